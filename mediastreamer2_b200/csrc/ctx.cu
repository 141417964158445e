// ctx.cu — context, memory helpers, device stopwatch (include/msb200dsp.h "context" section)
#include "msb200_internal.h"

static thread_local char g_err[512] = "";

void msb200_set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

#include <map>
#include <mutex>
cudaError_t msb200_smem_optin(const void *func, int device, size_t bytes) {
	static std::mutex mu;
	static std::map<std::pair<const void *, int>, size_t> cur;
	if (bytes <= 48 * 1024) return cudaSuccess;
	std::lock_guard<std::mutex> g(mu);
	size_t &have = cur[std::make_pair(func, device)];
	if (bytes <= have) return cudaSuccess;
	cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
	if (e == cudaSuccess) have = bytes;
	return e;
}

extern "C" {

int msb200_version(void) {
	return 100; // 0.1.0
}

const char *msb200_last_error(void) {
	return g_err;
}

static int ctx_create_impl(int device_ordinal, bool own, cudaStream_t ext, msb200_ctx **out) {
	MSB200_CHECK_ARG(out != nullptr);
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		// by design there is no CPU fallback: the named filters only exist on the GPU
		msb200_set_error("no CUDA device available (%s); libmsb200dsp has no CPU fallback",
		                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
		return MSB200_ENODEV;
	}
	MSB200_CHECK_ARG(device_ordinal >= 0 && device_ordinal < n);
	MSB200_CUDA(cudaSetDevice(device_ordinal));
	cudaDeviceProp prop;
	MSB200_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
	if (prop.major < 10) {
		msb200_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device_ordinal, prop.major,
		                 prop.minor);
		return MSB200_ENODEV;
	}
	msb200_ctx *c = new msb200_ctx();
	c->device = device_ordinal;
	c->sm_count = prop.multiProcessorCount;
	c->owns_stream = own;
	if (own) MSB200_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	else c->stream = ext;
	MSB200_CUDA(cudaEventCreate(&c->ev_start));
	MSB200_CUDA(cudaEventCreate(&c->ev_stop));
	*out = c;
	return MSB200_OK;
}

int msb200_ctx_make_current(msb200_ctx *c) {
	MSB200_CHECK_ARG(c != nullptr);
	MSB200_CUDA(cudaSetDevice(c->device));
	return MSB200_OK;
}

int msb200_ctx_create(int device_ordinal, msb200_ctx **out) {
	return ctx_create_impl(device_ordinal, true, nullptr, out);
}
int msb200_ctx_create_on_stream(int device_ordinal, void *cuda_stream, msb200_ctx **out) {
	return ctx_create_impl(device_ordinal, false, (cudaStream_t)cuda_stream, out);
}

void msb200_ctx_destroy(msb200_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	if (c->flush_buf) cudaFree(c->flush_buf);
	if (c->ev_start) cudaEventDestroy(c->ev_start);
	if (c->ev_stop) cudaEventDestroy(c->ev_stop);
	if (c->stream && c->owns_stream) cudaStreamDestroy(c->stream);
	delete c;
}

int msb200_ctx_sync(msb200_ctx *c) {
	MSB200_CHECK_ARG(c);
	MSB200_CUDA(cudaStreamSynchronize(c->stream));
	return MSB200_OK;
}

int msb200_ctx_set_deferred_sync(msb200_ctx *c, int on) {
	MSB200_CHECK_ARG(c);
	if (!on && c->defer_sync) MSB200_CUDA(cudaStreamSynchronize(c->stream));
	c->defer_sync = on != 0;
	return MSB200_OK;
}

uint64_t msb200_ctx_launch_count(msb200_ctx *c) {
	return c ? c->launches : 0;
}

int msb200_dev_alloc(msb200_ctx *c, size_t bytes, void **p) {
	MSB200_CHECK_ARG(c && p);
	MSB200_CUDA(cudaSetDevice(c->device));
	MSB200_CUDA(cudaMalloc(p, bytes ? bytes : 1));
	return MSB200_OK;
}
int msb200_dev_free(msb200_ctx *c, void *p) {
	MSB200_CHECK_ARG(c);
	MSB200_CUDA(cudaFree(p));
	return MSB200_OK;
}
int msb200_host_alloc_pinned(msb200_ctx *c, size_t bytes, void **p) {
	MSB200_CHECK_ARG(c && p);
	MSB200_CUDA(cudaSetDevice(c->device));
	MSB200_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
	return MSB200_OK;
}
int msb200_host_free_pinned(msb200_ctx *c, void *p) {
	MSB200_CHECK_ARG(c);
	MSB200_CUDA(cudaFreeHost(p));
	return MSB200_OK;
}
int msb200_memcpy_h2d(msb200_ctx *c, void *dev, const void *host, size_t bytes) {
	MSB200_CHECK_ARG(c);
	MSB200_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
	MSB200_CUDA(cudaStreamSynchronize(c->stream));
	return MSB200_OK;
}
int msb200_memcpy_d2h(msb200_ctx *c, void *host, const void *dev, size_t bytes) {
	MSB200_CHECK_ARG(c);
	MSB200_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
	MSB200_CUDA(cudaStreamSynchronize(c->stream));
	return MSB200_OK;
}
int msb200_memset_dev(msb200_ctx *c, void *dev, int value, size_t bytes) {
	MSB200_CHECK_ARG(c);
	MSB200_CUDA(cudaMemsetAsync(dev, value, bytes, c->stream));
	return MSB200_OK;
}
int msb200_timer_start(msb200_ctx *c) {
	MSB200_CHECK_ARG(c);
	MSB200_CUDA(cudaEventRecord(c->ev_start, c->stream));
	return MSB200_OK;
}
int msb200_timer_stop_ms(msb200_ctx *c, float *ms) {
	MSB200_CHECK_ARG(c && ms);
	MSB200_CUDA(cudaEventRecord(c->ev_stop, c->stream));
	MSB200_CUDA(cudaEventSynchronize(c->ev_stop));
	MSB200_CUDA(cudaEventElapsedTime(ms, c->ev_start, c->ev_stop));
	return MSB200_OK;
}
int msb200_ipc_export(msb200_ctx *c, void *dev_ptr, uint8_t handle[MSB200_IPC_HANDLE_BYTES]) {
	MSB200_CHECK_ARG(c && dev_ptr && handle);
	static_assert(sizeof(cudaIpcMemHandle_t) <= MSB200_IPC_HANDLE_BYTES, "handle size");
	cudaIpcMemHandle_t h;
	MSB200_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
	memset(handle, 0, MSB200_IPC_HANDLE_BYTES);
	memcpy(handle, &h, sizeof(h));
	return MSB200_OK;
}
int msb200_ipc_import(msb200_ctx *c, const uint8_t handle[MSB200_IPC_HANDLE_BYTES], void **dev_ptr) {
	MSB200_CHECK_ARG(c && handle && dev_ptr);
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	MSB200_CUDA(cudaSetDevice(c->device));
	MSB200_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return MSB200_OK;
}
int msb200_ipc_close(msb200_ctx *c, void *dev_ptr) {
	MSB200_CHECK_ARG(c && dev_ptr);
	MSB200_CUDA(cudaIpcCloseMemHandle(dev_ptr));
	return MSB200_OK;
}
int msb200_flush_l2(msb200_ctx *c) {
	MSB200_CHECK_ARG(c);
	if (!c->flush_buf) {
		c->flush_bytes = (size_t)256 << 20; // 2x the 126 MB L2
		MSB200_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
	}
	MSB200_CUDA(cudaMemsetAsync(c->flush_buf, 0x5a, c->flush_bytes, c->stream));
	return MSB200_OK;
}

} // extern "C"
