// msb200_internal.h — shared internals of libmsb200dsp.so (not installed; the public ABI is include/msb200dsp.h)
#pragma once
#include "msb200dsp.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct msb200_ctx {
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	bool owns_stream = true;
	cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
	uint64_t launches = 0;
	void *flush_buf = nullptr; // > L2, written by msb200_flush_l2
	size_t flush_bytes = 0;
	bool defer_sync = false; // msb200_ctx_set_deferred_sync: host-buffer entry points return after enqueueing
};

// what a host-buffer entry point ends with: wait for the stream, unless the caller collects several calls under one
// msb200_ctx_sync()
#define MSB200_HOST_DONE(ctx)                                                                                          \
	do {                                                                                                               \
		if (!(ctx)->defer_sync) MSB200_CUDA(cudaStreamSynchronize((ctx)->stream));                                     \
	} while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the FUNCTION (per device), not to the bank that sets it: banks of
// different geometry share it, so it is only ever raised. Returns a cudaError_t.
cudaError_t msb200_smem_optin(const void *func, int device, size_t bytes);
#define MSB200_SMEM_OPTIN(func, ctx, bytes) MSB200_CUDA(msb200_smem_optin((const void *)(func), (ctx)->device, (bytes)))

void msb200_set_error(const char *fmt, ...);

#define MSB200_CUDA(expr)                                                                                              \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess) {                                                                                       \
			msb200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));                    \
			return MSB200_ECUDA;                                                                                       \
		}                                                                                                              \
	} while (0)

#define MSB200_CHECK_ARG(cond)                                                                                         \
	do {                                                                                                               \
		if (!(cond)) {                                                                                                 \
			msb200_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);                                \
			return MSB200_EINVAL;                                                                                      \
		}                                                                                                              \
	} while (0)

// every kernel launch goes through this so the launch counter is honest
#define MSB200_LAUNCH(ctx, kernel, grid, block, smem, ...)                                                             \
	do {                                                                                                               \
		kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                                               \
		(ctx)->launches++;                                                                                             \
		cudaError_t _e = cudaGetLastError();                                                                           \
		if (_e != cudaSuccess) {                                                                                       \
			msb200_set_error("%s:%d: launch %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(_e));           \
			return MSB200_ECUDA;                                                                                       \
		}                                                                                                              \
	} while (0)

static inline int msb200_div_up(int a, int b) {
	return (a + b - 1) / b;
}

// device scratch that grows on demand (per object staging of host-path calls)
struct msb200_devbuf {
	void *p = nullptr;
	size_t cap = 0;
	int reserve(size_t bytes) {
		if (bytes <= cap) return MSB200_OK;
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
		cudaError_t e = cudaMalloc(&p, bytes);
		if (e != cudaSuccess) {
			msb200_set_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
			return MSB200_ENOMEM;
		}
		cap = bytes;
		return MSB200_OK;
	}
	void release() {
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

// ---- internal launchers used by chain.cu (ring-buffer addressing for the re-framing between stages); ring == 0: linear
int msb200i_resample_launch(msb200_resample *r, const void *d_in, int in_frames, int in_stride, void *d_out,
                            int out_stride, int ring_off, int ring_cap, int *out_frames);
int msb200i_resample_launch_pair(msb200_resample *a, msb200_resample *b, const void *d_in_a, const void *d_in_b, int in_frames,
                                 int in_stride, void *d_out_a, void *d_out_b, int out_stride, int ring_off, int ring_cap, int *out_frames);
int msb200i_aec_launch(msb200_aec *a, const void *d_mic, const void *d_ref, int in_stride, int in_frame0,
                       int in_ring_frames, void *d_out, int out_stride, int out_frame0, int out_ring_frames, int nframes,
                       const int *d_counts = nullptr);
// d_copy_out (optional, with `copied`): the processed blocks also go to d_copy_out[stream][blk * nsamples ...] (rows copy_stride
// samples apart) when the lane kernel runs; *copied tells whether it did
int msb200i_volume_launch(msb200_volume *v, void *d_io, int nsamples, int stride, int nblocks, int block0, int ring_blocks,
                          const int *d_counts = nullptr, void *d_copy_out = nullptr, int copy_stride = 0, int *copied = nullptr);
int msb200i_mixer_launch(msb200_mixer *m, const void *d_in, long in_pin_stride, const void *d_present, void *d_out);
int msb200i_packed422_to_i420(msb200_ctx *ctx, int n_frames, const void *d_src, int w, int h, int uyvy, void *d_dst);
int msb200i_rgb24_to_i420(msb200_ctx *ctx, int n_frames, const void *d_src, int w, int h, int fmt, void *d_dst, int x86_vertical = 0);
