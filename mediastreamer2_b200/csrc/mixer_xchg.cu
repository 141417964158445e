// mixer_xchg.cu — the ONE exchange step of the hot path (BASELINE cfg3): an MSAudioMixer conference whose pins
// live on different GPUs. Two implementations behind the C ABI:
//
//  (1) msb200_comm_* / msb200_mixer_process_striped_dev: partial sums -> ncclAllReduce(int32, SUM) -> finish. NCCL is
//      dlopen'ed (no link-time dependency for 1-GPU users); the collective runs on the context's stream.
//  (2) msb200_mixer_xchg_*: ONE kernel per tick over NVLink peer memory. Every CTA
//        A. builds the int32 partial sums of its (room, column) items over the LOCAL pins and PUSHES them (16-byte stores)
//           into slot [parity][my_rank] of every rank's receive area — remote stores are posted, nothing waits on NVLink;
//        B. fences, then publishes `epoch` in flags[my_rank][cta] of every rank;
//        C. waits until the LOCAL flags[src][cta] of every source rank reach `epoch` (polling local memory only),
//        D. adds the world's partial sums from its LOCAL receive area and emits sat(total - own) for the local pins.
//      No CTA waits before it has pushed, the grid is sized to be co-resident, and the item -> CTA map is identical on all
//      ranks, so CTA c only ever waits for CTA c of the peers: no deadlock, no grid-wide barrier, no NCCL launch.
//      Integer sums are order-independent: bit-exact with the single-GPU mixer (audiomixer.c:288-346, :113-130, :40-44).
//      Receive slots alternate with the tick parity: a peer can push tick t+2 only after it finished tick t+1, which needs
//      our tick t+1 push, which our stream issues after our tick t kernel completed — so slot `parity` is never overwritten
//      while it is still being read.
#include "mixer_internal.cuh"

#include <dlfcn.h>

// ============================================================================================ fused exchange kernel
struct XchgPeers {
	int *recv[MSB200_MAX_PEERS];       // rank g's receive area as mapped here: [parity][src][rooms * nwords] int32
	unsigned *flags[MSB200_MAX_PEERS]; // rank g's flags as mapped here: [src][cta]
	int n, rank, parity;
	unsigned epoch;
	unsigned *error;                   // local: count of flag waits that ran into the timeout
	unsigned long long timeout_ns;
};

__device__ __forceinline__ int4 ld_volatile_v4(const int *p) { // written by a peer: never from a stale L1 line
	int4 v;
	asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

constexpr int XCHG_THREADS = 256;

__global__ void __launch_bounds__(XCHG_THREADS) mixer_xchg_kernel(const short *__restrict__ in, const uint8_t *__restrict__ present,
                                                                  const float *__restrict__ gain,
                                                                  const uint8_t *__restrict__ active, short *__restrict__ out,
                                                                  int n_rooms, int n_pins, int nwords, XchgPeers x) {
	const int nvec = nwords >> 2; // 4 samples per item: one int2 of s16 in, one int4 of int32 on the wire
	const long total = (long)n_rooms * nvec;
	const long stride = (long)gridDim.x * XCHG_THREADS;
	const size_t slot_words = (size_t)n_rooms * nwords;
	// ---- A: partial sums of the local pins, pushed to every rank's slot [parity][rank]
	for (long item = (long)blockIdx.x * XCHG_THREADS + threadIdx.x; item < total; item += stride) {
		const int room = (int)(item / nvec), col = (int)(item % nvec);
		const size_t chan0 = (size_t)room * n_pins;
		const int2 *inv = reinterpret_cast<const int2 *>(in) + chan0 * nvec + col;
		int sum[4] = {0, 0, 0, 0};
		for (int p = 0; p < n_pins; ++p) {
			if (!present[chan0 + p] || !active[chan0 + p]) continue; // channel_process_in :78-90
			const float g = gain[chan0 + p];
			int s[4];
			unpack_s16<4>(inv[(size_t)p * nvec], s);
#pragma unroll
			for (int k = 0; k < 4; ++k) sum[k] += mix_contrib(s[k], g);
		}
		const size_t off = ((size_t)(x.parity * x.n + x.rank)) * slot_words + (size_t)room * nwords + (size_t)col * 4;
		const int4 v = make_int4(sum[0], sum[1], sum[2], sum[3]);
		for (int g = 0; g < x.n; ++g) *reinterpret_cast<int4 *>(x.recv[g] + off) = v;
	}
	// ---- B: release — every thread's pushes are ordered before the CTA's flag stores
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x < x.n)
		*reinterpret_cast<volatile unsigned *>(x.flags[threadIdx.x] + (size_t)x.rank * gridDim.x + blockIdx.x) = x.epoch;
	// ---- C: acquire — the same CTA of every source rank has pushed (local polling; the wait is bounded)
	if (threadIdx.x < x.n) {
		const volatile unsigned *f = x.flags[x.rank] + (size_t)threadIdx.x * gridDim.x + blockIdx.x;
		const unsigned long long t0 = global_ns();
		while ((int)(*f - x.epoch) < 0) {
			if (global_ns() - t0 > x.timeout_ns) {
				atomicAdd(x.error, 1u);
				break;
			}
		}
		__threadfence_system();
	}
	__syncthreads();
	// ---- D: total = sum over the world's slots (local memory), outputs of the local pins (channel_process_out :113-130)
	const int *mine = x.recv[x.rank] + (size_t)x.parity * x.n * slot_words;
	for (long item = (long)blockIdx.x * XCHG_THREADS + threadIdx.x; item < total; item += stride) {
		const int room = (int)(item / nvec), col = (int)(item % nvec);
		const size_t chan0 = (size_t)room * n_pins;
		const size_t off = (size_t)room * nwords + (size_t)col * 4;
		int sum[4] = {0, 0, 0, 0};
		for (int g = 0; g < x.n; ++g) {
			const int4 v = ld_volatile_v4(mine + (size_t)g * slot_words + off);
			sum[0] += v.x; sum[1] += v.y; sum[2] += v.z; sum[3] += v.w;
		}
		const int2 *inv = reinterpret_cast<const int2 *>(in) + chan0 * nvec + col;
		int2 *outv = reinterpret_cast<int2 *>(out) + chan0 * nvec + col;
		for (int p = 0; p < n_pins; ++p) {
			int o[4];
			if (active[chan0 + p] && present[chan0 + p]) {
				const float g = gain[chan0 + p];
				int s[4];
				unpack_s16<4>(inv[(size_t)p * nvec], s);
#pragma unroll
				for (int k = 0; k < 4; ++k) o[k] = mix_sat(sum[k] - mix_contrib(s[k], g));
			} else {
#pragma unroll
				for (int k = 0; k < 4; ++k) o[k] = mix_sat(sum[k]);
			}
			outv[(size_t)p * nvec] = pack_s16<4>(o);
		}
	}
}

struct msb200_mixer_xchg {
	msb200_mixer *m;
	int rank, world, grid;
	uint32_t epoch;
	size_t slot_words, recv_bytes, flags_bytes, block_bytes;
	uint8_t *block;                         // one allocation: receive area | flags | error word
	uint8_t *mapped[MSB200_MAX_PEERS];      // every rank's block as addressable from this rank (own: block)
	bool ipc_mapped[MSB200_MAX_PEERS];
	bool connected;
	unsigned long long timeout_ns;
};

static int xchg_grid(const msb200_mixer *m) {
	// identical on every rank (same bank shape, same constant): the CTA index is part of the flag address. 4 CTAs of 256
	// threads per SM of a 148-SM part are co-resident by a wide margin (the kernel uses < 40 registers, no shared memory).
	const long total = (long)m->n_rooms * (m->nwords / 4);
	const long need = (total + XCHG_THREADS - 1) / XCHG_THREADS;
	const long cap = 148L * 4;
	return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

extern "C" {

int msb200_mixer_xchg_create(msb200_mixer *m, int rank, int world, msb200_mixer_xchg **out) {
	MSB200_CHECK_ARG(m && out && world >= 1 && world <= MSB200_MAX_PEERS && rank >= 0 && rank < world);
	MSB200_CHECK_ARG(m->conf_mode && m->nwords % 4 == 0);
	msb200_mixer_xchg *x = new msb200_mixer_xchg();
	x->m = m;
	x->rank = rank;
	x->world = world;
	x->grid = xchg_grid(m);
	x->epoch = 0;
	x->slot_words = (size_t)m->n_rooms * m->nwords;
	x->recv_bytes = 2 * (size_t)world * x->slot_words * sizeof(int);
	x->flags_bytes = (((size_t)world * x->grid * sizeof(unsigned)) + 255) & ~(size_t)255;
	x->block_bytes = x->recv_bytes + x->flags_bytes + 256;
	x->connected = false;
	x->timeout_ns = 2000000000ull; // 2 s: a peer that is this late is gone; outputs of the tick are then undefined
	if (const char *e = getenv("MSB200_XCHG_TIMEOUT_MS")) x->timeout_ns = strtoull(e, nullptr, 10) * 1000000ull;
	memset(x->mapped, 0, sizeof(x->mapped));
	memset(x->ipc_mapped, 0, sizeof(x->ipc_mapped));
	cudaError_t e = cudaSetDevice(m->ctx->device);
	if (e == cudaSuccess) e = cudaMalloc(&x->block, x->block_bytes);
	if (e == cudaSuccess) e = cudaMemsetAsync(x->block, 0, x->block_bytes, m->ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(m->ctx->stream);
	if (e != cudaSuccess) {
		msb200_set_error("mixer_xchg_create: %s", cudaGetErrorString(e));
		delete x;
		return MSB200_ECUDA;
	}
	x->mapped[rank] = x->block;
	*out = x;
	return MSB200_OK;
}

int msb200_mixer_xchg_export(msb200_mixer_xchg *x, uint8_t handle[MSB200_IPC_HANDLE_BYTES]) {
	MSB200_CHECK_ARG(x && handle);
	return msb200_ipc_export(x->m->ctx, x->block, handle);
}

int msb200_mixer_xchg_connect(msb200_mixer_xchg *x, const uint8_t *handles) {
	MSB200_CHECK_ARG(x && handles && !x->connected);
	for (int g = 0; g < x->world; ++g) {
		if (g == x->rank) continue;
		void *p = nullptr;
		int r = msb200_ipc_import(x->m->ctx, handles + (size_t)g * MSB200_IPC_HANDLE_BYTES, &p);
		if (r) return r;
		x->mapped[g] = (uint8_t *)p;
		x->ipc_mapped[g] = true;
	}
	x->connected = true;
	return MSB200_OK;
}

int msb200_mixer_xchg_connect_local(msb200_mixer_xchg *x, msb200_mixer_xchg *const *all) {
	MSB200_CHECK_ARG(x && all && !x->connected);
	for (int g = 0; g < x->world; ++g) {
		MSB200_CHECK_ARG(all[g] && all[g]->rank == g && all[g]->world == x->world && all[g]->block_bytes == x->block_bytes);
		if (g == x->rank) continue;
		const int peer_dev = all[g]->m->ctx->device;
		if (peer_dev != x->m->ctx->device) {
			MSB200_CUDA(cudaSetDevice(x->m->ctx->device));
			cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if (e != cudaSuccess) {
				msb200_set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", x->m->ctx->device, peer_dev, cudaGetErrorString(e));
				return MSB200_ECUDA;
			}
		}
		x->mapped[g] = all[g]->block;
	}
	x->connected = true;
	return MSB200_OK;
}

int msb200_mixer_xchg_process_dev(msb200_mixer_xchg *x, const void *d_in, const void *d_present, void *d_out) {
	MSB200_CHECK_ARG(x && d_in && d_present && d_out && x->connected);
	MSB200_CHECK_ARG(((uintptr_t)d_in % 8) == 0 && ((uintptr_t)d_out % 8) == 0);
	msb200_mixer *m = x->m;
	int r = msb200i_mixer_upload(m);
	if (r) return r;
	XchgPeers p;
	memset(&p, 0, sizeof(p));
	for (int g = 0; g < x->world; ++g) {
		p.recv[g] = (int *)x->mapped[g];
		p.flags[g] = (unsigned *)(x->mapped[g] + x->recv_bytes);
	}
	p.n = x->world;
	p.rank = x->rank;
	p.parity = (int)(x->epoch & 1u);
	p.epoch = ++x->epoch;
	p.error = (unsigned *)(x->block + x->recv_bytes + x->flags_bytes);
	p.timeout_ns = x->timeout_ns;
	// the whole bank takes part (n_rooms, not `live`): the flag addresses depend on the grid, which must match on all ranks
	MSB200_LAUNCH(m->ctx, mixer_xchg_kernel, x->grid, XCHG_THREADS, 0, (const short *)d_in, (const uint8_t *)d_present,
	              m->d_gain, m->d_active, (short *)d_out, m->n_rooms, m->n_pins, m->nwords, p);
	return MSB200_OK;
}

int msb200_mixer_xchg_status(msb200_mixer_xchg *x, uint32_t *timeouts) {
	MSB200_CHECK_ARG(x && timeouts);
	MSB200_CUDA(cudaMemcpyAsync(timeouts, x->block + x->recv_bytes + x->flags_bytes, sizeof(uint32_t), cudaMemcpyDeviceToHost,
	                            x->m->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(x->m->ctx->stream));
	return MSB200_OK;
}

size_t msb200_mixer_xchg_wire_bytes_per_tick(msb200_mixer_xchg *x) {
	return x ? (size_t)(x->world - 1) * x->slot_words * sizeof(int) : 0;
}

void msb200_mixer_xchg_destroy(msb200_mixer_xchg *x) {
	if (!x) return;
	cudaSetDevice(x->m->ctx->device);
	cudaStreamSynchronize(x->m->ctx->stream);
	for (int g = 0; g < x->world; ++g)
		if (x->ipc_mapped[g] && x->mapped[g]) cudaIpcCloseMemHandle(x->mapped[g]);
	cudaFree(x->block);
	delete x;
}

} // extern "C"

// ============================================================================================ NCCL (dlopen'ed)
// Only what the conference exchange needs, declared here so that neither the build nor 1-GPU users need NCCL:
// ncclUniqueId is 128 opaque bytes, ncclInt32 = 2, ncclSum = 0 (nccl.h, stable since NCCL 2.0).
namespace {
typedef struct { char internal[128]; } nccl_unique_id;
typedef void *nccl_comm_t;
struct NcclApi {
	void *dl = nullptr;
	int (*GetUniqueId)(nccl_unique_id *) = nullptr;
	int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
	int (*CommDestroy)(nccl_comm_t) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	int (*GetVersion)(int *) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
	if (g_nccl.dl) return MSB200_OK;
	static_assert(sizeof(nccl_unique_id) == MSB200_COMM_ID_BYTES, "unique id size");
	void *dl = nullptr;
	const char *env = getenv("MSB200_NCCL_LIB");
	if (env && *env) dl = dlopen(env, RTLD_NOW | RTLD_LOCAL);
	if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy the host process (e.g. torch) already loaded
	if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
	if (!dl) dl = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
	if (!dl) {
		msb200_set_error("NCCL not found (set MSB200_NCCL_LIB to libnccl.so.2): %s", dlerror());
		return MSB200_ENODEV;
	}
	NcclApi a;
	a.dl = dl;
	*(void **)&a.GetUniqueId = dlsym(dl, "ncclGetUniqueId");
	*(void **)&a.CommInitRank = dlsym(dl, "ncclCommInitRank");
	*(void **)&a.CommDestroy = dlsym(dl, "ncclCommDestroy");
	*(void **)&a.AllReduce = dlsym(dl, "ncclAllReduce");
	*(void **)&a.GetErrorString = dlsym(dl, "ncclGetErrorString");
	*(void **)&a.GetVersion = dlsym(dl, "ncclGetVersion");
	if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
		msb200_set_error("the NCCL library lacks a required symbol");
		dlclose(dl);
		return MSB200_ENODEV;
	}
	g_nccl = a;
	return MSB200_OK;
}
} // namespace

#define MSB200_NCCL(expr)                                                                                              \
	do {                                                                                                               \
		int _r = (expr);                                                                                               \
		if (_r != 0) {                                                                                                 \
			msb200_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));                 \
			return MSB200_ECUDA;                                                                                       \
		}                                                                                                              \
	} while (0)

struct msb200_comm {
	msb200_ctx *ctx;
	nccl_comm_t comm;
	int rank, world;
};

extern "C" {

int msb200_comm_available(void) {
	return nccl_load() == MSB200_OK ? 1 : 0;
}

int msb200_comm_nccl_version(void) {
	int v = 0;
	if (nccl_load() != MSB200_OK || !g_nccl.GetVersion || g_nccl.GetVersion(&v) != 0) return 0;
	return v;
}

int msb200_comm_unique_id(uint8_t id[MSB200_COMM_ID_BYTES]) {
	MSB200_CHECK_ARG(id);
	int r = nccl_load();
	if (r) return r;
	nccl_unique_id u;
	MSB200_NCCL(g_nccl.GetUniqueId(&u));
	memcpy(id, &u, sizeof(u));
	return MSB200_OK;
}

int msb200_comm_create(msb200_ctx *ctx, const uint8_t id[MSB200_COMM_ID_BYTES], int rank, int world, msb200_comm **out) {
	MSB200_CHECK_ARG(ctx && id && out && world >= 1 && rank >= 0 && rank < world);
	int r = nccl_load();
	if (r) return r;
	MSB200_CUDA(cudaSetDevice(ctx->device));
	nccl_unique_id u;
	memcpy(&u, id, sizeof(u));
	nccl_comm_t c = nullptr;
	MSB200_NCCL(g_nccl.CommInitRank(&c, world, u, rank));
	msb200_comm *m = new msb200_comm();
	m->ctx = ctx;
	m->comm = c;
	m->rank = rank;
	m->world = world;
	*out = m;
	return MSB200_OK;
}

void msb200_comm_destroy(msb200_comm *c) {
	if (!c) return;
	cudaSetDevice(c->ctx->device);
	cudaStreamSynchronize(c->ctx->stream);
	if (g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
	delete c;
}

int msb200_comm_allreduce_sum_i32_dev(msb200_comm *c, void *d_buf, size_t count) {
	MSB200_CHECK_ARG(c && d_buf);
	MSB200_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, /*ncclInt32*/ 2, /*ncclSum*/ 0, c->comm, c->ctx->stream));
	return MSB200_OK;
}

int msb200_mixer_process_striped_dev(msb200_mixer *m, msb200_comm *c, const void *d_in, const void *d_present,
                                     void *d_sum_i32, void *d_out) {
	MSB200_CHECK_ARG(m && c && d_in && d_present && d_sum_i32 && d_out && m->ctx == c->ctx);
	int r = msb200_mixer_partial_dev(m, d_in, d_present, d_sum_i32);
	if (r) return r;
	if ((r = msb200_comm_allreduce_sum_i32_dev(c, d_sum_i32, (size_t)m->live * m->nwords))) return r;
	return msb200_mixer_finish_dev(m, d_in, d_present, d_sum_i32, d_out);
}

} // extern "C"
