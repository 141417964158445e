// g711.cu — G.711 A-law / mu-law companding for batched RTP payloads (include/msb200dsp.h "G.711" section).
// Replaces the per-sample loops of the reference's codec filters:
//   decoders /root/reference/src/audiofilters/alaw.c:199-211, ulaw.c (same shape): Snack_Alaw2Lin / Snack_Mulaw2Lin
//   encoders alaw.c:56-94, ulaw.c: Snack_Lin2Alaw / Snack_Lin2Mulaw (g711.c:119-262)
// Pure integer bit manipulation, bit-exact (oracle/oracle_g711.c, pinned exhaustively against the unmodified g711.c).
// HBM-bound: 3 bytes per sample (1 code + 2 PCM). One thread handles 16 samples: one 16-byte code vector and two 16-byte
// PCM vectors, fully coalesced; the segment number comes from the leading-one position (clz), no tables, no branches
// on data. Tails and unaligned buffers take a scalar path.
#include "msb200_internal.h"

template <int LAW>
__device__ __forceinline__ int g711_dec1(unsigned c) {
	if (LAW == 0) { // Snack_Alaw2Lin
		c ^= 0x55u;
		const int seg = (int)(c & 0x70u) >> 4;
		int t = (int)(c & 0x0Fu) << 4;
		t = seg == 0 ? t + 8 : (t + 0x108) << max(seg - 1, 0);
		return (c & 0x80u) ? t : -t;
	} else { // Snack_Mulaw2Lin
		c = ~c & 0xFFu;
		const int t = (int)(((c & 0x0Fu) << 3) + 0x84u) << ((c & 0x70u) >> 4);
		return (c & 0x80u) ? (0x84 - t) : (t - 0x84);
	}
}
template <int LAW>
__device__ __forceinline__ unsigned g711_enc1(int pcm) { // pcm: sign-extended 16-bit sample
	if (LAW == 0) { // Snack_Lin2Alaw
		int x = pcm >> 3;
		const unsigned mask = x >= 0 ? 0xD5u : 0x55u;
		x = x < 0 ? ~x : x; // -x - 1
		const int seg = x < 32 ? 0 : 27 - __clz(x); // leading-one position - 4
		const unsigned q = (unsigned)(x >> (seg < 2 ? 1 : seg)) & 0xFu;
		return (((unsigned)seg << 4) | q) ^ mask;
	} else { // Snack_Lin2Mulaw
		int x = pcm >> 2;
		const unsigned mask = x < 0 ? 0x7Fu : 0xFFu;
		x = min(abs(x), 8159) + 33;
		const int seg = x < 64 ? 0 : 26 - __clz(x); // leading-one position - 5
		const unsigned v = (((unsigned)seg << 4) | ((unsigned)(x >> (seg + 1)) & 0xFu)) ^ mask;
		return seg >= 8 ? (0x7Fu ^ mask) : v;
	}
}

template <int LAW>
__global__ void __launch_bounds__(256) g711_decode_kernel(const uint8_t *__restrict__ code, short *__restrict__ pcm, size_t n, int vec) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (vec) {
		const size_t nv = n / 16;
		auto expand = [&](const uint4 c, size_t at) {
			const unsigned w[4] = {c.x, c.y, c.z, c.w};
			unsigned o[8];
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const unsigned lo = (w[k >> 1] >> (16 * (k & 1))) & 0xFFu, hi = (w[k >> 1] >> (16 * (k & 1) + 8)) & 0xFFu;
				o[k] = ((unsigned)g711_dec1<LAW>(lo) & 0xFFFFu) | ((unsigned)g711_dec1<LAW>(hi) << 16);
			}
			uint4 *dst = reinterpret_cast<uint4 *>(pcm) + 2 * at;
			__stcs(dst, make_uint4(o[0], o[1], o[2], o[3])); // streaming stores: the output is not re-read by this kernel
			__stcs(dst + 1, make_uint4(o[4], o[5], o[6], o[7]));
		};
		// four independent 16-byte loads in flight per thread before any of them is consumed
		for (; i + 3 * stride < nv; i += 4 * stride) {
			const uint4 c0 = __ldcs(reinterpret_cast<const uint4 *>(code) + i), c1 = __ldcs(reinterpret_cast<const uint4 *>(code) + i + stride);
			const uint4 c2 = __ldcs(reinterpret_cast<const uint4 *>(code) + i + 2 * stride), c3 = __ldcs(reinterpret_cast<const uint4 *>(code) + i + 3 * stride);
			expand(c0, i);
			expand(c1, i + stride);
			expand(c2, i + 2 * stride);
			expand(c3, i + 3 * stride);
		}
		for (; i < nv; i += stride) expand(__ldcs(reinterpret_cast<const uint4 *>(code) + i), i);
		i = nv * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; // tail
	}
	for (; i < n; i += stride) pcm[i] = (short)g711_dec1<LAW>(code[i]);
}
template <int LAW>
__global__ void __launch_bounds__(256) g711_encode_kernel(const short *__restrict__ pcm, uint8_t *__restrict__ code, size_t n, int vec) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (vec) {
		const size_t nv = n / 16;
		auto compress = [&](const uint4 a, const uint4 b, size_t at) {
			const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
			unsigned o[4] = {0, 0, 0, 0};
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const int lo = (int)(short)(w[k] & 0xFFFFu), hi = (int)w[k] >> 16;
				o[k >> 1] |= (g711_enc1<LAW>(lo) | (g711_enc1<LAW>(hi) << 8)) << (16 * (k & 1));
			}
			__stcs(reinterpret_cast<uint4 *>(code) + at, make_uint4(o[0], o[1], o[2], o[3]));
		};
		const uint4 *src = reinterpret_cast<const uint4 *>(pcm);
		for (; i + stride < nv; i += 2 * stride) { // four independent 16-byte loads in flight per thread
			const uint4 a0 = __ldcs(src + 2 * i), b0 = __ldcs(src + 2 * i + 1);
			const uint4 a1 = __ldcs(src + 2 * (i + stride)), b1 = __ldcs(src + 2 * (i + stride) + 1);
			compress(a0, b0, i);
			compress(a1, b1, i + stride);
		}
		for (; i < nv; i += stride) compress(__ldcs(src + 2 * i), __ldcs(src + 2 * i + 1), i);
		i = nv * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	}
	for (; i < n; i += stride) code[i] = (uint8_t)g711_enc1<LAW>((int)pcm[i]);
}

static int g711_grid(msb200_ctx *ctx, size_t n) {
	const size_t want = (n / 16 + 255) / 256 + 1;
	const size_t cap = (size_t)ctx->sm_count * 8; // grid-stride: 8 CTAs of 256 threads per SM (full occupancy at 32 registers)
	return (int)(want < cap ? want : cap);
}

extern "C" {

int msb200_g711_decode_dev(msb200_ctx *ctx, int law, const void *d_code, void *d_pcm, size_t n) {
	MSB200_CHECK_ARG(ctx && d_code && d_pcm && (law == MSB200_G711_ALAW || law == MSB200_G711_ULAW));
	if (n == 0) return MSB200_OK;
	const int vec = ((uintptr_t)d_code % 16) == 0 && ((uintptr_t)d_pcm % 16) == 0;
	if (law == MSB200_G711_ALAW)
		MSB200_LAUNCH(ctx, g711_decode_kernel<0>, g711_grid(ctx, n), 256, 0, (const uint8_t *)d_code, (short *)d_pcm, n, vec);
	else
		MSB200_LAUNCH(ctx, g711_decode_kernel<1>, g711_grid(ctx, n), 256, 0, (const uint8_t *)d_code, (short *)d_pcm, n, vec);
	return MSB200_OK;
}
int msb200_g711_encode_dev(msb200_ctx *ctx, int law, const void *d_pcm, void *d_code, size_t n) {
	MSB200_CHECK_ARG(ctx && d_code && d_pcm && (law == MSB200_G711_ALAW || law == MSB200_G711_ULAW));
	if (n == 0) return MSB200_OK;
	const int vec = ((uintptr_t)d_code % 16) == 0 && ((uintptr_t)d_pcm % 16) == 0;
	if (law == MSB200_G711_ALAW)
		MSB200_LAUNCH(ctx, g711_encode_kernel<0>, g711_grid(ctx, n), 256, 0, (const short *)d_pcm, (uint8_t *)d_code, n, vec);
	else
		MSB200_LAUNCH(ctx, g711_encode_kernel<1>, g711_grid(ctx, n), 256, 0, (const short *)d_pcm, (uint8_t *)d_code, n, vec);
	return MSB200_OK;
}

// host-buffer variants: H2D -> kernel -> D2H -> sync on the context's stream (staging grows on demand, kept per context
// user: the buffers belong to the call, so two threads must not share one ctx without external locking — as for banks)
static int g711_host(msb200_ctx *ctx, int law, int encode, const void *in, void *out, size_t n) {
	MSB200_CHECK_ARG(ctx && in && out);
	if (n == 0) return MSB200_OK;
	const size_t in_bytes = encode ? n * 2 : n, out_bytes = encode ? n : n * 2;
	void *d_in = nullptr, *d_out = nullptr;
	cudaStream_t s = ctx->stream;
	MSB200_CUDA(cudaMallocAsync(&d_in, in_bytes, s));
	MSB200_CUDA(cudaMallocAsync(&d_out, out_bytes, s));
	MSB200_CUDA(cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, s));
	int r = encode ? msb200_g711_encode_dev(ctx, law, d_in, d_out, n) : msb200_g711_decode_dev(ctx, law, d_in, d_out, n);
	if (r == MSB200_OK) {
		cudaError_t e = cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s);
		if (e != cudaSuccess) {
			msb200_set_error("g711: D2H copy failed: %s", cudaGetErrorString(e));
			r = MSB200_ECUDA;
		}
	}
	cudaFreeAsync(d_in, s);
	cudaFreeAsync(d_out, s);
	MSB200_HOST_DONE(ctx);
	return r;
}
int msb200_g711_decode(msb200_ctx *ctx, int law, const uint8_t *code, int16_t *pcm, size_t n) {
	return g711_host(ctx, law, 0, code, pcm, n);
}
int msb200_g711_encode(msb200_ctx *ctx, int law, const int16_t *pcm, uint8_t *code, size_t n) {
	return g711_host(ctx, law, 1, pcm, code, n);
}

} // extern "C"
