// aec.cu — MSSpeexEC arithmetic: MDF two-path echo canceller + preprocessor (denoise / residual-echo suppression),
// one CTA per call stream, `nframes` consecutive frames per launch.
//
// Replaces, per framesize block, speex_echo_cancellation() + speex_preprocess_run() as called from
// /root/reference/src/audiofilters/speexec.c:297-298 with the configuration of speex_ec_preprocess() :188-216.
// speexdsp itself is not under /root/reference; the algorithm is restated in oracle/oracle_aec.c (see its header for
// provenance and "parity unpinned"). This file follows that restatement statement by statement; arithmetic that the
// oracle does in a fixed order (FFT butterflies, spectral products, filterbank sums, state recurrences) is done in the
// same order with -fmad=false, so the only deviations are the block-parallel reductions (inner products, Pey/Pyy,
// |W_j|^2), which sum in tree order instead of sequentially.
//
// Data layout in HBM (per stream, contiguous "page" so one CTA streams it linearly; SURVEY §8d: ~0.39 MB/frame):
//   X   [(M+1)][F] float2   far-end spectra, ring buffer over blocks (slot = (head + j) % (M+1); every stream keeps its
//                           own head in its state page, so streams may run different numbers of frames per launch) —
//                           replaces speex's per-frame memmove
//   W   [M][F] float2       background (adaptive) filter
//   FG  [M][F] float2       foreground filter
//   small state             window halves, previous error spectrum, power spectra, preprocessor state (~20 KB)
// Spectra use float2 per bin with bin 0 = (DC, Nyquist): 8-byte aligned, coalesced accesses (256 threads x 8 B = one
// 2 KB row per block). The hot loop over the M blocks reads X, FG, W once and writes W once per frame.
#include "msb200_internal.h"

#include <cmath>
#include <type_traits>

#define NB_BANDS 24

struct AecLayout {        // offsets in floats inside one stream's small-state page
	int xprev, E, last_y, power, power_1, Eh, Yh, prop, wnorm, scal, ints;
	int inbuf, outbuf, old_ps, noise, echo_noise, zeta, S, Smin, Stmp;
	int total;
};
enum { // scal[] indices
	SC_DAVG1, SC_DAVG2, SC_DVAR1, SC_DVAR2, SC_PEY, SC_PYY, SC_SUM_ADAPT, SC_LEAK, SC_MEMX, SC_MEMD, SC_MEME,
	SC_NOTCH0, SC_NOTCH1, SC_COUNT = 16
};
enum { IN_ADAPTED, IN_SATURATED, IN_SCREWED, IN_CANCEL_COUNT, IN_NB_ADAPT, IN_MIN_COUNT, IN_FG_PENDING, IN_HEAD, IN_COUNT = 8 };

struct AecParams {
	int F, N, M, L, log2L, rate;
	float spec_average, beta0, beta_max, notch_radius, preemph;
	int noise_suppress, echo_suppress, echo_suppress_active;
	float noise_floor; // (float)exp(.2302585f * noise_suppress)
	AecLayout lay;
	size_t x_stride, w_stride; // float2 per stream for X and for W / FG
	// bank-wide constant tables (device pointers)
	const float2 *tw;   // [L/2]  (cos, sin)(2 pi j / L)
	const float2 *spl;  // [L+1]  (cos, sin)(2 pi k / N)
	const float *window;  // [N] hann
	const float *pwindow; // [N] preprocessor conj window
	const int *bank_left; // [F]
	const float *filter_left, *filter_right; // [F]
	const int *band_start; // [NB_BANDS+1] first bin whose left band is b
	const float *prop0;    // [M] initial proportional weights
};

// ------------------------------------------------------------------------------------------------ cp.async (LDGSTS)
// The block pass keeps AEC_STAGES - 1 blocks (X_{j+1}, FG_j, W_j: 3 rows of 8F bytes each) in flight per CTA.
// Measured on B200 (profiles/r2c_aec_kernel_deep_pipe.*): going from 2 to 5 blocks in flight (-DAEC_STAGES=6
// -DAEC_OWN_STAGES=4: two extra ring slots aliased onto scratch that is dead during the pass, still 4 CTAs per SM) changes
// nothing — 851 vs 832 us, long-scoreboard stalls are 9 % of the warp latency either way. The pass is not latency bound;
// the kernel sits on the L1 / shared-memory pipe (65 %), on CTA barriers (28 % of warp latency) and on issue slots (59 %).
// The shallow ring stays the default (less shared memory); the deep one remains a build option for other shapes.
#ifndef AEC_STAGES
#define AEC_STAGES 3
#define AEC_OWN_STAGES 2 // ring slots with storage of their own; slots [AEC_OWN_STAGES, AEC_STAGES) alias dead scratch
#endif
// Occupancy, measured A/B in one session on B200 (MSB200_AEC_CTAS, 200 ticks of 4096 streams each): 4 CTAs of 256 threads
// per SM at 64 registers: 0.784 ms per launch; 5 CTAs at 48 registers (ptxas spills 20 bytes; shared memory cut to 43 KB by
// aliasing the third ring slot onto FFT scratch and loading the late state vectors into registers): 0.801 ms — the extra
// warps do not pay for the coarser wave quantisation (4096 CTAs = 5.5 waves of 740 instead of 6.9 of 592); 40 registers:
// 0.839 ms. The default stays 4.
#define AEC_CTAS_PER_SM_256 4
#define AEC_DEFAULT_SKEW_US 6 // measured after the tight pass went in (one box, 4096 streams, steady state): 0 -> 0.7678 ms per launch, 4 / 5 / 6 / 8 / 10 us -> 0.7617 / 0.7601 / 0.7593 / 0.7600 / 0.7611
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
	const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
	asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait() {
	asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------ barriers
// The F "main" threads (one per bin) synchronise on named barrier 1, never on barrier 0: the 48 kHz kernel carries a
// NINTH warp (the serial warp, see the kernel) that must not take part in their barriers. Barriers 2 - 4 are the
// producer / consumer hand-offs between that warp and the main threads (arrive on one side, sync on the other).
template <int NT> __device__ __forceinline__ void bar_main() {
	asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}
template <int NT> __device__ __forceinline__ int bar_main_or(int pred) {
	int r;
	asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.or.pred q, 1, %2, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
	             : "=r"(r)
	             : "r"(pred), "n"(NT)
	             : "memory");
	return r;
}
template <int ID, int NT> __device__ __forceinline__ void bar_arrive() { // prior shared-memory writes first
	__threadfence_block();
	asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NT) : "memory");
}
template <int ID, int NT> __device__ __forceinline__ void bar_wait() {
	asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory");
}
#define MAIN_SYNC() bar_main<(1 << LOG2L)>()
enum { BAR_INPUT = 2, BAR_DEEMPH_IN = 3, BAR_DEEMPH_OUT = 4, BAR_RING_OWN = 5, BAR_RING_ALIAS = 6 };

// ------------------------------------------------------------------------------------------------ bulk copies (TMA)
// TMA builds: the serial warp's lane 0 is also the PRODUCER of the block pass. It moves every 2 KB row X_{j+1}, FG_j, W_j
// with one cp.async.bulk (UBLKCP) into the ring and the per-bin threads wait on the slot's "full" mbarrier, instead of
// every thread copying its own 8 bytes with LDGSTS (3 LDGSTS per thread and block were 2/3 of the pass's load on the
// shared-memory pipe, the busiest unit of the kernel, plus the address bookkeeping). A slot goes back to the producer
// through its "empty" mbarrier (one arrival per warp of per-bin threads).
__device__ __forceinline__ unsigned smem_addr(const void *p) {
	return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
	asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(bar),
	             "r"(parity)
	             : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
	             "r"(bytes), "r"(bar)
	             : "memory");
}
// generic-proxy accesses (plain loads / stores, LDGSTS) before, async-proxy accesses (the bulk copies) after
__device__ __forceinline__ void fence_proxy_async_all() {
	asm volatile("fence.proxy.async;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ block helpers
// All helpers are called by every main thread of the CTA (threadIdx.x < F).

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
// deterministic block sum: shuffle tree inside each warp, then warp partials added in warp order. `red` = smem[32]
template <int LOG2L> __device__ float block_sum(float v, float *red) {
	constexpr int nw = ((1 << LOG2L) + 31) >> 5;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	v = warp_sum(v);
	MAIN_SYNC(); // protect `red` from the previous use
	if (lane == 0) red[warp] = v;
	MAIN_SYNC();
	float s = 0.f;
	for (int w = 0; w < nw; ++w) s += red[w];
	return s;
}
// three block sums sharing one pair of barriers; each sum has exactly block_sum()'s order. blockDim.x <= 256
template <int LOG2L> __device__ void block_sum3(float &a, float &b, float &c, float *red) {
	constexpr int nw = ((1 << LOG2L) + 31) >> 5;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	a = warp_sum(a);
	b = warp_sum(b);
	c = warp_sum(c);
	MAIN_SYNC();
	if (lane == 0) {
		red[warp] = a;
		red[8 + warp] = b;
		red[16 + warp] = c;
	}
	MAIN_SYNC();
	float sa = 0.f, sb = 0.f, sc3 = 0.f;
	for (int w = 0; w < nw; ++w) {
		sa += red[w];
		sb += red[8 + w];
		sc3 += red[16 + w];
	}
	a = sa;
	b = sb;
	c = sc3;
}
template <int LOG2L> __device__ int block_any(int pred) {
	return bar_main_or<(1 << LOG2L)>(pred);
}

// complex radix-2 Stockham FFT of length L = 1 << LOG2L over shared memory. Thread b < L/2 of a group computes one
// whole butterfly per stage: inputs x[b], x[b + L/2] (autosort: the same two slots at every stage), outputs
// y[q + 2sp] = a + c and y[q + 2sp + s] = (a - c) * w^p with q = b mod s, p = b div s. Arithmetic per output is
// oracle/oracle_aec.c:cfft()'s, operation for operation. Half the CTA works per transform, so a pair of transforms
// (cfft_pair) costs the same instructions and barriers as one.
// One thread carries FOUR points through TWO consecutive stages in registers: the pair of stage-st butterflies (u, u + H/2)
// feeds exactly the pair of stage-(st+1) butterflies (u + ps, u + ps + s), so the intermediate values never touch shared
// memory and every other CTA barrier disappears — while each value is still produced by the very same operations on the very
// same operands as in the stage-by-stage form (bit-identical spectra: test_aec_first_frames_match_oracle, probe "X").
// Thread u < H/2 of a group works; outputs land at u + 3 ps + {0, s, 2s, 3s}. An odd stage count ends with one plain stage.
template <int LOG2L>
__device__ __forceinline__ void cfft_bfly(float2 *x, float2 *y, const float2 *tw, int sign, bool active) {
	constexpr int H = 1 << (LOG2L - 1), Q = H >> 1;
	const int b = threadIdx.x & (H - 1);
	auto bfly = [&](const float2 a, const float2 c, const float2 w, float2 &o0, float2 &o1) {
		const float wr = w.x, wi = sign < 0 ? -w.y : w.y;
		o0.x = a.x + c.x;
		o0.y = a.y + c.y;
		const float dr = a.x - c.x, di = a.y - c.y;
		o1.x = dr * wr - di * wi;
		o1.y = dr * wi + di * wr;
	};
#ifdef AEC_FFT_PLAIN // A/B build: one stage per barrier, as in round 1
	constexpr int DOUBLE_END = 0;
#else
	constexpr int DOUBLE_END = LOG2L & ~1;
#endif
#pragma unroll
	for (int st = 0; st < DOUBLE_END; st += 2) {
		if (active && b < Q) {
			const int s = 1 << st;
			const int ps = b & ~(s - 1); // p << st
			float2 y0, y1, y0b, y1b, z0, z1, z2, z3;
			bfly(x[b], x[b + H], tw[ps], y0, y1);               // stage st, butterfly b
			bfly(x[b + Q], x[b + Q + H], tw[ps + Q], y0b, y1b); // stage st, butterfly b + H/2
			const float2 w2 = tw[2 * ps];
			bfly(y0, y0b, w2, z0, z2);                          // stage st + 1, butterfly b + ps
			bfly(y1, y1b, w2, z1, z3);                          // stage st + 1, butterfly b + ps + s
			const int o = b + 3 * ps;
			y[o] = z0;
			y[o + s] = z1;
			y[o + 2 * s] = z2;
			y[o + 3 * s] = z3;
		}
		MAIN_SYNC();
		float2 *sw = x;
		x = y;
		y = sw;
	}
#pragma unroll
	for (int st = DOUBLE_END; st < LOG2L; ++st) {
		if (active) {
			const int s = 1 << st;
			const int ps = b & ~(s - 1);
			float2 o0, o1;
			bfly(x[b], x[b + H], tw[ps], o0, o1);
			const int oi = b + ps;
			y[oi] = o0;
			y[oi + s] = o1;
		}
		MAIN_SYNC();
		float2 *sw = x;
		x = y;
		y = sw;
	}
}
// The same transform as ONE out-of-line function for all seven call sites of a frame (AEC_FFT_INLINE builds keep the inlined
// copies): the inlined butterflies were 1846 of the 48 kHz kernel's 7776 instructions (29 of 124 KB) — with four CTAs of an
// SM in four different phases of the frame the kernel ran at 69 % of the instruction cache hierarchy's request rate and
// 7.5 % of the warp stall samples were instruction fetches. Operands are shared-memory addresses (ld / st.shared, no generic
// pointers across the call); the inverse transform's conjugated twiddle is a sign-bit flip, as `-w.y` is.
__device__ __forceinline__ float2 lds_f2(unsigned a) {
	float2 v;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
	return v;
}
__device__ __forceinline__ void sts_f2(unsigned a, float2 v) {
	asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
template <int LOG2L> __device__ __noinline__ void cfft_core(unsigned x, unsigned y, unsigned tw, unsigned conj_mask, int active) {
	constexpr int H = 1 << (LOG2L - 1), Q = H >> 1;
	const int b = threadIdx.x & (H - 1);
	auto bfly = [&](const float2 a, const float2 c, const float2 w, float2 &o0, float2 &o1) {
		const float wr = w.x, wi = __int_as_float(__float_as_int(w.y) ^ (int)conj_mask);
		o0.x = a.x + c.x;
		o0.y = a.y + c.y;
		const float dr = a.x - c.x, di = a.y - c.y;
		o1.x = dr * wr - di * wi;
		o1.y = dr * wi + di * wr;
	};
	constexpr int DOUBLE_END = LOG2L & ~1;
#pragma unroll
	for (int st = 0; st < DOUBLE_END; st += 2) {
		if (active && b < Q) {
			const int s = 1 << st;
			const int ps = b & ~(s - 1);
			float2 y0, y1, y0b, y1b, z0, z1, z2, z3;
			bfly(lds_f2(x + 8 * b), lds_f2(x + 8 * (b + H)), lds_f2(tw + 8 * ps), y0, y1);
			bfly(lds_f2(x + 8 * (b + Q)), lds_f2(x + 8 * (b + Q + H)), lds_f2(tw + 8 * (ps + Q)), y0b, y1b);
			const float2 w2 = lds_f2(tw + 16 * ps);
			bfly(y0, y0b, w2, z0, z2);
			bfly(y1, y1b, w2, z1, z3);
			const unsigned o = y + 8 * (b + 3 * ps);
			sts_f2(o, z0);
			sts_f2(o + 8 * s, z1);
			sts_f2(o + 16 * s, z2);
			sts_f2(o + 24 * s, z3);
		}
		MAIN_SYNC();
		const unsigned sw = x;
		x = y;
		y = sw;
	}
#pragma unroll
	for (int st = DOUBLE_END; st < LOG2L; ++st) {
		if (active) {
			const int s = 1 << st;
			const int ps = b & ~(s - 1);
			float2 o0, o1;
			bfly(lds_f2(x + 8 * b), lds_f2(x + 8 * (b + H)), lds_f2(tw + 8 * ps), o0, o1);
			const unsigned oi = y + 8 * (b + ps);
			sts_f2(oi, o0);
			sts_f2(oi + 8 * s, o1);
		}
		MAIN_SYNC();
		const unsigned sw = x;
		x = y;
		y = sw;
	}
}
#if defined(AEC_FFT_INLINE) || defined(AEC_FFT_PLAIN)
#define CFFT_RUN(LOG2L, x, y, tw, sign, active) cfft_bfly<LOG2L>(x, y, tw, sign, active)
#else
#define CFFT_RUN(LOG2L, x, y, tw, sign, active) \
	cfft_core<LOG2L>(smem_addr(x), smem_addr(y), smem_addr(tw), (sign) < 0 ? 0x80000000u : 0u, (active) ? 1 : 0)
#endif
// number of buffer swaps of cfft_bfly: one per (double or single) pass
#ifdef AEC_FFT_PLAIN
template <int LOG2L> struct CfftSwaps { static constexpr int value = LOG2L; };
#else
template <int LOG2L> struct CfftSwaps { static constexpr int value = (LOG2L + 1) / 2; };
#endif
// single transform: threads < L/2 work. Returns the buffer holding the result.
template <int LOG2L> __device__ __forceinline__ float2 *cfft(float2 *x, float2 *y, const float2 *tw, int sign) {
	CFFT_RUN(LOG2L, x, y, tw, sign, threadIdx.x < (1 << (LOG2L - 1)));
	return (CfftSwaps<LOG2L>::value & 1) ? y : x;
}

// real forward FFT (spx_fft semantics: input scaled by 1/N): in[N] real (smem) -> spec[L] float2 (smem, bin0=(DC,Nyq))
// bufa/bufb: float2[L] scratch. in may alias nothing; spec may alias neither scratch.
template <int LOG2L>
__device__ void rfft(const float *in, float2 *spec, float2 *bufa, float2 *bufb, const AecParams &P, const float2 *tw,
                     const float2 *spl) {
	const int k = threadIdx.x, L = 1 << LOG2L;
	const float scale = (float)(1. / P.N);
	bufa[k] = make_float2(scale * in[2 * k], scale * in[2 * k + 1]);
	MAIN_SYNC();
	const float2 *Z = cfft<LOG2L>(bufa, bufb, tw, -1);
	float2 out;
	if (k == 0) {
		out.x = Z[0].x + Z[0].y;
		out.y = Z[0].x - Z[0].y;
	} else {
		const float zr = Z[k].x, zi = Z[k].y, yr = Z[L - k].x, yi = -Z[L - k].y;
		const float er = 0.5f * (zr + yr), ei = 0.5f * (zi + yi);
		const float dr = 0.5f * (zr - yr), di = 0.5f * (zi - yi);
		const float c = spl[k].x, s = spl[k].y;
		out.x = er + (c * di - s * dr);
		out.y = ei - (s * di + c * dr);
	}
	MAIN_SYNC(); // all reads of Z done before spec (which may be reused scratch by the caller later) is written
	spec[k] = out;
	MAIN_SYNC();
}

// real inverse FFT (spx_ifft, unscaled): spec[L] float2 (bin0=(DC,Nyq)) -> out[N] real (smem)
template <int LOG2L>
__device__ void irfft(const float2 *spec, float *out, float2 *bufa, float2 *bufb, const AecParams &P, const float2 *tw,
                      const float2 *spl) {
	const int k = threadIdx.x, L = 1 << LOG2L;
	float2 z;
	if (k == 0) {
		z.x = spec[0].x + spec[0].y;
		z.y = spec[0].x - spec[0].y;
	} else {
		const float xr = spec[k].x, xi = spec[k].y;
		const float yr = spec[L - k].x, yi = -spec[L - k].y;
		const float er = xr + yr, ei = xi + yi, dr = xr - yr, di = xi - yi;
		const float c = spl[k].x, s = spl[k].y;
		const float orr = dr * c - di * s, oi = dr * s + di * c;
		z.x = er - oi;
		z.y = ei + orr;
	}
	MAIN_SYNC();
	bufa[k] = z;
	MAIN_SYNC();
	const float2 *r = cfft<LOG2L>(bufa, bufb, tw, +1);
	const float2 v = r[k];
	MAIN_SYNC();
	out[2 * k] = v.x;
	out[2 * k + 1] = v.y;
	MAIN_SYNC();
}

// ---- paired transforms: two independent FFTs, one per half of the CTA, sharing every stage's barrier
template <int LOG2L>
__device__ __forceinline__ void cfft_pair(float2 *&xa, float2 *&ya, float2 *&xb, float2 *&yb, const float2 *tw, int sign) {
	const bool second = threadIdx.x >= (1 << (LOG2L - 1)); // lower half of the CTA: transform a, upper half: transform b
	CFFT_RUN(LOG2L, second ? xb : xa, second ? yb : ya, tw, sign, true);
	if (CfftSwaps<LOG2L>::value & 1) {
		float2 *sw = xa; xa = ya; ya = sw;
		sw = xb; xb = yb; yb = sw;
	}
}
// two real forward FFTs (same arithmetic as rfft<> per transform); scratch: a0,a1 for the first, b0,b1 for the second
template <int LOG2L>
__device__ void rfft_pair(const float *in_a, float2 *spec_a, const float *in_b, float2 *spec_b, float2 *a0, float2 *a1,
                          float2 *b0, float2 *b1, const AecParams &P, const float2 *tw, const float2 *spl) {
	const int k = threadIdx.x, L = 1 << LOG2L;
	const float scale = (float)(1. / P.N);
	a0[k] = make_float2(scale * in_a[2 * k], scale * in_a[2 * k + 1]);
	b0[k] = make_float2(scale * in_b[2 * k], scale * in_b[2 * k + 1]);
	MAIN_SYNC();
	float2 *xa = a0, *ya = a1, *xb = b0, *yb = b1;
	cfft_pair<LOG2L>(xa, ya, xb, yb, tw, -1);
	float2 oa, ob;
	if (k == 0) {
		oa.x = xa[0].x + xa[0].y; oa.y = xa[0].x - xa[0].y;
		ob.x = xb[0].x + xb[0].y; ob.y = xb[0].x - xb[0].y;
	} else {
		const float c = spl[k].x, sn = spl[k].y;
		{
			const float zr = xa[k].x, zi = xa[k].y, yr = xa[L - k].x, yi = -xa[L - k].y;
			const float er = 0.5f * (zr + yr), ei = 0.5f * (zi + yi), dr = 0.5f * (zr - yr), di = 0.5f * (zi - yi);
			oa.x = er + (c * di - sn * dr);
			oa.y = ei - (sn * di + c * dr);
		}
		{
			const float zr = xb[k].x, zi = xb[k].y, yr = xb[L - k].x, yi = -xb[L - k].y;
			const float er = 0.5f * (zr + yr), ei = 0.5f * (zi + yi), dr = 0.5f * (zr - yr), di = 0.5f * (zi - yi);
			ob.x = er + (c * di - sn * dr);
			ob.y = ei - (sn * di + c * dr);
		}
	}
	MAIN_SYNC();
	spec_a[k] = oa;
	spec_b[k] = ob;
	MAIN_SYNC();
}
// two real inverse FFTs
template <int LOG2L>
__device__ void irfft_pair(const float2 *spec_a, float *out_a, const float2 *spec_b, float *out_b, float2 *a0, float2 *a1,
                           float2 *b0, float2 *b1, const AecParams &P, const float2 *tw, const float2 *spl) {
	const int k = threadIdx.x, L = 1 << LOG2L;
	float2 za, zb;
	if (k == 0) {
		za.x = spec_a[0].x + spec_a[0].y; za.y = spec_a[0].x - spec_a[0].y;
		zb.x = spec_b[0].x + spec_b[0].y; zb.y = spec_b[0].x - spec_b[0].y;
	} else {
		const float c = spl[k].x, sn = spl[k].y;
		{
			const float xr = spec_a[k].x, xi = spec_a[k].y, yr = spec_a[L - k].x, yi = -spec_a[L - k].y;
			const float er = xr + yr, ei = xi + yi, dr = xr - yr, di = xi - yi;
			const float orr = dr * c - di * sn, oi = dr * sn + di * c;
			za.x = er - oi; za.y = ei + orr;
		}
		{
			const float xr = spec_b[k].x, xi = spec_b[k].y, yr = spec_b[L - k].x, yi = -spec_b[L - k].y;
			const float er = xr + yr, ei = xi + yi, dr = xr - yr, di = xi - yi;
			const float orr = dr * c - di * sn, oi = dr * sn + di * c;
			zb.x = er - oi; zb.y = ei + orr;
		}
	}
	MAIN_SYNC();
	a0[k] = za;
	b0[k] = zb;
	MAIN_SYNC();
	float2 *xa = a0, *ya = a1, *xb = b0, *yb = b1;
	cfft_pair<LOG2L>(xa, ya, xb, yb, tw, +1);
	const float2 va = xa[k], vb = xb[k];
	MAIN_SYNC();
	out_a[2 * k] = va.x; out_a[2 * k + 1] = va.y;
	out_b[2 * k] = vb.x; out_b[2 * k + 1] = vb.y;
	MAIN_SYNC();
}

__device__ __forceinline__ short word2int(float x) {
	return (short)(x < -32767.5f ? -32768 : (x > 32766.5f ? 32767 : (int)floor(.5 + (double)x)));
}

// double-precision library routines as ONE out-of-line copy each (code size: see cfft_core)
__device__ __noinline__ double exp_d(double x) {
	return exp(x);
}
__device__ __noinline__ double sqrt_d(double x) {
	return sqrt(x);
}
__device__ __noinline__ float hypergeom_gain(float xx) {
	const float table[21] = {0.82157f, 1.02017f, 1.20461f, 1.37534f, 1.53363f, 1.68092f, 1.81865f,
	                         1.94811f, 2.07038f, 2.18638f, 2.29688f, 2.40255f, 2.50391f, 2.60144f,
	                         2.69551f, 2.78647f, 2.87458f, 2.96015f, 3.04333f, 3.12431f, 3.20326f};
	float x = xx;
	float integer = (float)floor((double)(2 * x));
	int ind = (int)integer;
	if (ind < 0) return 1.f;
	if (ind > 19) return (float)(1 + .1296 / (double)x);
	float frac = 2 * x - integer;
	return (float)((double)((1 - frac) * table[ind] + frac * table[ind + 1]) / sqrt_d((double)(x + .0001f)));
}
__device__ __forceinline__ float qcurve(float x) {
	return 1.f / (1.f + .15f / x);
}

// ------------------------------------------------------------------------------------------------ the kernel
// dynamic shared memory map (floats): see carve-up at the top of the kernel body
// SW (48 kHz only): the CTA has a ninth warp, the SERIAL warp. Lane 0 of it runs the two sequential IIR filters of the
// frame — DC notch + pre-emphasis of the microphone, de-emphasis of the output — that a main thread used to run while
// the other 255 waited at a barrier (ncu, profiles/r2c_aec_kernel_deep_pipe.hotlines.txt: 18 % of all warp stall samples
// sat on those two barriers). The notch of frame k+1 now runs under the tail of frame k and the far-end transforms /
// block pass of frame k+1 (none of them needs the microphone), the de-emphasis under the E / Y transforms and the
// adaptation statistics. Same operations on the same operands in the same order: results are bit-identical.
template <int LOG2L, int CTAS = AEC_CTAS_PER_SM_256, bool SW = false, bool TMA = false>
__global__ void __launch_bounds__((1 << LOG2L) + (SW ? 32 : 0), ((256 * CTAS) >> LOG2L) > 32 ? 32 : ((256 * CTAS) >> LOG2L))
    aec_kernel(const short *__restrict__ mic, const short *__restrict__ ref, short *__restrict__ out, int nframes,
               int io_stride, float2 *__restrict__ gX, float2 *__restrict__ gW, float2 *__restrict__ gFG,
               float *__restrict__ gS, AecParams P, const int *__restrict__ counts, int in_frame0, int in_ring, int out_stride, int out_frame0,
               int out_ring, int skew_ns, int skew_ctas, int pass_generic) {
	extern __shared__ float sm[];
	constexpr int F = 1 << LOG2L, N = 2 * F, L = F;
	const int M = P.M;
#ifdef AEC_DIAG_NO_PASS // timing experiment only (results are meaningless): the frame without its block pass
	const int M_PASS = 0;
#else
	const int M_PASS = M;
#endif
	const int t = threadIdx.x;
	const int stream = blockIdx.x;
	// ragged batches: this stream's own frame count (a stream that staged fewer frames than the bank's maximum in this
	// tick must NOT be fed padding: a made-up frame would enter its far-end history and its adaptive filter)
	if (counts) nframes = min(nframes, counts[stream]);
	if (nframes <= 0) return;
	// Phase skew. All CTAs of the first wave start together and would walk through a frame in lock step: everybody in the
	// transforms and serial sections (no HBM traffic at all), then everybody in the block pass (HBM saturated) — the memory
	// system idles while the SMs compute and the SMs idle while it streams. Delaying the k-th CTA of every SM by k x skew
	// spreads the passes over the frame time; later waves inherit the spread (a CTA starts when a slot frees).
	if (skew_ns > 0 && (int)blockIdx.x < 4 * skew_ctas && (int)blockIdx.x >= skew_ctas) {
		if (threadIdx.x == 0) {
			const unsigned long long wait = (unsigned long long)(blockIdx.x / skew_ctas) * (unsigned long long)skew_ns;
			unsigned long long t0, t1;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
			do {
				__nanosleep(500);
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			} while (t1 - t0 < wait);
		}
		__syncthreads(); // every thread of the CTA, the serial warp included
	}
	// ---- shared memory carve-up
	float2 *tw = reinterpret_cast<float2 *>(sm);           // [L/2]
	float2 *spl = tw + L / 2;                              // [L+1] (+1 pad)
	float2 *specB = spl + L + 2;                           // [L] spectrum scratch B
	float2 *Eprev = specB + L;                             // [L] previous frame's error spectrum
	float2 *pipe = Eprev + L;                              // [AEC_STAGES][3][L] cp.async ring: X_{j+1}, FG_j, W_j
	// ring slots AEC_OWN_STAGES.. live on top of the seven arrays below, all dead between the constraint pre-pass and the
	// end of the block pass: 5 * 8F + 2 * 4(F + NB_BANDS + 1) bytes >= 2 slots of 24F bytes
	float2 *bufa = pipe + AEC_OWN_STAGES * 3 * L;          // [L]
	float2 *bufb = bufa + L;                               // [L]
	float2 *specA = bufb + L;                              // [L] spectrum scratch A
	float *ebuf = reinterpret_cast<float *>(specA + L);    // [N]
	float *ybuf = ebuf + N;                                // [N]
	constexpr int VLEN = (F + NB_BANDS + 1 + 3) & ~3;      // F + NB_BANDS + 1 entries, rounded so that what follows stays 16-byte aligned
	float *vec1 = ybuf + N;                                // [VLEN]
	float *vec2 = vec1 + VLEN;                             // [VLEN]
	static_assert(AEC_STAGES - AEC_OWN_STAGES >= 0 && AEC_STAGES - AEC_OWN_STAGES <= 2, "the aliased scratch holds at most two ring slots");
	float *vec3 = vec2 + VLEN;                             // [VLEN]
	float *vec4 = vec3 + VLEN;                             // [VLEN]
	float *vec5 = vec4 + VLEN;                             // [VLEN]
	float *xw = vec5 + VLEN;                               // [N] far-end window
	float *input_own = xw + N;                             // [F] (builds without the serial warp)
	float *tmpv = input_own + F;                               // [N] generic real scratch
	float *power_1 = tmpv + N;                             // [F+1]
	float *prop = power_1 + F + 1;                         // [M]
	float *wpart = prop + M;                               // [M][8] per-warp |W_j|^2 partials
	float *red = wpart + M * 8;                            // [32]
	float *sc = red + 32;                                  // [SC_COUNT]
	int *si = reinterpret_cast<int *>(sc + SC_COUNT);      // [IN_COUNT]
	// serial-warp hand-off (SW only): reset flag, the microphone frame as floats, two filtered input frames (frame parity)
	int *sw_reset = si + IN_COUNT;                                                                  // [4]
	float *micf = reinterpret_cast<float *>((reinterpret_cast<size_t>(sw_reset + 4) + 15) & ~(size_t)15); // [F]
	float *inq = micf + F;                                                                          // [2][F]
	unsigned long long *mbars = reinterpret_cast<unsigned long long *>(inq + 2 * F);                // [AEC_STAGES] full, [AEC_STAGES] empty (TMA)
	constexpr int NT_ALL = F + 32;
	static_assert(!TMA || SW, "the bulk-copy producer is the serial warp");
	const unsigned pipe_s = smem_addr(pipe), full_s = smem_addr(mbars), empty_s = full_s + 8 * AEC_STAGES;
	if (TMA) {
		if (t == 0) {
			for (int i = 0; i < AEC_STAGES; ++i) {
				mbar_init(full_s + 8 * i, 1);               // the producer's arrive.expect_tx
				mbar_init(empty_s + 8 * i, (F + 31) >> 5);  // lane 0 of every warp of per-bin threads
			}
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads(); // every thread of the CTA: producer and consumers see the initialised barriers
	}

	float2 *X = gX + (size_t)stream * P.x_stride;
	float2 *W = gW + (size_t)stream * P.w_stride;
	float2 *FG = gFG + (size_t)stream * P.w_stride;
	float *S = gS + (size_t)stream * P.lay.total;
	const AecLayout &ly = P.lay;
	const int nwarps = (F + 31) >> 5, lane = t & 31, warp = t >> 5;

	if (SW && t >= F) {
		// ================================================================================ the serial warp
		// Filter memories live in registers for the whole launch (read from and written back to the state page here:
		// the main threads leave these four scalars alone in SW builds).
		float m0 = S[ly.scal + SC_NOTCH0], m1 = S[ly.scal + SC_NOTCH1], memD = S[ly.scal + SC_MEMD], memE = S[ly.scal + SC_MEME];
		const float radius = P.notch_radius, pre = P.preemph;
		const float den2 = radius * radius + .7f * (1 - radius) * (1 - radius);
		// producer state (TMA builds): this stream's far-end ring head, whether the foreground array is stale (then W_j
		// doubles as FG_j and no FG row is fetched), and the phase parity of every slot's "empty" barrier
		int p_head = reinterpret_cast<const int *>(S + ly.ints)[IN_HEAD];
		int p_fg_pending = reinterpret_cast<const int *>(S + ly.ints)[IN_FG_PENDING];
		unsigned p_empty_par = (1u << AEC_STAGES) - 1; // a fresh barrier passes a wait for parity 1: every slot starts free
		for (int fr = 0; fr < nframes; ++fr) {
			const int fin = in_ring > 0 ? (in_frame0 + fr) % in_ring : fr;
			const short *in_mic = mic + (size_t)stream * io_stride + (size_t)fin * F;
#pragma unroll
			for (int k = 0; k < F / 32; ++k) micf[lane + 32 * k] = (float)in_mic[lane + 32 * k];
			__syncwarp();
			if (lane == 0) {
				// DC notch (filter_dc_notch16) then pre-emphasis; the recurrence m0 -> vout -> m0 is the critical path: add,
				// mul, add, fma (2*a is exact, so fma(2, a, m1) rounds exactly like m1 + 2*a)
				float *dst = inq + (fr & 1) * F;
#pragma unroll 1
				for (int i = 0; i < F; i += 4) {
					const float4 vin4 = *reinterpret_cast<const float4 *>(micf + i);
					const float vin[4] = {vin4.x, vin4.y, vin4.z, vin4.w};
					float r[4], o[4];
#pragma unroll
					for (int k = 0; k < 4; ++k) {
						const float vout = m0 + vin[k];
						r[k] = radius * vout;
						m0 = __fmaf_rn(2.f, r[k] - vin[k], m1);
						m1 = vin[k] - den2 * vout;
						o[k] = r[k] - pre * memD;
						memD = r[k];
					}
					*reinterpret_cast<float4 *>(dst + i) = make_float4(o[0], o[1], o[2], o[3]);
				}
			}
			__syncwarp();
			bar_arrive<BAR_INPUT, NT_ALL>();
			if (TMA) {
				// ---- the block pass's producer: rows X_{j+1}, FG_j, W_j of block j into ring slot j % AEC_STAGES
				p_head = (p_head + M) % (M + 1);
				int xslot = p_head + 1 > M ? 0 : p_head + 1;
				const float2 *srcW = W, *srcF = FG;
				const unsigned row_bytes = F * sizeof(float2);
				const unsigned tx = p_fg_pending ? 2 * row_bytes : 3 * row_bytes;
				for (int j0 = 0; j0 < M; j0 += AEC_STAGES) {
#pragma unroll
					for (int sidx = 0; sidx < AEC_STAGES; ++sidx) {
						const int j = j0 + sidx;
						if (j < M) {
							// the ring's memory is the per-bin threads' scratch outside the pass: the slots with storage of
							// their own are handed over at the top of the frame, the aliased ones when the pre-pass is over
							if (j == 0) bar_wait<BAR_RING_OWN, NT_ALL>();
							if (j == AEC_OWN_STAGES) bar_wait<BAR_RING_ALIAS, NT_ALL>();
							if (lane == 0) {
								mbar_wait(empty_s + 8 * sidx, (p_empty_par >> sidx) & 1);
								const unsigned dst = pipe_s + sidx * 3 * row_bytes, bar = full_s + 8 * sidx;
								mbar_expect_tx(bar, tx);
								bulk_g2s(dst, X + (size_t)xslot * F, row_bytes, bar);
								if (!p_fg_pending) bulk_g2s(dst + row_bytes, srcF, row_bytes, bar);
								bulk_g2s(dst + 2 * row_bytes, srcW, row_bytes, bar);
							}
							p_empty_par ^= 1u << sidx;
							xslot = xslot == M ? 0 : xslot + 1;
							srcW += F;
							srcF += F;
							__syncwarp();
						}
					}
				}
				if (M <= AEC_OWN_STAGES) bar_wait<BAR_RING_ALIAS, NT_ALL>(); // keep the hand-off counts in step
			}
			bar_wait<BAR_DEEMPH_IN, NT_ALL>(); // tmpv = this frame's error signal, sw_reset = its verdict
			const int reset = sw_reset[0];
			if (TMA) p_fg_pending = reset ? 0 : si[IN_FG_PENDING]; // as the next frame's pass will find it
			if (lane == 0) {
#pragma unroll 1
				for (int i = 0; i < F; i += 4) {
					float4 v = *reinterpret_cast<const float4 *>(tmpv + i);
					v.x = v.x + pre * memE;
					v.y = v.y + pre * v.x;
					v.z = v.z + pre * v.y;
					v.w = v.w + pre * v.z;
					memE = v.w;
					*reinterpret_cast<float4 *>(tmpv + i) = v;
				}
			}
			__syncwarp();
			bar_arrive<BAR_DEEMPH_OUT, NT_ALL>();
			if (reset) m0 = m1 = memD = memE = 0.f; // speex_echo_state_reset() of this frame, before the next frame's notch
		}
		if (lane == 0) {
			S[ly.scal + SC_NOTCH0] = m0;
			S[ly.scal + SC_NOTCH1] = m1;
			S[ly.scal + SC_MEMD] = memD;
			S[ly.scal + SC_MEME] = memE;
		}
		return;
	}

	// ---- load constants and per-stream small state
#pragma unroll 1
	for (int i = t; i < L / 2; i += F) tw[i] = P.tw[i];
#pragma unroll 1
	for (int i = t; i <= L; i += F) spl[i] = P.spl[i];
	Eprev[t] = reinterpret_cast<float2 *>(S + ly.E)[t];
	xw[F + t] = S[ly.xprev + t];
#pragma unroll 1
	for (int i = t; i <= F; i += F) power_1[i] = S[ly.power_1 + i];
#pragma unroll 1
	for (int i = t; i < M; i += F) prop[i] = S[ly.prop + i];
	if (t < SC_COUNT) sc[t] = S[ly.scal + t];
	if (t < IN_COUNT) si[t] = reinterpret_cast<int *>(S + ly.ints)[t];
	MAIN_SYNC();

	int head = si[IN_HEAD];
	unsigned full_par = 0; // phase parity of every ring slot's "full" barrier (TMA builds)
	for (int fr = 0; fr < nframes; ++fr) {
		// frame addressing: linear, or frame-aligned circular buffers (chain re-framing between 10 ms ticks and frames)
		const int fin = in_ring > 0 ? (in_frame0 + fr) % in_ring : fr;
		const int fout = out_ring > 0 ? (out_frame0 + fr) % out_ring : fr;
		const short *in_mic = mic + (size_t)stream * io_stride + (size_t)fin * F;
		const short *in_ref = ref + (size_t)stream * io_stride + (size_t)fin * F;
		short *o16 = out + (size_t)stream * out_stride + (size_t)fout * F;
		const int mic_i = in_mic[t];
		const int ref_i = in_ref[t];
		head = (head + M) % (M + 1); // head - 1 mod (M+1): newest block goes to slot `head`
		const float ss = .35f / (float)M, ss_1 = 1 - ss;

		// ---- frame constants + the block pass's cp.async prologue, issued before the serial input filters so that the
		// first blocks of X / FG / W are already in shared memory when the pass starts (they do not depend on this frame's
		// input: X_{j+1} are older ring slots, FG and W the previous frame's filters).
		// Deferred foreground refresh: when the previous frame decided "foreground := background" (IN_FG_PENDING), the
		// copy is not a separate pass; the W_j streamed in (still the previous frame's final value) IS the foreground
		// block: it is used as such and written to FG_j on the way (no FG read, no extra W read).
		const bool do_update = si[IN_SATURATED] == 0;
		const bool fg_pending = si[IN_FG_PENDING] != 0;
		const int cc = si[IN_CANCEL_COUNT] + 1; // st->cancel_count++ at the top of the frame
		const int constr_j = M > 1 ? cc % (M - 1) + 1 : 0;
		// |W_j|^2 is only consumed by mdf_adjust_prop once the filter counts as adapted (or is about to)
		const bool need_wnorm = si[IN_ADAPTED] || sc[SC_SUM_ADAPT] > (float)M - 1.f;
		const int xs0 = head + 1 > M ? 0 : head + 1;
#ifdef AEC_PASS_DIRECT
		const float2 *gX_pf = X + (size_t)xs0 * F + t; // X_{j+1} of the block being prefetched
		// ring wrap: a uniform count-down and a constant step back, instead of comparing against (and re-deriving) the
		// ring's end and base addresses at every block
		int x_left = M + 1 - xs0;
		const long x_ring = (long)(M + 1) * F;
#endif
		// W and FG are walked through ONE pointer each that moves once per unrolled group of AEC_STAGES blocks: the
		// prefetch of block j + AEC_STAGES - 1 and the stores of block j are that pointer plus compile-time offsets
		// (immediates in the LDGSTS / STG encodings) instead of four running 64-bit pointers bumped every block
		float2 *gW_grp = W + t, *gF_grp = FG + t;
		float2 *const pipe_t = pipe + t;
		// The far-end ring is walked with a 32-bit BYTE offset from this thread's column (add a row, back to 0 at the ring's
		// end): an add, a compare and a select per block instead of a 64-bit pointer stepped back by the ring's length
		const char *const xg = reinterpret_cast<const char *>(X + t);
		constexpr unsigned ROW_BYTES = F * (unsigned)sizeof(float2);
		unsigned xoff = (unsigned)xs0 * ROW_BYTES;
		const unsigned x_ring_bytes = (unsigned)(M + 1) * ROW_BYTES;
		auto prefetch = [&](int stage, int rel) { // rel: block index relative to the group the pointers stand at
			float2 *dst = pipe_t + stage * 3 * F;
			cp_async8(dst, xg + xoff);
			if (!fg_pending) cp_async8(dst + F, gF_grp + rel * F);
			cp_async8(dst + 2 * F, gW_grp + rel * F);
			xoff += ROW_BYTES;
			if (xoff == x_ring_bytes) xoff = 0;
		};
#ifndef AEC_PASS_DIRECT
		if (TMA) {
			// the producer may fill the ring's own slots from here on. Everything it will read (last frame's W / FG / X
			// stores) and overwrite (the slots, last frame's scratch) was accessed through the generic proxy: fence first
			fence_proxy_async_all();
			bar_arrive<BAR_RING_OWN, NT_ALL>();
		} else {
		// early prologue: the ring slots with storage of their own (the aliased ones are still scratch until the pre-pass ends)
#pragma unroll
		for (int pj = 0; pj < (AEC_OWN_STAGES < AEC_STAGES - 1 ? AEC_OWN_STAGES : AEC_STAGES - 1); ++pj) {
			if (pj < M_PASS) prefetch(pj, pj);
			cp_async_commit();
		}
		}
#endif

		// ---- DC notch (serial IIR, filter_dc_notch16) then pre-emphasis on the microphone: the serial warp's job in SW
		// builds (the filtered frame arrives in inq[frame parity], first needed after the block pass)
		float *const input = SW ? inq + (fr & 1) * F : input_own;
		if (!SW) {
		tmpv[t] = (float)mic_i;
		MAIN_SYNC();
		if (t == 0) {
			const float radius = P.notch_radius;
			const float den2 = radius * radius + .7f * (1 - radius) * (1 - radius);
			float m0 = sc[SC_NOTCH0], m1 = sc[SC_NOTCH1];
			// the recurrence m0 -> vout -> m0 is the critical path: keep it at add, mul, add, fma. 2*a is exact, so
			// fma(2, a, m1) rounds exactly like m1 + 2*a; samples move through registers four at a time
#pragma unroll 2
			for (int i = 0; i < F; i += 4) {
				const float4 vin4 = *reinterpret_cast<const float4 *>(tmpv + i);
				const float vin[4] = {vin4.x, vin4.y, vin4.z, vin4.w};
				float r[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const float vout = m0 + vin[k];
					r[k] = radius * vout;
					m0 = __fmaf_rn(2.f, r[k] - vin[k], m1);
					m1 = vin[k] - den2 * vout;
				}
				*reinterpret_cast<float4 *>(input + i) = make_float4(r[0], r[1], r[2], r[3]);
			}
			sc[SC_NOTCH0] = m0;
			sc[SC_NOTCH1] = m1;
		}
		MAIN_SYNC();
		{
			const float prev = t == 0 ? sc[SC_MEMD] : input[t - 1];
			const float v = input[t] - P.preemph * prev;
			const float lastv = input[F - 1];
			MAIN_SYNC();
			input[t] = v;
			if (t == 0) sc[SC_MEMD] = lastv;
		}
		}
		// ---- far-end window shift + pre-emphasis
		{
			const float old = xw[F + t];
			tmpv[t] = (float)ref_i;
			MAIN_SYNC();
			const float prev = t == 0 ? sc[SC_MEMX] : tmpv[t - 1];
			xw[t] = old;
			xw[F + t] = (float)ref_i - P.preemph * prev;
			MAIN_SYNC();
			if (t == 0) sc[SC_MEMX] = tmpv[F - 1];
		}
		// ---- X_0 = FFT(x) into the ring
		rfft<LOG2L>(xw, specA, bufa, bufb, P, tw, spl);
		X[(size_t)head * F + t] = specA[t];
		float Sxx = block_sum<LOG2L>(xw[F + t] * xw[F + t], red);

		// ---- adjust proportional adaptation rate (uses |W_j|^2 of the previous frame's final W)
		if (si[IN_ADAPTED]) {
			// mdf_adjust_prop: prop_j = sqrt(1 + |W_j|^2); += .1*max; normalise to .99. One warp does it: the maximum by
			// shuffles (exact in any order), the sum of the M terms in the library's sequential order by every lane at once
			// (M dependent adds; one thread walking all three loops alone cost 4.5 % of the kernel's warp time in the
			// adapted regime, everybody else at the barrier), the rest per element. (float)sqrt((double)x) is the correctly
			// rounded float square root (53 >= 2 * 24 + 2 bits: no double rounding), i.e. __fsqrt_rn.
#pragma unroll 1
			for (int i = t; i < M; i += F) prop[i] = __fsqrt_rn(1.f + S[ly.wnorm + i]);
			MAIN_SYNC();
			if (warp == 0) {
				float mx = 1.f;
#pragma unroll 1
				for (int i = lane; i < M; i += 32) mx = fmaxf(mx, prop[i]);
#pragma unroll
				for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll 1
				for (int i = lane; i < M; i += 32) prop[i] += .1f * mx;
				__syncwarp();
				float prop_sum = 1.f;
#pragma unroll 4
				for (int i = 0; i < M; ++i) prop_sum += prop[i];
				__syncwarp();
#pragma unroll 1
				for (int i = lane; i < M; i += 32) prop[i] = (.99f * prop[i]) / prop_sum;
			}
			MAIN_SYNC();
		}

		// ---- AUMDF constraint pre-pass. Block 0 and one rotating block get their weight update followed by
		// IFFT -> zero the second half -> FFT. Their gradient only needs X_{j+1}, E and prop_j, all known here, so both are
		// updated and constrained now as ONE paired transform; the block pass below then runs without any barrier.
		// Results stay in shared memory: block 0 in specB, block constr_j in tmpv (viewed as float2[L]).
		const float2 Ep = Eprev[t];
		const float p1 = power_1[t], p1n = power_1[F]; // p1n only meaningful for thread 0 (Nyquist)
		float2 xj = specA[t];                          // X_0 (just computed)
		float2 *cspec = reinterpret_cast<float2 *>(tmpv);
		{
			const int s0 = head + 1 > M ? head - M : head + 1; // ring slot of X_1
			int sj = head + constr_j + 1;                      // ring slot of X_{constr_j+1}
			if (sj > M) sj -= M + 1;
			float2 w0 = W[t], w1 = W[(size_t)constr_j * F + t];
			if (do_update) {
				const float2 xa = X[(size_t)s0 * F + t], xb = X[(size_t)sj * F + t];
				const float pj0 = prop[0], pj1 = prop[constr_j];
				if (t == 0) {
					w0.x += (pj0 * p1) * (xa.x * Ep.x);
					w0.y += (pj0 * p1n) * (xa.y * Ep.y);
					w1.x += (pj1 * p1) * (xb.x * Ep.x);
					w1.y += (pj1 * p1n) * (xb.y * Ep.y);
				} else {
					const float Wg0 = pj0 * p1, Wg1 = pj1 * p1;
					w0.x += Wg0 * (xa.x * Ep.x + xa.y * Ep.y);
					w0.y += Wg0 * (-xa.y * Ep.x + xa.x * Ep.y);
					w1.x += Wg1 * (xb.x * Ep.x + xb.y * Ep.y);
					w1.y += Wg1 * (-xb.y * Ep.x + xb.x * Ep.y);
				}
			}
			specB[t] = w0;
			cspec[t] = w1;
			MAIN_SYNC();
			float2 *sa = reinterpret_cast<float2 *>(ebuf), *sb = reinterpret_cast<float2 *>(ybuf);
			irfft_pair<LOG2L>(specB, reinterpret_cast<float *>(specB), cspec, tmpv, bufa, bufb, sa, sb, P, tw, spl);
			reinterpret_cast<float *>(specB)[F + t] = 0.f;
			tmpv[F + t] = 0.f;
			MAIN_SYNC();
			rfft_pair<LOG2L>(reinterpret_cast<float *>(specB), specB, tmpv, cspec, bufa, bufb, sa, sb, P, tw, spl);
		}
#ifndef AEC_PASS_DIRECT
		// late prologue: from here to the end of the pass bufa / bufb / specA / ebuf / ybuf / vec1 / vec2 are ring slots
		if (TMA) {
			fence_proxy_async_all();
			bar_arrive<BAR_RING_ALIAS, NT_ALL>();
		} else {
#pragma unroll
		for (int pj = AEC_OWN_STAGES; pj < AEC_STAGES - 1; ++pj) {
			if (pj < M_PASS) prefetch(pj, pj);
			cp_async_commit();
		}
		}
#endif

		// ---- the pass over the M blocks: foreground output, weight update, background output.
		// X_{j+1}, FG_j and W_j are streamed HBM -> shared memory with cp.async, AEC_STAGES-1 blocks ahead of their use;
		// every thread copies and later reads only its own bin, so the pipeline needs no block barrier, only
		// cp.async.wait_group. The stage loop is unrolled so that all shared-memory offsets are immediates.
		float2 yfg = make_float2(0.f, 0.f), ybg = make_float2(0.f, 0.f);
		{
			static_assert(AEC_STAGES % 3 == 0, "the grouped |W_j|^2 reduction below folds three blocks at a time");
			float nrm[3] = {0.f, 0.f, 0.f};
			// |W_j|^2 of blocks jb .. jb + 2 (nrm[0 .. 2]) summed over the warp: three warp sums for the price of one and a
			// bit: after the first two folds the three quantities live in disjoint lane groups (block jb: lanes 0-7, jb+1:
			// 16-23, jb+2: 8-15 and 24-31) and share the remaining folds. Same pairing order (xor 16, 8, 4, 2, 1) as warp_sum:
			// identical sums, 6 SHFL not 15.
			auto wnorm_fold = [&](const int jb) {
				const bool hi16 = lane & 16, hi8 = lane & 8;
				float a = hi16 ? nrm[1] : nrm[0];
				a += __shfl_xor_sync(0xffffffffu, hi16 ? nrm[0] : nrm[1], 16);
				float c2 = nrm[2] + __shfl_xor_sync(0xffffffffu, nrm[2], 16);
				float c = hi8 ? c2 : a;
				c += __shfl_xor_sync(0xffffffffu, hi8 ? a : c2, 8);
				c += __shfl_xor_sync(0xffffffffu, c, 4);
				c += __shfl_xor_sync(0xffffffffu, c, 2);
				c += __shfl_xor_sync(0xffffffffu, c, 1);
				const int jw = lane == 0 ? jb : (lane == 16 ? jb + 1 : jb + 2);
				if ((lane == 0 || lane == 16 || lane == 8) && jw < M) wpart[jw * 8 + warp] = c;
			};
#ifdef AEC_PASS_DIRECT
			// A/B build: X_{j+1}, FG_j, W_j go HBM -> REGISTERS (two blocks ahead, a ring of two register sets) instead of
			// HBM -> shared memory -> registers: no LDGSTS, no LDS in the pass (the shared-memory pipe is the busiest unit)
			constexpr int GROUP = 6;
			float2 rX[2], rF[2], rW[2];
			auto fetch = [&](int slot, int rel) {
				rX[slot] = *gX_pf;
				rW[slot] = gW_grp[rel * F];
				if (!fg_pending) rF[slot] = gF_grp[rel * F];
				gX_pf += F;
				if (--x_left == 0) gX_pf -= x_ring;
			};
			if (0 < M) fetch(0, 0);
			if (1 < M) fetch(1, 1);
#else
			constexpr int GROUP = AEC_STAGES;
#endif
			int j0_first = 0;
#if !defined(AEC_PASS_DIRECT) && !defined(AEC_NO_TIGHT) && AEC_STAGES == 3
			// ---- the tight form of the pass for the common frame: no foreground refresh pending, adaptation on, and only the
			// groups whose three blocks AND their prefetches (two blocks ahead) all exist — 15 of the 16 groups at 48 kHz; the
			// generic loop below finishes the rest (and runs every other kind of frame). Same operations on the same operands in
			// the same order as the generic loop (bit-identical: test_aec_pass_forms_are_bit_identical); what it drops are the
			// per-block tests for the end of the filter, the end of the prefetches, a pending refresh and adaptation being off.
			// Measured (B200, 4096 streams, steady state, one box, A/B interleaved): 0.7818 ms per launch before, 0.7752 with the
			// 32-bit ring offset alone, 0.7676 with this loop — although it executes a third fewer instructions per block (ncu
			// on the generic loop, profiles/r2z_aec_kernel.*: 95 per block and warp, 30 of them arithmetic): the pass is paced
			// by its loads, not by its instructions. Variants that cost MORE than they saved: a second copy of the loop without
			// the bin-0 code for warps 1-7 (0.8126: every change of the kernel's size or register allocation moves the
			// non-pass code by a few per cent at the 56-register cap), the whole pass as an out-of-line function with the
			// register file to itself (0.8054).
			if (!TMA && !pass_generic && !fg_pending && do_update && M_PASS >= 5) {
				char *wg = reinterpret_cast<char *>(gW_grp);
				const char *fgp = reinterpret_cast<const char *>(gF_grp);
				const int n_fast = (M - 5) / 3 + 1; // groups j0 = 0, 3, ... with j0 + 4 < M
				auto block = [&](auto sidx_c, const int j0) {
					constexpr int SIDX = decltype(sidx_c)::value;
					const int j = j0 + SIDX;
					{ // block j + 2 into the slot block j - 1 left
						float2 *dst = pipe_t + ((SIDX + 2) % 3) * 3 * F;
						cp_async8(dst, xg + xoff);
						cp_async8(dst + F, fgp + (SIDX + 2) * ROW_BYTES);
						cp_async8(dst + 2 * F, wg + (SIDX + 2) * ROW_BYTES);
					}
					xoff += ROW_BYTES;
					if (xoff == x_ring_bytes) xoff = 0;
					cp_async_commit();
					cp_async_wait<AEC_STAGES - 1>();
					const float2 *src = pipe_t + SIDX * 3 * F;
					const float2 xj1 = src[0];
					float2 w = src[2 * F];
					const float2 fg = src[F];
					if (t == 0) {
						yfg.x += xj.x * fg.x;
						yfg.y += xj.y * fg.y;
					} else {
						yfg.x += (xj.x * fg.x - xj.y * fg.y);
						yfg.y += (xj.y * fg.x + xj.x * fg.y);
					}
					if (j == 0 || j == constr_j) {
						w = j == 0 ? specB[t] : cspec[t];
					} else {
						const float pj = prop[j];
						if (t == 0) {
							w.x += (pj * p1) * (xj1.x * Ep.x);
							w.y += (pj * p1n) * (xj1.y * Ep.y);
						} else {
							const float Wg = pj * p1;
							w.x += Wg * (xj1.x * Ep.x + xj1.y * Ep.y);
							w.y += Wg * (-xj1.y * Ep.x + xj1.x * Ep.y);
						}
					}
					*reinterpret_cast<float2 *>(wg + SIDX * ROW_BYTES) = w;
					if (need_wnorm) nrm[SIDX] = w.x * w.x + w.y * w.y;
					if (t == 0) {
						ybg.x += xj.x * w.x;
						ybg.y += xj.y * w.y;
					} else {
						ybg.x += (xj.x * w.x - xj.y * w.y);
						ybg.y += (xj.y * w.x + xj.x * w.y);
					}
					xj = xj1;
				};
#pragma unroll 1
				for (int g = 0; g < n_fast; ++g) {
					const int j0 = 3 * g;
					block(std::integral_constant<int, 0>{}, j0);
					block(std::integral_constant<int, 1>{}, j0);
					block(std::integral_constant<int, 2>{}, j0);
					if (need_wnorm) wnorm_fold(j0);
					wg += 3 * ROW_BYTES;
					fgp += 3 * ROW_BYTES;
				}
				gW_grp = reinterpret_cast<float2 *>(wg);
				gF_grp += (size_t)n_fast * 3 * F;
				j0_first = 3 * n_fast;
			}
#endif
			for (int j0 = j0_first; j0 < M_PASS; j0 += GROUP) {
#pragma unroll
				for (int sidx = 0; sidx < GROUP; ++sidx) {
					const int j = j0 + sidx;
					if (j < M) {
#ifdef AEC_PASS_DIRECT
						const float2 xj1 = rX[sidx & 1];
						float2 w = rW[sidx & 1];
						float2 fg = fg_pending ? w : rF[sidx & 1];
						if (fg_pending) gF_grp[sidx * F] = w;
						if (j + 2 < M) fetch(sidx & 1, sidx + 2);
#else
						if (TMA) {
							mbar_wait(full_s + 8 * sidx, (full_par >> sidx) & 1);
							full_par ^= 1u << sidx;
						} else {
							// keep AEC_STAGES-1 blocks in flight: the slot freed by block j-1 receives block j+STAGES-1
							if (j + AEC_STAGES - 1 < M) prefetch((sidx + AEC_STAGES - 1) % AEC_STAGES, sidx + AEC_STAGES - 1);
							cp_async_commit();
							cp_async_wait<AEC_STAGES - 1>();
						}
						const float2 *src = pipe_t + sidx * 3 * F;
						const float2 xj1 = src[0];
						float2 w = src[2 * F];
						float2 fg;
						if (fg_pending) {
							fg = w;
							gF_grp[sidx * F] = w;
						} else {
							fg = src[F];
						}
#endif
						// foreground: Y += X_j * FG_j (spectral_mul_accum; bin 0 carries two real products)
						if (t == 0) {
							yfg.x += xj.x * fg.x;
							yfg.y += xj.y * fg.y;
						} else {
							yfg.x += (xj.x * fg.x - xj.y * fg.y);
							yfg.y += (xj.y * fg.x + xj.x * fg.y);
						}
						const bool constrained = j == 0 || j == constr_j;
						if (constrained) {
							w = j == 0 ? specB[t] : cspec[t]; // updated + constrained by the pre-pass
						} else if (do_update) {
							// gradient: W_j += prop_j * power_1 * conj(X_{j+1}) * E  (weighted_spectral_mul_conj)
							const float pj = prop[j];
							if (t == 0) {
								w.x += (pj * p1) * (xj1.x * Ep.x);
								w.y += (pj * p1n) * (xj1.y * Ep.y);
							} else {
								const float Wg = pj * p1;
								w.x += Wg * (xj1.x * Ep.x + xj1.y * Ep.y);
								w.y += Wg * (-xj1.y * Ep.x + xj1.x * Ep.y);
							}
						}
						if (do_update || constrained) gW_grp[sidx * F] = w;
						// |W_j|^2 partial for next frame's mdf_adjust_prop: reduced three blocks at a time after the group
						if (need_wnorm) nrm[sidx % 3] = w.x * w.x + w.y * w.y;
						// background: Y += X_j * W_j
						if (t == 0) {
							ybg.x += xj.x * w.x;
							ybg.y += xj.y * w.y;
						} else {
							ybg.x += (xj.x * w.x - xj.y * w.y);
							ybg.y += (xj.y * w.x + xj.x * w.y);
						}
						xj = xj1;
						if (TMA) {
							// the slot goes back to the producer once the whole warp has read its three rows
							__syncwarp();
							if (lane == 0) mbar_arrive(empty_s + 8 * sidx);
						}
					}
					if (sidx % 3 == 2 && need_wnorm && j0 + sidx - 2 < M) wnorm_fold(j0 + sidx - 2);
				}
				gW_grp += GROUP * F;
				gF_grp += GROUP * F;
			}
			cp_async_wait<0>();
		}
		MAIN_SYNC();
		if (need_wnorm && t < M) {
			float s = 0.f;
			for (int w8 = 0; w8 < nwarps; ++w8) s += wpart[t * 8 + w8];
			S[ly.wnorm + t] = s;
		}
		if (t == 0) {
			if (!do_update) si[IN_SATURATED]--;
			si[IN_FG_PENDING] = 0; // the foreground array is materialised again
		}

		// ---- per-stream state for the statistics and the preprocessor: HBM -> the idle cp.async ring, consumed much
		// later by the thread that copied it (no barrier needed, only wait_group). Slot k holds F floats at stg[k*F].
		float *stg = reinterpret_cast<float *>(pipe + 2 * L);
		enum { SG_POWER, SG_EH, SG_YH, SG_LASTY1, SG_ECHO, SG_INBUF, SG_S, SG_SMIN, SG_SLOTS };
		static_assert((2 + SG_SLOTS / 2) * 1 <= AEC_OWN_STAGES * 3, "pc0, pc1 and the staged state fit the ring's own slots");
		cp_async4(stg + SG_POWER * F + t, S + ly.power + t);
		cp_async4(stg + SG_EH * F + t, S + ly.Eh + t);
		cp_async4(stg + SG_YH * F + t, S + ly.Yh + t);
		cp_async4(stg + SG_LASTY1 * F + t, S + ly.last_y + F + t);
		cp_async4(stg + SG_ECHO * F + t, S + ly.echo_noise + t);
		cp_async4(stg + SG_INBUF * F + t, S + ly.inbuf + t);
		cp_async4(stg + SG_S * F + t, S + ly.S + t);
		cp_async4(stg + SG_SMIN * F + t, S + ly.Smin + t);
		cp_async_commit();

		// ---- foreground and background filter outputs (one paired inverse transform)
		float2 *pc0 = pipe, *pc1 = pipe + L; // the cp.async ring is idle outside the block pass: scratch for the pair
		specA[t] = yfg;
		specB[t] = ybg;
		MAIN_SYNC();
		irfft_pair<LOG2L>(specA, ebuf, specB, ybuf, bufa, bufb, pc0, pc1, P, tw, spl);
		if (SW) bar_wait<BAR_INPUT, NT_ALL>(); // the serial warp's filtered microphone frame
		float Sff, Dbf, See;
		{
			const float ef = input[t] - ebuf[t + F], dd = ebuf[t + F] - ybuf[t + F], eb = input[t] - ybuf[t + F];
			Sff = ef * ef;
			Dbf = dd * dd;
			See = eb * eb;
			ebuf[t] = eb;
			block_sum3<LOG2L>(Sff, Dbf, See, red);
			Dbf = 10 + Dbf;
		}

		// ---- two-path logic (every thread evaluates the same scalars)
		float Davg1 = .6f * sc[SC_DAVG1] + .4f * (Sff - See);
		float Davg2 = .85f * sc[SC_DAVG2] + .15f * (Sff - See);
		float Dvar1 = .36f * sc[SC_DVAR1] + .16f * Sff * Dbf;
		float Dvar2 = .7225f * sc[SC_DVAR2] + .0225f * Sff * Dbf;
		int update_foreground = 0, fg_refresh = 0;
		if ((Sff - See) * fabsf(Sff - See) > (Sff * Dbf)) update_foreground = 1;
		else if ((Davg1 * fabsf(Davg1)) > (.5f * Dvar1)) update_foreground = 1;
		else if ((Davg2 * fabsf(Davg2)) > (.25f * Dvar2)) update_foreground = 1;
		MAIN_SYNC(); // everyone has read sc[] before it is rewritten
		if (update_foreground) {
			Davg1 = Davg2 = 0;
			Dvar1 = Dvar2 = 0;
			fg_refresh = 1; // foreground := background, materialised lazily by the next frame's block pass
			ebuf[t + F] = P.window[t + F] * ebuf[t + F] + P.window[t] * ybuf[t + F];
		} else {
			int reset_background = 0;
			if ((-(Sff - See) * fabsf(Sff - See)) > (4.f * (Sff * Dbf))) reset_background = 1;
			if ((-Davg1 * fabsf(Davg1)) > (4.f * Dvar1)) reset_background = 1;
			if ((-Davg2 * fabsf(Davg2)) > (4.f * Dvar2)) reset_background = 1;
			if (reset_background) {
				for (int j = 0; j < M; ++j) {
					const float2 w = FG[(size_t)j * F + t];
					W[(size_t)j * F + t] = w;
					float n2 = warp_sum(w.x * w.x + w.y * w.y);
					if (lane == 0) wpart[j * 8 + warp] = n2;
				}
				MAIN_SYNC();
				if (t < M) {
					float s = 0.f;
					for (int w8 = 0; w8 < nwarps; ++w8) s += wpart[t * 8 + w8];
					S[ly.wnorm + t] = s;
				}
				ybuf[t + F] = ebuf[t + F];
				ebuf[t] = input[t] - ybuf[t + F];
				See = Sff;
				Davg1 = Davg2 = 0;
				Dvar1 = Dvar2 = 0;
			}
		}
		if (t == 0) {
			if (fg_refresh) si[IN_FG_PENDING] = 1;
			sc[SC_DAVG1] = Davg1;
			sc[SC_DAVG2] = Davg2;
			sc[SC_DVAR1] = Dvar1;
			sc[SC_DVAR2] = Dvar2;
		}
		MAIN_SYNC();

		// ---- output signal and saturation test. The error / echo-estimate signals and their energies are settled first
		// (every thread touches its own elements only), so that the frame's sanity verdict is known BEFORE the serial
		// de-emphasis is handed over: a reset also clears that filter's memory.
		const int sat = block_any<LOG2L>(mic_i <= -32000 || mic_i >= 32000);
		const float de_in = input[t] - ebuf[t + F];
		{
			const float ev = ebuf[t];
			ebuf[t + F] = ev;
			ebuf[t] = 0.f;
		}
		float Sey = ebuf[t + F] * ybuf[t + F], Syy = ybuf[t + F] * ybuf[t + F], Sdd = input[t] * input[t];
		block_sum3<LOG2L>(Sey, Syy, Sdd, red);
		// ---- sanity checks
		float lasty0, lasty1; // st->last_y halves as the preprocessor's residual-echo estimate will see them
		int screwed = si[IN_SCREWED];
		bool zero_out = false;
		if (!(Syy >= 0 && Sxx >= 0 && See >= 0) || !(Sff < N * 1e9 && Syy < N * 1e9 && Sxx < N * 1e9)) {
			screwed += 50;
			zero_out = true;
		} else if (Sff > Sdd + (float)(N * 10000)) {
			screwed++;
		} else {
			screwed = 0;
		}
		// ---- de-emphasis (serial IIR) of the output, in place in tmpv
		tmpv[t] = de_in;
		ybuf[t] = 0.f;
		if (t == 0 && sat && si[IN_SATURATED] == 0) si[IN_SATURATED] = 1;
		if (SW) {
			if (t == 0) sw_reset[0] = screwed >= 50;
			bar_arrive<BAR_DEEMPH_IN, NT_ALL>(); // the serial warp filters while the transforms and statistics below run
		} else {
			MAIN_SYNC();
			if (t == 0) {
				float memE = sc[SC_MEME];
				const float pre = P.preemph;
#pragma unroll 2
				for (int i = 0; i < F; i += 4) {
					float4 v = *reinterpret_cast<const float4 *>(tmpv + i);
					v.x = v.x + pre * memE;
					v.y = v.y + pre * v.x;
					v.z = v.z + pre * v.y;
					v.w = v.w + pre * v.z;
					memE = v.w;
					*reinterpret_cast<float4 *>(tmpv + i) = v;
				}
				sc[SC_MEME] = memE;
			}
		}
		MAIN_SYNC();
		// E (kept for the next frame's gradient) and Y in one paired transform
		rfft_pair<LOG2L>(ebuf, Eprev, ybuf, specB, bufa, bufb, pc0, pc1, P, tw, spl);
		// Rf -> vec1, Yf -> vec2, Xf -> vec3 (F+1 bins; bin F is the Nyquist term held by thread 0)
		{
			const float2 e = Eprev[t], y = specB[t], x0 = X[(size_t)head * F + t];
			if (t == 0) {
				vec1[0] = e.x * e.x; vec1[F] = e.y * e.y;
				vec2[0] = y.x * y.x; vec2[F] = y.y * y.y;
				vec3[0] = x0.x * x0.x; vec3[F] = x0.y * x0.y;
			} else {
				vec1[t] = e.x * e.x + e.y * e.y;
				vec2[t] = y.x * y.x + y.y * y.y;
				vec3[t] = x0.x * x0.x + x0.y * x0.y;
			}
		}
		MAIN_SYNC();
		int adapted_now = 0;
		if (screwed >= 50) {
			// speex_echo_state_reset(): filters, history and statistics back to their initial values
			for (int j = 0; j < M; ++j) {
				W[(size_t)j * F + t] = make_float2(0.f, 0.f);
				FG[(size_t)j * F + t] = make_float2(0.f, 0.f);
			}
			for (int j = 0; j <= M; ++j) X[(size_t)j * F + t] = make_float2(0.f, 0.f);
#pragma unroll 1
			for (int i = t; i <= F; i += F) {
				S[ly.power + i] = 0;
				power_1[i] = 1.f;
				S[ly.Eh + i] = 0;
				S[ly.Yh + i] = 0;
			}
			S[ly.last_y + t] = 0; // first half only, as the library does
			cp_async_wait<0>();
			lasty0 = 0.f;
			lasty1 = stg[SG_LASTY1 * F + t];
			Eprev[t] = make_float2(0.f, 0.f);
			xw[t] = 0;
			xw[F + t] = 0;
			if (t < M) S[ly.wnorm + t] = 0;
			MAIN_SYNC();
			if (t == 0) {
				si[IN_CANCEL_COUNT] = 0;
				si[IN_SCREWED] = 0;
				si[IN_SATURATED] = 0;
				si[IN_ADAPTED] = 0;
				si[IN_FG_PENDING] = 0;
				sc[SC_NOTCH0] = sc[SC_NOTCH1] = 0;
				sc[SC_MEMD] = sc[SC_MEME] = sc[SC_MEMX] = 0;
				sc[SC_SUM_ADAPT] = 0;
				sc[SC_PEY] = sc[SC_PYY] = 1.f;
				sc[SC_DAVG1] = sc[SC_DAVG2] = sc[SC_DVAR1] = sc[SC_DVAR2] = 0;
			}
			MAIN_SYNC();
			// the library returns before the adaptation statistics; the preprocessor still runs on `out`
		} else {
			if (t == 0) {
				si[IN_SCREWED] = screwed;
				si[IN_CANCEL_COUNT] = cc;
			}
			See = See > (float)(N * 100) ? See : (float)(N * 100);
			Sxx += Sxx; // the library accumulates the far-end energy a second time at this point
			// ---- smoothed far-end power, filtered spectra, leak estimate
			float pey_part = 0.f, pyy_part = 0.f;
			cp_async_wait<0>(); // the staged state (own column only)
#pragma unroll 1
			for (int j = t; j <= F; j += F) {
				const bool nyq = j == F; // thread 0's second trip: the Nyquist entries are not staged
				const float pw = ss_1 * (nyq ? S[ly.power + j] : stg[SG_POWER * F + j]) + 1 + ss * vec3[j];
				S[ly.power + j] = pw;
				vec4[j] = pw;
				const float Eh_old = nyq ? S[ly.Eh + j] : stg[SG_EH * F + j];
				const float Yh_old = nyq ? S[ly.Yh + j] : stg[SG_YH * F + j];
				const float Eh = vec1[j] - Eh_old, Yh = vec2[j] - Yh_old;
				pey_part += Eh * Yh;
				pyy_part += Yh * Yh;
				S[ly.Eh + j] = (1 - P.spec_average) * Eh_old + P.spec_average * vec1[j];
				S[ly.Yh + j] = (1 - P.spec_average) * Yh_old + P.spec_average * vec2[j];
			}
			float unused3 = 0.f;
			block_sum3<LOG2L>(pey_part, pyy_part, unused3, red);
			float Pey = 1.f + pey_part;
			float Pyy = 1.f + pyy_part;
			Pyy = __fsqrt_rn(Pyy); // == (float)sqrt((double)Pyy), see mdf_adjust_prop above
			Pey = Pey / Pyy;
			float tmp32 = P.beta0 * Syy;
			if (tmp32 > P.beta_max * See) tmp32 = P.beta_max * See;
			const float alpha = tmp32 / See, alpha_1 = 1.f - alpha;
			float sPey = alpha_1 * sc[SC_PEY] + alpha * Pey;
			float sPyy = alpha_1 * sc[SC_PYY] + alpha * Pyy;
			if (sPyy < 1.f) sPyy = 1.f;
			if (sPey < .005f * sPyy) sPey = .005f * sPyy;
			if (sPey > sPyy) sPey = sPyy;
			float leak = sPey / sPyy;
			if (leak > 16383) leak = 32767;
			float RER = (.0001f * Sxx + 3.f * leak * Syy) / See;
			if (RER < Sey * Sey / (1 + See * Syy)) RER = Sey * Sey / (1 + See * Syy);
			if (RER > .5f) RER = .5f;
			int adapted = si[IN_ADAPTED];
			float sum_adapt = sc[SC_SUM_ADAPT];
			if (!adapted && sum_adapt > (float)M && leak * Syy > .03f * Syy) adapted = 1;
			MAIN_SYNC();
			if (adapted) {
#pragma unroll 1
				for (int i = t; i <= F; i += F) {
					float r = leak * vec2[i];
					const float e = vec1[i] + 1;
					if (r > .5f * e) r = .5f * e;
					r = .7f * r + .3f * (RER * e);
					power_1[i] = r / (e * (vec4[i] + 10));
				}
			} else {
				float adapt_rate = 0;
				if (Sxx > (float)(N * 1000)) {
					float tq = .25f * Sxx;
					if (tq > .25f * See) tq = .25f * See;
					adapt_rate = tq / See;
				}
#pragma unroll 1
				for (int i = t; i <= F; i += F) power_1[i] = adapt_rate / (vec4[i] + 10);
				sum_adapt = sum_adapt + adapt_rate;
			}
			if (t == 0) {
				sc[SC_PEY] = sPey;
				sc[SC_PYY] = sPyy;
				sc[SC_LEAK] = leak;
				sc[SC_SUM_ADAPT] = sum_adapt;
				si[IN_ADAPTED] = adapted;
			}
			adapted_now = adapted;
		}
		// ---- the output frame (the de-emphasis is done by now), then last_y (residual echo input for the preprocessor)
		if (SW) bar_wait<BAR_DEEMPH_OUT, NT_ALL>();
		const int out_i = zero_out ? 0 : word2int(tmpv[t]);
		if (screwed < 50) {
			lasty0 = lasty1 = stg[SG_LASTY1 * F + t];
			S[ly.last_y + t] = lasty0;
			if (adapted_now) {
				lasty1 = (float)(mic_i - out_i);
				S[ly.last_y + F + t] = lasty1;
			}
		}
		MAIN_SYNC();

		// ========================================================================= speex_preprocess_run
		{
			const int Mb = NB_BANDS;
			// state vectors consumed late in the preprocessor: loaded into registers now, their latency hides behind the
			// paired transform below (the same thread reads and later rewrites each element: no hazard)
			const float r_stmp = S[ly.Stmp + t], r_noise = S[ly.noise + t], r_oldps = S[ly.old_ps + t], r_zeta = S[ly.zeta + t];
			const float r_outbuf = S[ly.outbuf + t];
			const float r_band_oldps = t < NB_BANDS ? S[ly.old_ps + F + t] : 0.f, r_band_zeta = t < NB_BANDS ? S[ly.zeta + F + t] : 0.f;
			int nb_adapt = si[IN_NB_ADAPT] + 1;
			if (nb_adapt > 20000) nb_adapt = 20000;
			int min_count = si[IN_MIN_COUNT] + 1;
			float beta = 1.0f / (float)nb_adapt;
			if (beta < .03f) beta = .03f;
			const float beta_1 = 1.f - beta;
			float *ps = vec1, *echo_noise = vec2, *noise = vec3, *prior = vec4, *gains = vec5;
			// filterbank weights of this bin, and (threads < 3*Mb) the bin range of the band this thread will sum
			const float fl = P.filter_left[t], fr = P.filter_right[t];
			const int bl = P.bank_left[t];
			const int fb_which = t / Mb, fb_b = t - fb_which * Mb;
			int bs0 = 0, bs1 = 0, bs2 = 0;
			if (t < 3 * Mb) {
				bs0 = fb_b > 0 ? P.band_start[fb_b - 1] : 0;
				bs1 = P.band_start[fb_b];
				bs2 = P.band_start[fb_b + 1];
			}
			// ---- speex_echo_get_residual's frame (window * last_y) and preprocess_analysis' frame (inbuf | out, windowed):
			// one paired forward transform -> specB (residual), specA (ft; bin 0 = (ft[0], ft[2N-1]))
			tmpv[t] = P.window[t] * lasty0;
			tmpv[t + F] = P.window[t + F] * lasty1;
			ebuf[t] = stg[SG_INBUF * F + t] * P.pwindow[t];
			ebuf[F + t] = (float)out_i * P.pwindow[F + t];
			S[ly.inbuf + t] = (float)out_i;
			MAIN_SYNC();
			rfft_pair<LOG2L>(tmpv, specB, ebuf, specA, bufa, bufb, pc0, pc1, P, tw, spl);
			// weighted copies for the three filterbank_compute_bank32() calls: the per-band sums below then only add
			float *eR = reinterpret_cast<float *>(pc0), *eL = eR + F, *pR = reinterpret_cast<float *>(pc1), *pL = pR + F;
			float *nR = reinterpret_cast<float *>(bufa), *nL = nR + F;
			const float leak_now = sc[SC_LEAK];
			const float leak2 = leak_now > .5f ? 1.f : 2 * leak_now;
			{
				// residual echo |Y|^2 * leak2, truncated to integers; NaN / absurd-value guard on the DC term
				const float2 y = specB[t], y0 = specB[0];
				float re = t == 0 ? y.x * y.x : y.x * y.x + y.y * y.y;
				re = (float)(int)(leak2 * re);
				const float r0 = (float)(int)(leak2 * (y0.x * y0.x));
				if (!(r0 >= 0 && r0 < F * 1e9f)) re = 0;
				const float a = .6f * stg[SG_ECHO * F + t];
				const float en = a > re ? a : re;
				echo_noise[t] = en;
				eR[t] = fr * en;
				eL[t] = fl * en;
				const float2 f = specA[t];
				const float pv = t == 0 ? f.x * f.x : f.x * f.x + f.y * f.y;
				ps[t] = pv;
				pR[t] = fr * pv;
				pL[t] = fl * pv;
			}
			MAIN_SYNC();
			// ---- update_noise_prob
			const int min_range = nb_adapt < 100 ? 15 : (nb_adapt < 1000 ? 50 : (nb_adapt < 10000 ? 150 : 300));
			{
				float Sv;
				const float Sold = stg[SG_S * F + t];
				if (t == 0) Sv = .8f * Sold + .2f * ps[0];
				else if (t == F - 1) Sv = .8f * Sold + .2f * ps[F - 1];
				else Sv = .8f * Sold + .05f * ps[t - 1] + .1f * ps[t] + .05f * ps[t + 1];
				S[ly.S + t] = Sv;
				float smin = stg[SG_SMIN * F + t], stmp = r_stmp;
				if (nb_adapt == 1) smin = stmp = 0;
				if (min_count > min_range) {
					smin = stmp < Sv ? stmp : Sv;
					stmp = Sv;
				} else {
					smin = smin < Sv ? smin : Sv;
					stmp = stmp < Sv ? stmp : Sv;
				}
				S[ly.Smin + t] = smin;
				S[ly.Stmp + t] = stmp;
				const int update_prob = (.4f * Sv > smin) ? 1 : 0;
				float nz = r_noise;
				if (!update_prob || ps[t] < nz) {
					const float v = beta_1 * nz + beta * ps[t];
					nz = v > 0 ? v : 0;
				}
				noise[t] = nz;
				nR[t] = fr * nz;
				nL[t] = fl * nz;
				S[ly.noise + t] = nz;
			}
			if (min_count > min_range) min_count = 0;
			MAIN_SYNC();
			// ---- the three Bark filterbanks at once (3 * Mb threads; filterbank_compute_bank32's accumulation order)
#pragma unroll 1
			for (int u = t; u < 3 * Mb; u += F) { // one trip unless F < 3 * Mb (8 kHz)
				const int which = u / Mb, b = u - which * Mb;
				int c0 = bs0, c1 = bs1, c2 = bs2;
				if (u != t) {
					c0 = b > 0 ? P.band_start[b - 1] : 0;
					c1 = P.band_start[b];
					c2 = P.band_start[b + 1];
				}
				const float *R = which == 0 ? eR : (which == 1 ? pR : nR);
				const float *Lw = which == 0 ? eL : (which == 1 ? pL : nL);
				float *dst = which == 0 ? echo_noise : (which == 1 ? ps : noise);
				float acc = 0.f;
				if (b > 0)
					for (int i = c0; i < c1; ++i) acc += R[i];
				for (int i = c1; i < c2; ++i) acc += Lw[i];
				dst[F + b] = acc;
			}
			MAIN_SYNC();
			// ---- SNRs over F + Mb entries (thread t: bin t; threads < Mb also band F + t)
			float post_me[2], old_ps_me[2];
			for (int r = 0; r < 2; ++r) {
				const int i = r == 0 ? t : F + t;
				if (r == 1 && t >= Mb) break;
				float old_ps = r == 0 ? r_oldps : r_band_oldps;
				if (nb_adapt == 1) old_ps = ps[i];
				const float tot_noise = 1.f + noise[i] + echo_noise[i] + 0.f;
				float post = ps[i] / tot_noise - 1.f;
				if (post > 100.f) post = 100.f;
				const float rr = old_ps / (old_ps + tot_noise);
				const float gamma = .1f + .89f * (rr * rr);
				float pr = gamma * (post > 0 ? post : 0) + (1.f - gamma) * (old_ps / tot_noise);
				if (pr > 100.f) pr = 100.f;
				prior[i] = pr;
				post_me[r] = post;
				old_ps_me[r] = old_ps;
			}
			MAIN_SYNC();
			// ---- zeta
			{
				float z = r_zeta;
				if (t == 0) z = .7f * z + .3f * prior[0];
				else if (t < F - 1) z = .7f * z + .15f * prior[t] + .075f * prior[t - 1] + .075f * prior[t + 1];
				else z = .7f * z + .3f * prior[t];
				S[ly.zeta + t] = z;
				if (t < Mb) {
					float zb = .7f * r_band_zeta + .3f * prior[F + t];
					S[ly.zeta + F + t] = zb;
					gains[F + t] = zb; // stash band zeta for Zframe
				}
				MAIN_SYNC();
			}
			// every thread sums the Mb band zetas itself (same order): no broadcast barrier
			float Zframe = 0;
			for (int i = 0; i < Mb; ++i) Zframe = Zframe + gains[F + i];
			const float Pframe = .1f + .899f * qcurve(Zframe / (float)Mb);
			const float effective_echo_suppress =
			    (1.f - Pframe) * (float)P.echo_suppress + Pframe * (float)P.echo_suppress_active;
			// ---- Bark-band gains: gain -> tmpv[0..Mb), gain2 -> tmpv[Mb..2Mb), gain_floor -> tmpv[2Mb..3Mb)
			if (t < Mb) {
				const int i = F + t;
				const float noise_floor = P.noise_floor; // exp(.2302585 * noise_suppress): a constant of the bank (host)
				const float echo_floor = (float)exp_d((double)(.2302585f * effective_echo_suppress));
				const float gfl = (float)(sqrt_d((double)(noise_floor * noise[i] + echo_floor * echo_noise[i])) /
				                          sqrt_d((double)(1 + noise[i] + echo_noise[i])));
				const float prior_ratio = prior[i] / (prior[i] + 1.f);
				const float theta = prior_ratio * (1.f + post_me[1]);
				const float MM = hypergeom_gain(theta);
				float g = prior_ratio * MM;
				if (g > 1.f) g = 1.f;
				S[ly.old_ps + i] = .2f * old_ps_me[1] + (.8f * (g * g)) * ps[i];
				const float zb = gains[F + t];
				const float P1 = .199f + .8f * qcurve(zb);
				const float q = 1.f - Pframe * P1;
				const double ee = exp_d((double)(-theta));
				const float gain2b = (float)(1 / (1.f + (double)((q / (1.f - q)) * (1 + prior[i])) * ee));
				tmpv[t] = g;
				tmpv[Mb + t] = gain2b;
				tmpv[2 * Mb + t] = gfl;
			}
			MAIN_SYNC();
			// ---- linear-frequency gains
			{
				const float gain_b = tmpv[bl] * fl + tmpv[bl + 1] * fr;
				const float p = tmpv[Mb + bl] * fl + tmpv[Mb + bl + 1] * fr;
				const float gfl = tmpv[2 * Mb + bl] * fl + tmpv[2 * Mb + bl + 1] * fr;
				const float prior_ratio = prior[t] / (prior[t] + 1.f);
				const float theta = prior_ratio * (1.f + post_me[0]);
				const float MM = hypergeom_gain(theta);
				float g = prior_ratio * MM;
				if (g > 1.f) g = 1.f;
				if (.333f * g > gain_b) g = 3.f * gain_b;
				S[ly.old_ps + t] = .2f * old_ps_me[0] + (.8f * (g * g)) * ps[t];
				if (g < gfl) g = gfl;
				const float tq = p * __fsqrt_rn(g) + (1.f - p) * __fsqrt_rn(gfl); // correctly rounded, as the double detour is
				gains[t] = tq * tq;
				S[ly.echo_noise + t] = echo_noise[t];
				if (t < Mb) S[ly.echo_noise + F + t] = echo_noise[F + t]; // band part is recomputed each frame
				MAIN_SYNC();
			}
			// ---- apply gain, inverse FFT, synthesis window, overlap-add
			{
				float2 f = specA[t];
				if (t == 0) {
					f.x = gains[0] * f.x;
					f.y = gains[F - 1] * f.y;
				} else {
					f.x = gains[t] * f.x;
					f.y = gains[t] * f.y;
				}
				specA[t] = f;
				MAIN_SYNC();
			}
			irfft<LOG2L>(specA, tmpv, bufa, bufb, P, tw, spl);
			{
				const float a = tmpv[t] * P.pwindow[t];
				const float b = tmpv[F + t] * P.pwindow[F + t];
				o16[t] = word2int(r_outbuf + a);
				S[ly.outbuf + t] = b;
			}
			if (t == 0) {
				si[IN_NB_ADAPT] = nb_adapt;
				si[IN_MIN_COUNT] = min_count;
			}
			MAIN_SYNC();
		}
	}

	// ---- store per-stream small state kept in shared memory
	reinterpret_cast<float2 *>(S + ly.E)[t] = Eprev[t];
	S[ly.xprev + t] = xw[F + t];
#pragma unroll 1
	for (int i = t; i <= F; i += F) S[ly.power_1 + i] = power_1[i];
#pragma unroll 1
	for (int i = t; i < M; i += F) S[ly.prop + i] = prop[i];
	// (the two input filters' memories belong to the serial warp in SW builds)
	if (t < SC_COUNT && !(SW && (t == SC_NOTCH0 || t == SC_NOTCH1 || t == SC_MEMD || t == SC_MEME))) S[ly.scal + t] = sc[t];
	if (t < IN_COUNT) reinterpret_cast<int *>(S + ly.ints)[t] = t == IN_HEAD ? head : si[t];
}

// ------------------------------------------------------------------------------------------------ host side
struct msb200_aec {
	msb200_ctx *ctx;
	int n;
	int live; // streams [0, live) are processed (msb200_aec_set_live); == n by default
	AecParams P;
	msb200_devbuf counts; // per-stream frame counts of a ragged host call
	size_t smem_bytes;
	size_t smem_bytes_sw; // the build with the serial warp (48 kHz)
	int path;             // msb200_aec_set_path: 0 serial warp (default), 1 the 256-thread build, 3 serial warp at 3 CTAs per SM
	float2 *dX, *dW, *dFG;
	float *dS;
	void *d_tables;
	std::vector<float> init_page;
	msb200_devbuf mic, ref, out;
	int tail_ms, filter_length;
};

static float to_bark(float n) {
	return 13.1f * (float)atan(.00074f * n) + 2.24f * (float)atan(n * n * 1.85e-8f) + 1e-4f * n;
}

static size_t aec_smem_floats(int F, int M, bool sw = false) {
	const int N = 2 * F, L = F;
	size_t f2 = (size_t)(L / 2) + (L + 2) + 5 * (size_t)L + (size_t)AEC_OWN_STAGES * 3 * L; // tw, spl, bufa, bufb, specA, specB, Eprev, pipe (own slots)
	size_t fl = (size_t)N * 4 + F + (F + 1) + 5 * (size_t)((F + NB_BANDS + 1 + 3) & ~3) + M + (size_t)M * 8 + 32 + SC_COUNT + IN_COUNT;
	if (sw) fl += 4 + 4 + 3 * (size_t)F + 4 * AEC_STAGES; // reset flag, alignment slack, micf[F], inq[2][F], mbarriers
	return f2 * 2 + fl;
}

static int aec_write_init(msb200_aec *a, int first, int count) {
	// X, W, FG zero; small page = init_page
	const AecParams &P = a->P;
	cudaStream_t s = a->ctx->stream;
	MSB200_CUDA(cudaMemsetAsync(a->dX + (size_t)first * P.x_stride, 0, sizeof(float2) * P.x_stride * (size_t)count, s));
	MSB200_CUDA(cudaMemsetAsync(a->dW + (size_t)first * P.w_stride, 0, sizeof(float2) * P.w_stride * (size_t)count, s));
	MSB200_CUDA(cudaMemsetAsync(a->dFG + (size_t)first * P.w_stride, 0, sizeof(float2) * P.w_stride * (size_t)count, s));
	std::vector<float> pages((size_t)count * P.lay.total);
	for (int i = 0; i < count; ++i) memcpy(&pages[(size_t)i * P.lay.total], a->init_page.data(), sizeof(float) * (size_t)P.lay.total);
	MSB200_CUDA(cudaMemcpyAsync(a->dS + (size_t)first * P.lay.total, pages.data(), sizeof(float) * pages.size(), cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}

extern "C" {

int msb200_aec_frame_size_for_rate(int sample_rate, int framesize_at_8000) { // adjust_framesize, speexec.c:171-180
	int newsize = (framesize_at_8000 * sample_rate) / 8000, n = 1, next;
	while ((next = n << 1) <= newsize) n = next;
	return n;
}

int msb200_aec_create(msb200_ctx *ctx, int n_streams, int sample_rate, int tail_length_ms, int framesize_at_8000,
                      msb200_aec **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && sample_rate >= 8000 && tail_length_ms > 0 && framesize_at_8000 > 0);
	const int F = msb200_aec_frame_size_for_rate(sample_rate, framesize_at_8000);
	MSB200_CHECK_ARG(F >= 32 && F <= 256); // one thread per bin; 8 kHz..48 kHz with the reference's framesize 64
	const int filter_length = (tail_length_ms * sample_rate) / 1000; // speexec.c:194
	const int N = 2 * F, L = F, M = (filter_length + F - 1) / F;
	MSB200_CHECK_ARG(M >= 2 && M <= 512);
	msb200_aec *a = new msb200_aec();
	a->ctx = ctx;
	a->n = a->live = n_streams;
	a->tail_ms = tail_length_ms;
	a->path = getenv("MSB200_AEC_PATH") ? atoi(getenv("MSB200_AEC_PATH")) : 0;
	if (a->path != 1 && a->path != 3 && a->path != 4 && a->path != 5) a->path = 0;
	a->filter_length = filter_length;
	AecParams &P = a->P;
	P.F = F; P.N = N; P.M = M; P.L = L; P.rate = sample_rate;
	P.log2L = 0;
	while ((1 << P.log2L) < L) P.log2L++;
	P.spec_average = (float)F / (float)sample_rate;
	P.beta0 = (2.0f * (float)F) / (float)sample_rate;
	P.beta_max = (.5f * (float)F) / (float)sample_rate;
	P.notch_radius = sample_rate < 12000 ? .9f : (sample_rate < 24000 ? .982f : .992f);
	P.preemph = .9f;
	P.noise_suppress = -15; P.echo_suppress = -40; P.echo_suppress_active = -15;
	P.noise_floor = (float)exp((double)(.2302585f * (float)P.noise_suppress));
	// small-state layout
	AecLayout &ly = P.lay;
	int o = 0;
	auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
	ly.xprev = take(F); ly.E = take(N); ly.last_y = take(N); ly.power = take(F + 1); ly.power_1 = take(F + 1);
	ly.Eh = take(F + 1); ly.Yh = take(F + 1); ly.prop = take(M); ly.wnorm = take(M); ly.scal = take(SC_COUNT);
	ly.ints = take(IN_COUNT); ly.inbuf = take(F); ly.outbuf = take(F); ly.old_ps = take(F + NB_BANDS);
	ly.noise = take(F + NB_BANDS); ly.echo_noise = take(F + NB_BANDS); ly.zeta = take(F + NB_BANDS);
	ly.S = take(F); ly.Smin = take(F); ly.Stmp = take(F);
	ly.total = o;
	P.x_stride = (size_t)(M + 1) * F;
	P.w_stride = (size_t)M * F;
	// initial page (speex_echo_state_init + speex_preprocess_state_init)
	a->init_page.assign((size_t)ly.total, 0.f);
	for (int i = 0; i <= F; ++i) a->init_page[(size_t)ly.power_1 + i] = 1.f;
	std::vector<float> prop((size_t)M);
	{
		float sum, decay = (float)exp(-2.4 / M);
		prop[0] = .7f;
		sum = prop[0];
		for (int i = 1; i < M; i++) {
			prop[(size_t)i] = prop[(size_t)i - 1] * decay;
			sum = sum + prop[(size_t)i];
		}
		for (int i = M - 1; i >= 0; i--) prop[(size_t)i] = (.8f * prop[(size_t)i]) / sum;
	}
	for (int i = 0; i < M; ++i) a->init_page[(size_t)ly.prop + i] = prop[(size_t)i];
	a->init_page[(size_t)ly.scal + SC_PEY] = 1.f;
	a->init_page[(size_t)ly.scal + SC_PYY] = 1.f;
	for (int i = 0; i < F + NB_BANDS; ++i) {
		a->init_page[(size_t)ly.noise + i] = 1.f;
		a->init_page[(size_t)ly.old_ps + i] = 1.f;
	}
	// constant tables
	std::vector<float2> tw((size_t)L / 2), spl((size_t)L + 1);
	for (int j = 0; j < L / 2; ++j) tw[(size_t)j] = make_float2((float)cos(2.0 * M_PI * j / L), (float)sin(2.0 * M_PI * j / L));
	for (int k = 0; k <= L; ++k) spl[(size_t)k] = make_float2((float)cos(2.0 * M_PI * k / N), (float)sin(2.0 * M_PI * k / N));
	std::vector<float> window((size_t)N), pwindow((size_t)N);
	for (int i = 0; i < N; i++) window[(size_t)i] = (float)(.5 - .5 * cos(2 * M_PI * i / N));
	for (int i = 0; i < N; i++) { // conj_window(2*F)
		float tmp, x = (4.f * (float)i) / (float)N;
		int inv = 0;
		if (x < 1.f) {
		} else if (x < 2.f) {
			x = 2.f - x;
			inv = 1;
		} else if (x < 3.f) {
			x = x - 2.f;
			inv = 1;
		} else {
			x = 2.f - x + 2.f;
		}
		x = 1.271903f * x;
		tmp = .5f - .5f * (float)cos(.5 * M_PI * x);
		tmp = tmp * tmp;
		if (inv) tmp = 1.f - tmp;
		pwindow[(size_t)i] = (float)sqrt(tmp);
	}
	std::vector<int> bank_left((size_t)F, 0), band_start(NB_BANDS + 1, F);
	std::vector<float> fl((size_t)F, 0.f), frr((size_t)F, 0.f);
	{ // filterbank_new(NB_BANDS, rate, F, 1)
		float df = (float)sample_rate / (2.f * (float)F);
		float max_mel = to_bark((float)sample_rate / 2);
		float mel_interval = max_mel / (float)(NB_BANDS - 1);
		for (int i = 0; i < F; i++) {
			float curr_freq = (float)i * df;
			float mel = to_bark(curr_freq);
			float val;
			int id1;
			if (mel > max_mel) break;
			id1 = (int)(floor(mel / mel_interval));
			if (id1 > NB_BANDS - 2) {
				id1 = NB_BANDS - 2;
				val = 1.f;
			} else {
				val = (mel - (float)id1 * mel_interval) / mel_interval;
			}
			bank_left[(size_t)i] = id1;
			fl[(size_t)i] = 1.f - val;
			frr[(size_t)i] = val;
		}
		// band_start[b] = first bin with left band >= b (bins are monotonic in band index)
		for (int b = NB_BANDS; b >= 0; --b) {
			int first = F;
			for (int i = F - 1; i >= 0; --i)
				if (bank_left[(size_t)i] >= b) first = i;
			band_start[(size_t)b] = first;
		}
	}
	// one allocation for all tables
	size_t off = 0;
	auto place = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
	size_t o_tw = place(sizeof(float2) * tw.size()), o_spl = place(sizeof(float2) * spl.size());
	size_t o_win = place(sizeof(float) * N), o_pwin = place(sizeof(float) * N), o_bl = place(sizeof(int) * F);
	size_t o_fl = place(sizeof(float) * F), o_fr = place(sizeof(float) * F), o_bs = place(sizeof(int) * (NB_BANDS + 1));
	size_t o_prop = place(sizeof(float) * M);
	MSB200_CUDA(cudaMalloc(&a->d_tables, off));
	char *base = (char *)a->d_tables;
	MSB200_CUDA(cudaMemcpy(base + o_tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_spl, spl.data(), sizeof(float2) * spl.size(), cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_win, window.data(), sizeof(float) * N, cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_pwin, pwindow.data(), sizeof(float) * N, cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_bl, bank_left.data(), sizeof(int) * F, cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_fl, fl.data(), sizeof(float) * F, cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_fr, frr.data(), sizeof(float) * F, cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_bs, band_start.data(), sizeof(int) * (NB_BANDS + 1), cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemcpy(base + o_prop, prop.data(), sizeof(float) * M, cudaMemcpyHostToDevice));
	P.tw = (const float2 *)(base + o_tw); P.spl = (const float2 *)(base + o_spl);
	P.window = (const float *)(base + o_win); P.pwindow = (const float *)(base + o_pwin);
	P.bank_left = (const int *)(base + o_bl); P.filter_left = (const float *)(base + o_fl);
	P.filter_right = (const float *)(base + o_fr); P.band_start = (const int *)(base + o_bs);
	P.prop0 = (const float *)(base + o_prop);
	// state
	MSB200_CUDA(cudaMalloc(&a->dX, sizeof(float2) * P.x_stride * (size_t)n_streams));
	MSB200_CUDA(cudaMalloc(&a->dW, sizeof(float2) * P.w_stride * (size_t)n_streams));
	MSB200_CUDA(cudaMalloc(&a->dFG, sizeof(float2) * P.w_stride * (size_t)n_streams));
	MSB200_CUDA(cudaMalloc(&a->dS, sizeof(float) * (size_t)ly.total * (size_t)n_streams));
	int r = aec_write_init(a, 0, n_streams);
	if (r) return r;
	a->smem_bytes = aec_smem_floats(F, M) * sizeof(float);
	a->smem_bytes_sw = aec_smem_floats(F, M, true) * sizeof(float);
	MSB200_SMEM_OPTIN((aec_kernel<8, AEC_CTAS_PER_SM_256>), ctx, a->smem_bytes);
	MSB200_SMEM_OPTIN((aec_kernel<8, AEC_CTAS_PER_SM_256, true>), ctx, a->smem_bytes_sw);
	MSB200_SMEM_OPTIN((aec_kernel<8, 3, true>), ctx, a->smem_bytes_sw);
	MSB200_SMEM_OPTIN((aec_kernel<8, AEC_CTAS_PER_SM_256, true, true>), ctx, a->smem_bytes_sw);
	MSB200_SMEM_OPTIN((aec_kernel<8, 5>), ctx, a->smem_bytes);
	MSB200_SMEM_OPTIN((aec_kernel<8, 6>), ctx, a->smem_bytes);
	MSB200_SMEM_OPTIN(aec_kernel<7>, ctx, a->smem_bytes);
	MSB200_SMEM_OPTIN(aec_kernel<6>, ctx, a->smem_bytes);
	MSB200_SMEM_OPTIN(aec_kernel<5>, ctx, a->smem_bytes);
	*out = a;
	return MSB200_OK;
}

void msb200_aec_destroy(msb200_aec *a) {
	if (!a) return;
	cudaStreamSynchronize(a->ctx->stream);
	cudaFree(a->dX);
	cudaFree(a->dW);
	cudaFree(a->dFG);
	cudaFree(a->dS);
	cudaFree(a->d_tables);
	a->counts.release();
	a->mic.release();
	a->ref.release();
	a->out.release();
	delete a;
}

int msb200_aec_get_info(msb200_aec *a, msb200_aec_info *info) {
	MSB200_CHECK_ARG(a && info);
	info->frame_size = a->P.F;
	info->window_size = a->P.N;
	info->M = a->P.M;
	info->sample_rate = a->P.rate;
	info->filter_length = a->filter_length;
	info->state_bytes_per_stream = sizeof(float2) * (a->P.x_stride + 2 * a->P.w_stride) + sizeof(float) * (size_t)a->P.lay.total;
	return MSB200_OK;
}

int msb200_aec_reset(msb200_aec *a, int stream) {
	MSB200_CHECK_ARG(a && stream >= -1 && stream < a->n);
	return stream < 0 ? aec_write_init(a, 0, a->n) : aec_write_init(a, stream, 1);
}

int msb200_aec_process_dev(msb200_aec *a, const void *d_mic, const void *d_ref, void *d_out, int nframes, int stride) {
	MSB200_CHECK_ARG(a && nframes > 0 && stride >= nframes * a->P.F);
	return msb200i_aec_launch(a, d_mic, d_ref, stride, 0, 0, d_out, stride, 0, 0, nframes, nullptr);
}
int msb200_aec_process_counts_dev(msb200_aec *a, const void *d_mic, const void *d_ref, void *d_out, int nframes, int stride,
                                  const void *d_counts) {
	MSB200_CHECK_ARG(a && nframes > 0 && stride >= nframes * a->P.F);
	return msb200i_aec_launch(a, d_mic, d_ref, stride, 0, 0, d_out, stride, 0, 0, nframes, (const int *)d_counts);
}
} // extern "C"
int msb200i_aec_launch(msb200_aec *a, const void *d_mic, const void *d_ref, int in_stride, int in_frame0,
                       int in_ring_frames, void *d_out, int out_stride, int out_frame0, int out_ring_frames, int nframes,
                       const int *d_counts) {
	MSB200_CHECK_ARG(a && d_mic && d_ref && d_out && nframes > 0);
	// phase skew between the CTAs that share an SM (see the kernel): only worth it when the grid fills the chip
	static const int skew_env = getenv("MSB200_AEC_SKEW_US") ? atoi(getenv("MSB200_AEC_SKEW_US")) : AEC_DEFAULT_SKEW_US;
	// (three waves and more: on a grid of one or two waves the late starters of the first wave would end the launch late)
	const int skew_ns = a->live >= 12 * a->ctx->sm_count ? skew_env * 1000 : 0;
	// path 5 (cross-checks, A/B runs): the default build with the generic loop of the block pass for every frame
	static const int generic_env = getenv("MSB200_AEC_PASS_GENERIC") ? atoi(getenv("MSB200_AEC_PASS_GENERIC")) : 0;
	const int pass_generic = a->path == 5 || generic_env != 0;
#define AEC_ARGS                                                                                                       \
	(const short *)d_mic, (const short *)d_ref, (short *)d_out, nframes, in_stride, a->dX, a->dW, a->dFG, a->dS, a->P, \
	    d_counts, in_frame0, in_ring_frames, out_stride, out_frame0, out_ring_frames, skew_ns, a->ctx->sm_count, pass_generic
	if (a->live > 0) switch (a->P.F) {
		case 256: {
			// occupancy A/B (profiling): MSB200_AEC_CTAS=5 selects the 48-register build (5 CTAs per SM), 6 the 40-register one
			static const int ctas = getenv("MSB200_AEC_CTAS") ? atoi(getenv("MSB200_AEC_CTAS")) : AEC_CTAS_PER_SM_256;
			// serial warp (default): 288 threads per CTA; msb200_aec_set_path / MSB200_AEC_PATH select the 256-thread build
			// (1) or the serial-warp build at 3 CTAs per SM (3: 72 registers instead of 56) for A/B runs and cross-checks
			const int sw = (a->path == 0 || a->path == 5) ? 1 : (a->path == 1 ? 0 : a->path);
			if (sw == 4)
				MSB200_LAUNCH(a->ctx, (aec_kernel<8, AEC_CTAS_PER_SM_256, true, true>), a->live, 288, a->smem_bytes_sw, AEC_ARGS);
			else if (sw == 3) MSB200_LAUNCH(a->ctx, (aec_kernel<8, 3, true>), a->live, 288, a->smem_bytes_sw, AEC_ARGS);
			else if (sw && ctas == AEC_CTAS_PER_SM_256)
				MSB200_LAUNCH(a->ctx, (aec_kernel<8, AEC_CTAS_PER_SM_256, true>), a->live, 288, a->smem_bytes_sw, AEC_ARGS);
			else if (ctas == 5) MSB200_LAUNCH(a->ctx, (aec_kernel<8, 5>), a->live, 256, a->smem_bytes, AEC_ARGS);
			else if (ctas == 6) MSB200_LAUNCH(a->ctx, (aec_kernel<8, 6>), a->live, 256, a->smem_bytes, AEC_ARGS);
			else MSB200_LAUNCH(a->ctx, (aec_kernel<8, AEC_CTAS_PER_SM_256>), a->live, 256, a->smem_bytes, AEC_ARGS);
			break;
		}
		case 128: MSB200_LAUNCH(a->ctx, aec_kernel<7>, a->live, 128, a->smem_bytes, AEC_ARGS); break;
		case 64: MSB200_LAUNCH(a->ctx, aec_kernel<6>, a->live, 64, a->smem_bytes, AEC_ARGS); break;
		case 32: MSB200_LAUNCH(a->ctx, aec_kernel<5>, a->live, 32, a->smem_bytes, AEC_ARGS); break;
		default: msb200_set_error("unsupported AEC frame size %d", a->P.F); return MSB200_EINVAL;
	}
#undef AEC_ARGS
	return MSB200_OK;
}
extern "C" {

int msb200_aec_process(msb200_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out, int nframes) {
	MSB200_CHECK_ARG(a && mic && ref && out && nframes > 0);
	if (a->live == 0) return MSB200_OK;
	size_t bytes = (size_t)a->live * nframes * a->P.F * 2;
	int r;
	if ((r = a->mic.reserve(bytes)) || (r = a->ref.reserve(bytes)) || (r = a->out.reserve(bytes))) return r;
	cudaStream_t s = a->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(a->mic.p, mic, bytes, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaMemcpyAsync(a->ref.p, ref, bytes, cudaMemcpyHostToDevice, s));
	if ((r = msb200_aec_process_dev(a, a->mic.p, a->ref.p, a->out.p, nframes, nframes * a->P.F))) return r;
	MSB200_CUDA(cudaMemcpyAsync(out, a->out.p, bytes, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}

// host buffers laid out [stream][stride_samples] of which the first nframes * frame_size samples are this call's frames
// (a fixed-size staging arena whose frame count varies from tick to tick)
int msb200_aec_process_strided(msb200_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out, int nframes,
                               int stride_samples) {
	return msb200_aec_process_counts(a, mic, ref, out, nframes, stride_samples, nullptr);
}
int msb200_aec_process_counts(msb200_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out, int nframes,
                              int stride_samples, const int32_t *counts) {
	MSB200_CHECK_ARG(a && mic && ref && out && nframes > 0 && stride_samples >= nframes * a->P.F);
	// only the staged frames of every row cross PCIe: rows of nframes*F out of stride_samples
	const size_t pitch = (size_t)stride_samples * 2, row = (size_t)nframes * a->P.F * 2;
	size_t bytes = (size_t)a->live * pitch + 16;
	int r;
	if ((r = a->mic.reserve(bytes)) || (r = a->ref.reserve(bytes)) || (r = a->out.reserve(bytes))) return r;
	cudaStream_t s = a->ctx->stream;
	if (a->live > 0) {
		MSB200_CUDA(cudaMemcpy2DAsync(a->mic.p, pitch, mic, pitch, row, (size_t)a->live, cudaMemcpyHostToDevice, s));
		MSB200_CUDA(cudaMemcpy2DAsync(a->ref.p, pitch, ref, pitch, row, (size_t)a->live, cudaMemcpyHostToDevice, s));
	}
	if (counts && a->live > 0) {
		if ((r = a->counts.reserve(sizeof(int32_t) * (size_t)a->n))) return r;
		MSB200_CUDA(cudaMemcpyAsync(a->counts.p, counts, sizeof(int32_t) * (size_t)a->live, cudaMemcpyHostToDevice, s));
	}
	if ((r = msb200_aec_process_counts_dev(a, a->mic.p, a->ref.p, a->out.p, nframes, stride_samples, counts ? a->counts.p : nullptr)))
		return r;
	if (a->live > 0) MSB200_CUDA(cudaMemcpy2DAsync(out, pitch, a->out.p, pitch, row, (size_t)a->live, cudaMemcpyDeviceToHost, s));
	MSB200_HOST_DONE(a->ctx);
	return MSB200_OK;
}
int msb200_aec_set_path(msb200_aec *a, int path) {
	MSB200_CHECK_ARG(a && (path == 0 || path == 1 || path == 3 || path == 4 || path == 5));
	a->path = path;
	return MSB200_OK;
}
int msb200_aec_set_live(msb200_aec *a, int n_live) {
	MSB200_CHECK_ARG(a && n_live >= 0 && n_live <= a->n);
	a->live = n_live;
	return MSB200_OK;
}

// the foreground array lags by one frame when a refresh is pending (see the block pass): logically FG == W then
static int aec_fg_pending(msb200_aec *a, int stream, int *pending, int clear) {
	int *d = reinterpret_cast<int *>(a->dS + (size_t)stream * a->P.lay.total + a->P.lay.ints) + IN_FG_PENDING;
	cudaStream_t s = a->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(pending, d, sizeof(int), cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	if (clear && *pending) {
		int zero = 0;
		MSB200_CUDA(cudaMemcpyAsync(d, &zero, sizeof(int), cudaMemcpyHostToDevice, s));
		MSB200_CUDA(cudaStreamSynchronize(s));
	}
	return MSB200_OK;
}

// blob = {magic, F, M, rate} + W [M][F] float2 + FG [M][F] float2
size_t msb200_aec_state_blob_size(msb200_aec *a) {
	return a ? 16 + 2 * sizeof(float2) * a->P.w_stride : 0;
}
int msb200_aec_get_state_blob(msb200_aec *a, int stream, void *blob, size_t size) {
	MSB200_CHECK_ARG(a && blob && stream >= 0 && stream < a->n && size >= msb200_aec_state_blob_size(a));
	int32_t hdr[4] = {0x4D534145, a->P.F, a->P.M, a->P.rate};
	memcpy(blob, hdr, 16);
	size_t wb = sizeof(float2) * a->P.w_stride;
	cudaStream_t s = a->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync((char *)blob + 16, a->dW + (size_t)stream * a->P.w_stride, wb, cudaMemcpyDeviceToHost, s));
	int pending = 0, r = aec_fg_pending(a, stream, &pending, 0);
	if (r) return r;
	MSB200_CUDA(cudaMemcpyAsync((char *)blob + 16 + wb, (pending ? a->dW : a->dFG) + (size_t)stream * a->P.w_stride, wb, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}
int msb200_aec_set_state_blob(msb200_aec *a, int stream, const void *blob, size_t size) {
	MSB200_CHECK_ARG(a && blob && stream >= 0 && stream < a->n && size >= msb200_aec_state_blob_size(a));
	int32_t hdr[4];
	memcpy(hdr, blob, 16);
	if (hdr[0] != 0x4D534145 || hdr[1] != a->P.F || hdr[2] != a->P.M || hdr[3] != a->P.rate) {
		msb200_set_error("AEC state blob does not match this canceller (F/M/rate)");
		return MSB200_EINVAL;
	}
	size_t wb = sizeof(float2) * a->P.w_stride;
	cudaStream_t s = a->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(a->dW + (size_t)stream * a->P.w_stride, (const char *)blob + 16, wb, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaMemcpyAsync(a->dFG + (size_t)stream * a->P.w_stride, (const char *)blob + 16 + wb, wb, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	int pending = 0;
	return aec_fg_pending(a, stream, &pending, 1);
}

// probes return data in the ORACLE's layout (speex packed spectra, X ordered newest block first)
int msb200_aec_probe(msb200_aec *a, int stream, const char *what, float *out, int max_floats) {
	MSB200_CHECK_ARG(a && what && out && stream >= 0 && stream < a->n);
	const AecParams &P = a->P;
	const int F = P.F, N = P.N, M = P.M;
	cudaStream_t s = a->ctx->stream;
	auto unpack_blocks = [&](const float2 *dev, int nblocks, int ring_head) -> int {
		std::vector<float2> tmp((size_t)nblocks * F);
		if (cudaMemcpyAsync(tmp.data(), dev, sizeof(float2) * tmp.size(), cudaMemcpyDeviceToHost, s) != cudaSuccess) return MSB200_ECUDA;
		cudaStreamSynchronize(s);
		int n = 0;
		for (int j = 0; j < nblocks; ++j) {
			const int slot = ring_head < 0 ? j : (ring_head + j) % nblocks;
			const float2 *b = &tmp[(size_t)slot * F];
			for (int i = 0; i < N && n < max_floats; ++i, ++n) {
				float v;
				if (i == 0) v = b[0].x;
				else if (i == N - 1) v = b[0].y;
				else v = (i & 1) ? b[(i + 1) / 2].x : b[i / 2].y;
				out[n] = v;
			}
		}
		return n;
	};
	if (!strcmp(what, "W")) return unpack_blocks(a->dW + (size_t)stream * P.w_stride, M, -1);
	if (!strcmp(what, "foreground")) {
		int pending = 0, r = aec_fg_pending(a, stream, &pending, 0);
		if (r) return r;
		return unpack_blocks((pending ? a->dW : a->dFG) + (size_t)stream * P.w_stride, M, -1);
	}
	if (!strcmp(what, "X")) { // every stream keeps its own ring head in its state page
		int head = 0;
		MSB200_CUDA(cudaMemcpy(&head, reinterpret_cast<int *>(a->dS + (size_t)stream * P.lay.total + P.lay.ints) + IN_HEAD, sizeof(int),
		                       cudaMemcpyDeviceToHost));
		return unpack_blocks(a->dX + (size_t)stream * P.x_stride, M + 1, head);
	}
	std::vector<float> page((size_t)P.lay.total);
	MSB200_CUDA(cudaMemcpyAsync(page.data(), a->dS + (size_t)stream * P.lay.total, sizeof(float) * page.size(), cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	const AecLayout &ly = P.lay;
	int off = -1, n = 0;
	if (!strcmp(what, "power")) off = ly.power, n = F + 1;
	else if (!strcmp(what, "power_1")) off = ly.power_1, n = F + 1;
	else if (!strcmp(what, "prop")) off = ly.prop, n = M;
	else if (!strcmp(what, "last_y")) off = ly.last_y, n = N;
	else if (!strcmp(what, "noise")) off = ly.noise, n = F + NB_BANDS;
	else if (!strcmp(what, "echo_noise")) off = ly.echo_noise, n = F + NB_BANDS;
	else if (!strcmp(what, "old_ps")) off = ly.old_ps, n = F + NB_BANDS;
	else if (!strcmp(what, "scalars")) {
		const float *sc = &page[(size_t)ly.scal];
		const int *si = reinterpret_cast<const int *>(&page[(size_t)ly.ints]);
		float v[16] = {(float)si[IN_ADAPTED], sc[SC_SUM_ADAPT], sc[SC_LEAK], sc[SC_PEY], sc[SC_PYY], sc[SC_DAVG1],
		               sc[SC_DAVG2], sc[SC_DVAR1], sc[SC_DVAR2], (float)si[IN_SATURATED], (float)si[IN_SCREWED],
		               (float)si[IN_CANCEL_COUNT], sc[SC_MEME], sc[SC_MEMD], sc[SC_MEMX], (float)si[IN_NB_ADAPT]};
		n = max_floats < 16 ? max_floats : 16;
		memcpy(out, v, sizeof(float) * (size_t)n);
		return n;
	} else if (!strcmp(what, "E")) {
		const float2 *b = reinterpret_cast<const float2 *>(&page[(size_t)ly.E]);
		for (int i = 0; i < N && n < max_floats; ++i, ++n)
			out[n] = i == 0 ? b[0].x : (i == N - 1 ? b[0].y : ((i & 1) ? b[(i + 1) / 2].x : b[i / 2].y));
		return n;
	} else {
		msb200_set_error("unknown probe '%s'", what);
		return MSB200_EINVAL;
	}
	if (n > max_floats) n = max_floats;
	memcpy(out, &page[(size_t)off], sizeof(float) * (size_t)n);
	return n;
}

} // extern "C"
