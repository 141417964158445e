// plc.cu — MSGenericPLC for a bank of streams (include/msb200dsp.h "Generic PLC" section).
// Replaces, per stream and per tick, the body of the reference filter
//   /root/reference/src/audiofilters/msgenericplc.c:61-157 (generic_plc_process) and its signal model
//   /root/reference/src/audiofilters/genericplc.c:74-110 (generic_plc_fftbf: window, N-point real spectrum, every packed
//   bin moved to twice its index x 0.85, 2N-point inverse), :112-200 (generic_plc_generate_samples), :202-231 (history and
//   continuity buffers), :235-241 (cross-fade); transforms = ms_fft / ms_ifft (src/utils/dsptools.c:362-376) over the
//   mixed-radix float kiss_fft (src/utils/kiss_fft.c, kiss_fftr.c:204-296).
// The concealer clock that decides WHICH streams lost a packet is host control logic (src/base/mscommon.c:315-362) and
// stays with the caller: it hands a per-stream mode byte to the bank each tick.
//
// One CTA per stream; streams in mode 0 leave at once. All state (50 ms history, 2x5 ms continuity, 100 ms of generated
// signal, two 16-bit counters) is device-resident: 4N + 4T + 4 bytes per stream (8.3 KB at 16 kHz). The transforms run in
// shared memory: digit-reversal gather, then one pass per factor (2, 3, 4, 5) with one butterfly per thread per step.
// Each butterfly keeps the reference's order of float operations and this translation unit is built with -fmad=false,
// so the concealed samples are bit-identical to the reference's (oracle/oracle_plc.c, pinned against the unmodified
// filter). Twiddles, window and permutations are computed once on the host in double precision exactly as
// kiss_fft_alloc / kiss_fftr_alloc / generic_plc_create_context do, and shared by all streams (L2-resident).
#include "msb200_internal.h"

#include <cmath>

#define PLC_THREADS 256
#define PLC_MAX_FACTORS 8

struct PlcFft {
	int n, inverse, nf;
	int p[PLC_MAX_FACTORS], m[PLC_MAX_FACTORS], stride[PLC_MAX_FACTORS];
	const float2 *tw;          // [n]
	const float2 *super;       // [n] real-transform twiddles
	const unsigned short *perm; // [n] gather order
};

struct PlcParams {
	int rate, N, T, max_len, dec_start, fade_len;
	PlcFft fwd, inv;
	const float *window; // [N]
	short *hist;         // [streams][N]
	short *cont;         // [streams][2T]
	short *gen;          // [streams][2N]
	unsigned short *ctr; // [streams][2]: plc_index, plc_samples_used (16-bit, wrapping, as in genericplc.h:45-46)
};

__device__ __forceinline__ float2 c_mul(float2 a, float2 b) { // C_MUL, _kiss_fft_guts.h:109-113
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// one radix-p butterfly on F[0], F[m], .., F[(p-1)m]; j = position in the block, s = twiddle stride
__device__ __forceinline__ void plc_bfly(const PlcFft &k, float2 *F, int p, int m, int j, int s) {
	const float2 *tw = k.tw;
	if (p == 4) { // kf_bfly4, kiss_fft.c:85-149
		const float2 s0 = c_mul(F[m], __ldg(tw + j * s)), s1 = c_mul(F[2 * m], __ldg(tw + 2 * j * s)),
		             s2 = c_mul(F[3 * m], __ldg(tw + 3 * j * s));
		float2 f0 = F[0];
		const float2 s5 = c_sub(f0, s1);
		f0 = c_add(f0, s1);
		const float2 s3 = c_add(s0, s2), s4 = c_sub(s0, s2);
		F[2 * m] = c_sub(f0, s3);
		F[0] = c_add(f0, s3);
		if (k.inverse) {
			F[m] = make_float2(s5.x - s4.y, s5.y + s4.x);
			F[3 * m] = make_float2(s5.x + s4.y, s5.y - s4.x);
		} else {
			F[m] = make_float2(s5.x + s4.y, s5.y - s4.x);
			F[3 * m] = make_float2(s5.x - s4.y, s5.y + s4.x);
		}
	} else if (p == 2) { // kf_bfly2 :36-83
		const float2 t = c_mul(F[m], __ldg(tw + j * s)), f0 = F[0];
		F[m] = c_sub(f0, t);
		F[0] = c_add(f0, t);
	} else if (p == 3) { // kf_bfly3 :151-184
		const float2 epi3 = __ldg(tw + s * m);
		const float2 s1 = c_mul(F[m], __ldg(tw + j * s)), s2 = c_mul(F[2 * m], __ldg(tw + 2 * j * s));
		const float2 s3 = c_add(s1, s2);
		float2 s0 = c_sub(s1, s2);
		const float2 f0 = F[0];
		float2 f1 = make_float2(f0.x - s3.x * .5f, f0.y - s3.y * .5f);
		s0.x *= epi3.y;
		s0.y *= epi3.y;
		F[0] = c_add(f0, s3);
		F[2 * m] = make_float2(f1.x + s0.y, f1.y - s0.x);
		f1.x -= s0.y;
		f1.y += s0.x;
		F[m] = f1;
	} else { // p == 5: kf_bfly5 :186-243
		const float2 ya = __ldg(tw + s * m), yb = __ldg(tw + s * 2 * m);
		const float2 s0 = F[0];
		const float2 s1 = c_mul(F[m], __ldg(tw + j * s)), s2 = c_mul(F[2 * m], __ldg(tw + 2 * j * s));
		const float2 s3 = c_mul(F[3 * m], __ldg(tw + 3 * j * s)), s4 = c_mul(F[4 * m], __ldg(tw + 4 * j * s));
		const float2 s7 = c_add(s1, s4), s10 = c_sub(s1, s4), s8 = c_add(s2, s3), s9 = c_sub(s2, s3);
		F[0] = make_float2(s0.x + (s7.x + s8.x), s0.y + (s7.y + s8.y));
		float2 s5, s6, s11, s12;
		s5.x = s0.x + s7.x * ya.x + s8.x * yb.x;
		s5.y = s0.y + s7.y * ya.x + s8.y * yb.x;
		s6.x = s10.y * ya.y + s9.y * yb.y;
		s6.y = -(s10.x * ya.y) - s9.x * yb.y;
		F[m] = c_sub(s5, s6);
		F[4 * m] = c_add(s5, s6);
		s11.x = s0.x + s7.x * yb.x + s8.x * ya.x;
		s11.y = s0.y + s7.y * yb.x + s8.y * ya.x;
		s12.x = -(s10.y * yb.y) + s9.y * ya.y;
		s12.y = s10.x * yb.y - s9.x * ya.y;
		F[2 * m] = c_add(s11, s12);
		F[3 * m] = c_sub(s11, s12);
	}
}

// kiss_fft_stride (kiss_fft.c:474-484): gather `in` into `out` in digit-reversed order, then the levels from the
// innermost out (kf_work :300-403). in and out are distinct shared buffers; ends with a barrier.
__device__ void plc_cfft(const PlcFft &k, const float2 *in, float2 *out) {
	const int t = threadIdx.x;
	for (int o = t; o < k.n; o += PLC_THREADS) out[o] = in[__ldg(k.perm + o)];
	__syncthreads();
	for (int d = k.nf - 1; d >= 0; --d) {
		const int p = k.p[d], m = k.m[d], nb = k.stride[d];
		for (int b = t; b < nb * m; b += PLC_THREADS) {
			const int blk = b / m, j = b - blk * m;
			plc_bfly(k, out + blk * p * m + j, p, m, j, nb);
		}
		__syncthreads();
	}
}

// genericplc.c:74-110. src: N samples (global), dst: 2N samples (global; may alias src). bx, by: shared, N float2 each.
__device__ void plc_stretch(const PlcParams &P, const short *src, short *dst, float2 *bx, float2 *by) {
	const int t = threadIdx.x, N = P.N;
	float *fx = reinterpret_cast<float *>(bx), *fy = reinterpret_cast<float *>(by);
	for (int i = t; i < N; i += PLC_THREADS) fx[i] = (float)src[i] * __ldg(P.window + i);
	__syncthreads();
	// ---- ms_fft: kiss_fftr2 (kiss_fftr.c:204-259) then x 1/N
	plc_cfft(P.fwd, bx, by);
	{
		const int n = P.fwd.n; // N / 2
		const float scale = 1.f / (float)N;
		float *freq = fx + N; // second half of bx: the packed spectrum [r0, r1, i1, ..., r(N/2)]
		for (int k = t; k <= n / 2; k += PLC_THREADS) {
			if (k == 0) {
				const float2 t0 = by[0];
				freq[0] = (t0.x + t0.y) * scale;
				freq[2 * n - 1] = (t0.x - t0.y) * scale;
			} else {
				const float2 a = by[k], b = by[n - k], st = __ldg(P.fwd.super + k);
				const float f2r = a.x - b.x, f2i = a.y + b.y, f1r = a.x + b.x, f1i = a.y - b.y;
				const float twr = f2r * st.x - f2i * st.y, twi = f2i * st.x + f2r * st.y;
				freq[2 * k - 1] = (.5f * (f1r + twr)) * scale;
				freq[2 * k] = (.5f * (f1i + twi)) * scale;
				freq[2 * (n - k) - 1] = (.5f * (f1r - twr)) * scale; // k == n - k: these two win, as in the reference's loop
				freq[2 * (n - k)] = (.5f * (twi - f1i)) * scale;
			}
		}
	}
	__syncthreads();
	// ---- double the spectrum: packed entry i -> 2i (x ENERGY_ATTENUATION), odd entries zero (:92-96)
	for (int i = t; i < N; i += PLC_THREADS) {
		fy[2 * i] = fx[N + i] * 0.85f;
		fy[2 * i + 1] = 0.f;
	}
	__syncthreads();
	// ---- ms_ifft: kiss_fftri2 (kiss_fftr.c:261-296), 2N real points = N complex
	{
		const int n = P.inv.n; // N
		for (int k = t; k <= n / 2; k += PLC_THREADS) {
			if (k == 0) {
				bx[0] = make_float2(fy[0] + fy[2 * n - 1], fy[0] - fy[2 * n - 1]);
			} else {
				const float2 fk = make_float2(fy[2 * k - 1], fy[2 * k]);
				const float2 fnkc = make_float2(fy[2 * (n - k) - 1], -fy[2 * (n - k)]);
				const float2 fek = c_add(fk, fnkc), df = c_sub(fk, fnkc), fok = c_mul(df, __ldg(P.inv.super + k));
				bx[k] = c_add(fek, fok);
				float2 r = c_sub(fek, fok);
				r.y *= -1.f;
				bx[n - k] = r;
			}
		}
	}
	__syncthreads();
	plc_cfft(P.inv, bx, by);
	for (int i = t; i < 2 * N; i += PLC_THREADS) dst[i] = (short)(int)fy[i]; // (int16_t) of a float: truncate, wrap
	__syncthreads();
}

// genericplc.c:235-241
__device__ __forceinline__ short plc_fade(short from, short to, int i, int n) {
	const float progress = (float)i / (float)n;
	return (short)(int)((float)from * (1.f - progress) + (float)to * progress);
}

// genericplc.c:202-213: slide the history by n samples and append data (shared copy `blk` of the n new samples)
__device__ void plc_push_history(const PlcParams &P, short *hist, const short *blk, int n, short *tmp) {
	const int t = threadIdx.x, N = P.N;
	if (n < N) {
		for (int i = t; i < N; i += PLC_THREADS) tmp[i] = hist[i];
		__syncthreads();
		for (int i = t; i < N; i += PLC_THREADS) hist[i] = i < N - n ? tmp[i + n] : blk[i - (N - n)];
	} else {
		for (int i = t; i < N; i += PLC_THREADS) hist[i] = blk[n - N + i];
	}
	__syncthreads();
}

// mode byte per stream: 0 idle; bit0 = a block of n samples arrived (in place: delay by T through the continuity buffer,
// cross-fade out of a concealed stretch); bit1 = conceal n samples into io; bit2 (with bit0) = the filter was emitting
// comfort noise before this block (msgenericplc.c:77-88).
__global__ void __launch_bounds__(PLC_THREADS) plc_kernel(const __grid_constant__ PlcParams P, short *__restrict__ io, int n,
                                                          int stride, const uint8_t *__restrict__ mode) {
	extern __shared__ float2 smem[];
	const int stream = blockIdx.x, t = threadIdx.x;
	const unsigned md = mode[stream];
	if (md == 0) return;
	const int N = P.N, T = P.T;
	float2 *bx = smem, *by = smem + N;
	short *blk = reinterpret_cast<short *>(smem + 2 * N); // [n] the block being handled
	short *sc = blk + ((n + 7) & ~7);                     // [2T] continuity copy
	short *tmp = sc + ((2 * T + 7) & ~7);                 // [N] history shuffle space
	short *row = io + (size_t)stream * stride;
	short *hist = P.hist + (size_t)stream * N, *cont = P.cont + (size_t)stream * 2 * T, *gen = P.gen + (size_t)stream * 2 * N;
	unsigned short *ctr = P.ctr + 2 * stream;
	int index = ctr[0], used = ctr[1];
	__syncthreads(); // everybody has read the counters

	if (md & 1u) { // ---------------------------------------------------------------- received block
		for (int i = t; i < n; i += PLC_THREADS) blk[i] = row[i];
		for (int i = t; i < 2 * T; i += PLC_THREADS) sc[i] = cont[i];
		__syncthreads();
		plc_push_history(P, hist, blk, n, tmp);
		const int tb = T > n ? n : T; // genericplc.c:215-231
		const bool cng = (md & 4u) != 0, was_plc = used != 0;
		for (int i = t; i < n; i += PLC_THREADS) {
			short v = i < tb ? sc[i] : blk[i - tb];
			if (cng) {
				if (i < T) v = 0;
				else if (i < 2 * T) v = plc_fade((short)0, v, i - T, T);
			}
			if (was_plc && n >= 2 * T && i >= T && i < 2 * T) v = plc_fade(sc[i], v, i - T, T); // msgenericplc.c:96-101
			row[i] = v;
		}
		for (int i = t; i < T; i += PLC_THREADS) {
			short c = i < tb ? blk[n - tb + i] : sc[i];
			if (was_plc && n < 2 * T) c = plc_fade(sc[T + i], c, i, T); // :102-111
			cont[i] = c;
		}
		index = 0;
		used = 0;
	} else if (md & 2u) { // --------------------------------------------------------- conceal n samples
		if (used >= P.max_len) { // genericplc.c:116-122
			for (int i = t; i < n; i += PLC_THREADS) blk[i] = 0;
			for (int i = t; i < 2 * T; i += PLC_THREADS) cont[i] = 0;
			used = (used + n) & 0xFFFF;
		} else {
			if (used == 0) { // first missing block: generate from the history (:125-138)
				plc_stretch(P, hist, gen, bx, by);
				for (int i = t; i < T; i += PLC_THREADS) gen[i] = plc_fade(cont[i], gen[i], i, T);
				__syncthreads();
			}
			if (index + n + 2 * T > 2 * N) { // generated signal exhausted: stretch it again (:142-171)
				int ready = (2 * N - index - T) & 0xFFFF;
				if (ready > n) ready = n;
				for (int i = t; i < ready; i += PLC_THREADS) blk[i] = gen[index + i];
				for (int i = t; i < T; i += PLC_THREADS) sc[i] = gen[index + ready + i];
				__syncthreads();
				plc_stretch(P, gen, gen, bx, by);
				for (int i = t; i < T; i += PLC_THREADS) gen[i] = plc_fade(sc[i], gen[i], i, T);
				__syncthreads();
				for (int i = ready + t; i < n; i += PLC_THREADS) blk[i] = gen[i - ready];
				index = n - ready;
			} else {
				for (int i = t; i < n; i += PLC_THREADS) blk[i] = gen[index + i];
				index += n;
			}
			for (int i = t; i < 2 * T; i += PLC_THREADS) cont[i] = gen[index + i];
			__syncthreads();
			if (used + n > P.dec_start) { // fade to silence between 100 and 150 ms (:183-198); double arithmetic as there
				const int i0 = P.dec_start - used > 0 ? P.dec_start - used : 0;
				for (int i = i0 + t; i < n; i += PLC_THREADS) {
					if (used + i >= P.max_len) blk[i] = 0;
					else {
						const float q = (float)(P.dec_start - (used + i)) / (float)P.fade_len;
						blk[i] = (short)(int)((1.0 + (double)q) * (double)(float)blk[i]);
					}
				}
			}
			used = (used + n) & 0xFFFF;
		}
		__syncthreads();
		for (int i = t; i < n; i += PLC_THREADS) row[i] = blk[i];
		plc_push_history(P, hist, blk, n, tmp); // msgenericplc.c:149-150
	}
	if (t == 0) {
		ctr[0] = (unsigned short)index;
		ctr[1] = (unsigned short)used;
	}
}

// ------------------------------------------------------------------------------------------------------------ host side
struct msb200_plc {
	msb200_ctx *ctx = nullptr;
	int n = 0, live = 0, rate = 0, max_block = 0; // live: streams [0, live) are copied and run (msb200_plc_set_live)
	PlcParams P{};
	void *d_tables = nullptr, *d_state = nullptr;
	msb200_devbuf io, md;
	size_t smem = 0;
};

namespace {
struct HostFft {
	int n, inverse, nf;
	int p[16], m[16], stride[16];
	std::vector<float2> tw, super;
	std::vector<unsigned short> perm;
};
// kiss_fft_alloc (kiss_fft.c:438-472), kf_factor (:405-428), kiss_fftr_alloc (kiss_fftr.c:39-77); false if a factor > 5
bool host_fft_init(HostFft &k, int nfft_real, int inverse) {
	const double pi = 3.14159265358979323846264338327;
	const int n = nfft_real >> 1;
	k.n = n;
	k.inverse = inverse;
	k.tw.resize((size_t)n);
	k.super.resize((size_t)n);
	k.perm.resize((size_t)n);
	for (int i = 0; i < n; ++i) {
		double phase = (-2 * pi / n) * i;
		if (inverse) phase *= -1;
		k.tw[(size_t)i] = make_float2((float)cos(phase), (float)sin(phase));
		double ph2 = pi * (((double)i) / n + .5);
		if (!inverse) ph2 = -ph2;
		k.super[(size_t)i] = make_float2((float)cos(ph2), (float)sin(ph2));
	}
	int rem = n, p = 4, nf = 0, stride = 1;
	do {
		while (rem % p) {
			if (p == 4) p = 2;
			else if (p == 2) p = 3;
			else p += 2;
			if (p > 32000 || p * p > rem) p = rem;
		}
		rem /= p;
		if (p > 5 || nf >= PLC_MAX_FACTORS) return false;
		k.p[nf] = p;
		k.m[nf] = rem;
		k.stride[nf] = stride;
		stride *= p;
		++nf;
	} while (rem > 1);
	k.nf = nf;
	for (int o = 0; o < n; ++o) { // kf_shuffle (:276-298)
		int src = 0;
		for (int d = 0; d < nf; ++d) src += ((o / k.m[d]) % k.p[d]) * k.stride[d];
		k.perm[(size_t)o] = (unsigned short)src;
	}
	return true;
}
size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }
} // namespace

extern "C" {

int msb200_plc_create(msb200_ctx *ctx, int n_streams, int sample_rate, int max_block, msb200_plc **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && sample_rate >= 8000 && sample_rate <= 48000 && max_block > 0);
	const int N = ((sample_rate * 2 / 40) / 100) * 100; // genericplc.c:42-44 (PLC_BUFFER_LEN = 2 / 40, genericplc.h:31)
	const int T = sample_rate * 5 / 1000;               // TRANSITION_DELAY
	MSB200_CHECK_ARG(N >= 4 && (N & 3) == 0);
	if (max_block + 2 * T > 2 * N) {
		msb200_set_error("msb200_plc_create: blocks of %d samples do not fit the %d-sample concealment buffer", max_block, 2 * N);
		return MSB200_EINVAL;
	}
	HostFft hf, hi;
	if (!host_fft_init(hf, N, 0) || !host_fft_init(hi, 2 * N, 1)) {
		msb200_set_error("msb200_plc_create: %d Hz needs transform factors other than 2, 3, 4, 5 (use 8 / 16 / 32 / 48 kHz)",
		                 sample_rate);
		return MSB200_EINVAL;
	}
	MSB200_CUDA(cudaSetDevice(ctx->device));
	msb200_plc *p = new msb200_plc();
	p->ctx = ctx;
	p->n = n_streams;
	p->live = n_streams;
	p->rate = sample_rate;
	p->max_block = max_block;
	// tables: window | tw_f | super_f | tw_i | super_i | perm_f | perm_i
	std::vector<float> window((size_t)N);
	for (int i = 0; i < N; ++i) window[(size_t)i] = (float)(0.75 - 0.25 * cos(2 * 3.14159265 * i / N)); // genericplc.c:60-62
	const size_t o_win = 0, o_twf = align16(o_win + sizeof(float) * N), o_suf = align16(o_twf + 8 * (size_t)hf.n),
	             o_twi = align16(o_suf + 8 * (size_t)hf.n), o_sui = align16(o_twi + 8 * (size_t)hi.n),
	             o_pf = align16(o_sui + 8 * (size_t)hi.n), o_pi = align16(o_pf + 2 * (size_t)hf.n),
	             tab_bytes = align16(o_pi + 2 * (size_t)hi.n);
	std::vector<uint8_t> tab(tab_bytes, 0);
	memcpy(tab.data() + o_win, window.data(), sizeof(float) * (size_t)N);
	memcpy(tab.data() + o_twf, hf.tw.data(), 8 * (size_t)hf.n);
	memcpy(tab.data() + o_suf, hf.super.data(), 8 * (size_t)hf.n);
	memcpy(tab.data() + o_twi, hi.tw.data(), 8 * (size_t)hi.n);
	memcpy(tab.data() + o_sui, hi.super.data(), 8 * (size_t)hi.n);
	memcpy(tab.data() + o_pf, hf.perm.data(), 2 * (size_t)hf.n);
	memcpy(tab.data() + o_pi, hi.perm.data(), 2 * (size_t)hi.n);
	const size_t per_stream = 2 * ((size_t)N + 2 * (size_t)T + 2 * (size_t)N + 2), st_bytes = per_stream * (size_t)n_streams;
	if (cudaMalloc(&p->d_tables, tab_bytes) != cudaSuccess || cudaMalloc(&p->d_state, st_bytes) != cudaSuccess) {
		msb200_set_error("msb200_plc_create: cudaMalloc(%zu) failed", st_bytes);
		msb200_plc_destroy(p);
		return MSB200_ENOMEM;
	}
	cudaStream_t s = ctx->stream;
	if (cudaMemcpyAsync(p->d_tables, tab.data(), tab_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess ||
	    cudaMemsetAsync(p->d_state, 0, st_bytes, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
		msb200_set_error("msb200_plc_create: table upload failed");
		msb200_plc_destroy(p);
		return MSB200_ECUDA;
	}
	PlcParams &P = p->P;
	P.rate = sample_rate;
	P.N = N;
	P.T = T;
	P.max_len = 150 * sample_rate / 1000;   // MAX_PLC_LEN
	P.dec_start = 100 * sample_rate / 1000; // PLC_DECREASE_START
	P.fade_len = 50 * sample_rate / 1000;
	auto fill = [&](PlcFft &d, const HostFft &h, size_t o_tw, size_t o_su, size_t o_perm) {
		d.n = h.n;
		d.inverse = h.inverse;
		d.nf = h.nf;
		for (int i = 0; i < h.nf; ++i) {
			d.p[i] = h.p[i];
			d.m[i] = h.m[i];
			d.stride[i] = h.stride[i];
		}
		const uint8_t *base = static_cast<const uint8_t *>(p->d_tables);
		d.tw = reinterpret_cast<const float2 *>(base + o_tw);
		d.super = reinterpret_cast<const float2 *>(base + o_su);
		d.perm = reinterpret_cast<const unsigned short *>(base + o_perm);
	};
	fill(P.fwd, hf, o_twf, o_suf, o_pf);
	fill(P.inv, hi, o_twi, o_sui, o_pi);
	P.window = reinterpret_cast<const float *>(static_cast<const uint8_t *>(p->d_tables) + o_win);
	short *st = static_cast<short *>(p->d_state);
	P.hist = st;
	P.cont = P.hist + (size_t)n_streams * N;
	P.gen = P.cont + (size_t)n_streams * 2 * T;
	P.ctr = reinterpret_cast<unsigned short *>(P.gen + (size_t)n_streams * 2 * N);
	p->smem = 16 * (size_t)N + 2 * ((size_t)((max_block + 7) & ~7) + (size_t)((2 * T + 7) & ~7) + (size_t)N);
	if (msb200_smem_optin((const void *)plc_kernel, ctx->device, p->smem) != cudaSuccess) {
		msb200_set_error("msb200_plc_create: %zu bytes of shared memory per stream not available", p->smem);
		msb200_plc_destroy(p);
		return MSB200_ECUDA;
	}
	*out = p;
	return MSB200_OK;
}

void msb200_plc_destroy(msb200_plc *p) {
	if (!p) return;
	cudaSetDevice(p->ctx->device);
	if (p->d_tables) cudaFree(p->d_tables);
	if (p->d_state) cudaFree(p->d_state);
	p->io.release();
	p->md.release();
	delete p;
}

int msb200_plc_history_samples(const msb200_plc *p) {
	return p ? p->P.N : 0;
}

int msb200_plc_reset_stream(msb200_plc *p, int stream) {
	MSB200_CHECK_ARG(p && stream >= 0 && stream < p->n);
	const PlcParams &P = p->P;
	cudaStream_t s = p->ctx->stream;
	MSB200_CUDA(cudaMemsetAsync(P.hist + (size_t)stream * P.N, 0, 2 * (size_t)P.N, s));
	MSB200_CUDA(cudaMemsetAsync(P.cont + (size_t)stream * 2 * P.T, 0, 4 * (size_t)P.T, s));
	MSB200_CUDA(cudaMemsetAsync(P.gen + (size_t)stream * 2 * P.N, 0, 4 * (size_t)P.N, s));
	MSB200_CUDA(cudaMemsetAsync(P.ctr + 2 * (size_t)stream, 0, 4, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}

int msb200_plc_set_live(msb200_plc *p, int n_live) {
	MSB200_CHECK_ARG(p && n_live >= 0 && n_live <= p->n);
	p->live = n_live;
	return MSB200_OK;
}

int msb200_plc_process_dev(msb200_plc *p, void *d_io, int nsamples, int stride_samples, const void *d_mode) {
	MSB200_CHECK_ARG(p && d_io && d_mode && nsamples > 0 && nsamples <= p->max_block && stride_samples >= nsamples);
	if (p->live == 0) return MSB200_OK;
	MSB200_LAUNCH(p->ctx, plc_kernel, p->live, PLC_THREADS, p->smem, p->P, static_cast<short *>(d_io), nsamples, stride_samples,
	              static_cast<const uint8_t *>(d_mode));
	return MSB200_OK;
}

int msb200_plc_process_strided(msb200_plc *p, int16_t *io, int nsamples, int stride_samples, const uint8_t *mode) {
	MSB200_CHECK_ARG(p && io && mode && nsamples > 0 && nsamples <= p->max_block && stride_samples >= nsamples);
	if (p->live == 0) return MSB200_OK;
	const size_t row = (size_t)nsamples * 2, rows = (size_t)p->live;
	int r = p->io.reserve(row * rows);
	if (r || (r = p->md.reserve(rows))) return r;
	cudaStream_t s = p->ctx->stream;
	MSB200_CUDA(cudaMemcpy2DAsync(p->io.p, row, io, (size_t)stride_samples * 2, row, rows, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaMemcpyAsync(p->md.p, mode, rows, cudaMemcpyHostToDevice, s));
	if ((r = msb200_plc_process_dev(p, p->io.p, nsamples, nsamples, p->md.p))) return r;
	MSB200_CUDA(cudaMemcpy2DAsync(io, (size_t)stride_samples * 2, p->io.p, row, row, rows, cudaMemcpyDeviceToHost, s));
	MSB200_HOST_DONE(p->ctx);
	return MSB200_OK;
}

int msb200_plc_process(msb200_plc *p, int16_t *io, int nsamples, const uint8_t *mode) {
	return msb200_plc_process_strided(p, io, nsamples, nsamples, mode);
}

} // extern "C"
