// resample.cu — MSResample: batched polyphase sinc resampler (speexdsp quality 3 "VOIP" design).
//
// Replaces resample_process_ms2() /root/reference/src/audiofilters/msresample.c:122-179, i.e.
// speex_resampler_process_int() of the external speexdsp library (restated in oracle/oracle_resample.c — see the
// provenance note there). Filter design (Kaiser-windowed sinc table) runs once on the host at bank creation; the
// per-tick work is one kernel: CTA per (stream, channel), thread per output sample, 48-tap (or filt_len-tap)
// float32 dot product accumulated in the library's sequential order with separate multiply and add.
#include "msb200_internal.h"

#include <cmath>

// ----------------------------------------------------------------------------------------------- host: filter design
namespace {

const double kaiser8_table[36] = {
    0.99537781, 1.00000000, 0.99537781, 0.98162644, 0.95908712, 0.92831446, 0.89005583, 0.84522401, 0.79486424,
    0.74011713, 0.68217934, 0.62226347, 0.56155915, 0.50119680, 0.44221549, 0.38553619, 0.33194107, 0.28205962,
    0.23636152, 0.19515633, 0.15859932, 0.12670280, 0.09935205, 0.07632451, 0.05731132, 0.04193980, 0.02979584,
    0.02044510, 0.01345224, 0.00839739, 0.00481569, 0.00247437, 0.00112393, 0.00042834, 0.00011921, 0.00000000};
const int kaiser8_oversample = 32;

// window lookup with the library's cubic interpolation (float index arithmetic, double accumulation)
double window_func(float x) {
	float y = x * (float)kaiser8_oversample;
	int ind = (int)floor(y);
	float frac = y - (float)ind;
	double c3 = -0.1666666667 * frac + 0.1666666667 * (frac * frac * frac);
	double c2 = frac + 0.5 * (frac * frac) - 0.5 * (frac * frac * frac);
	double c0 = -0.3333333333 * frac + 0.5 * (frac * frac) - 0.1666666667 * (frac * frac * frac);
	double c1 = 1.f - c3 - c2 - c0;
	return c0 * kaiser8_table[ind] + c1 * kaiser8_table[ind + 1] + c2 * kaiser8_table[ind + 2] + c3 * kaiser8_table[ind + 3];
}
float sinc_point(float cutoff, float x, int N) {
	float xx = x * cutoff;
	if (fabs(x) < 1e-6) return cutoff;
	if (fabs(x) > .5 * N) return 0;
	return (float)(cutoff * sin(M_PI * xx) / (M_PI * xx) * window_func((float)fabs(2. * x / N)));
}
uint32_t gcd_u32(uint32_t a, uint32_t b) {
	while (b) {
		uint32_t t = a % b;
		a = b;
		b = t;
	}
	return a;
}

struct ResampleDesign {
	uint32_t num, den, filt_len, oversample;
	int int_advance, frac_advance, use_direct;
	std::vector<float> table;
};

// quality 3: base_length 48, oversample 8, downsample_bw 0.895, upsample_bw 0.917, Kaiser-8 window
void design(uint32_t in_rate, uint32_t out_rate, ResampleDesign &d) {
	uint32_t g = gcd_u32(in_rate, out_rate);
	d.num = in_rate / g;
	d.den = out_rate / g;
	d.int_advance = (int)(d.num / d.den);
	d.frac_advance = (int)(d.num % d.den);
	d.oversample = 8;
	d.filt_len = 48;
	float cutoff;
	if (d.num > d.den) {
		cutoff = 0.895f * (float)d.den / (float)d.num;
		d.filt_len = (uint32_t)(((uint64_t)d.filt_len * d.num) / d.den);
		d.filt_len = ((d.filt_len - 1) & (~0x7u)) + 8;
		if (2 * d.den < d.num) d.oversample >>= 1;
		if (4 * d.den < d.num) d.oversample >>= 1;
		if (8 * d.den < d.num) d.oversample >>= 1;
		if (16 * d.den < d.num) d.oversample >>= 1;
		if (d.oversample < 1) d.oversample = 1;
	} else {
		cutoff = 0.917f;
	}
	d.use_direct = d.filt_len * d.den <= d.filt_len * d.oversample + 8;
	if (d.use_direct) {
		d.table.assign((size_t)d.filt_len * d.den, 0.f);
		for (uint32_t i = 0; i < d.den; i++)
			for (int32_t j = 0; j < (int32_t)d.filt_len; j++)
				d.table[(size_t)i * d.filt_len + (size_t)j] =
				    sinc_point(cutoff, ((float)(j - (int32_t)d.filt_len / 2 + 1) - ((float)i) / (float)d.den), (int)d.filt_len);
	} else {
		d.table.assign((size_t)d.filt_len * d.oversample + 8, 0.f);
		for (int32_t i = -4; i < (int32_t)(d.oversample * d.filt_len + 4); i++)
			d.table[(size_t)(i + 4)] =
			    sinc_point(cutoff, ((float)i / (float)d.oversample - (float)(d.filt_len / 2)), (int)d.filt_len);
	}
}

} // namespace

// ----------------------------------------------------------------------------------------------- device
struct ResampleParams {
	int filt_len, den, oversample, int_advance, frac_advance, use_direct, table_len, nch;
};

__device__ __forceinline__ short word2int(float x) { // speexdsp arch.h WORD2INT (float build)
	return (short)(x < -32767.5f ? -32768 : (x > 32766.5f ? 32767 : (int)floorf(.5f + x)));
}

// grid = (n_streams * nch); hist: [stream][ch][filt_len-1] s16 (exact: the library keeps the same integers as float)
__global__ void __launch_bounds__(256)
    resample_kernel(const short *__restrict__ in, int in_frames, int in_stride, short *__restrict__ out, int out_frames,
                    int out_stride, short *__restrict__ hist, const float *__restrict__ table, ResampleParams p,
                    int last_sample0, int samp_frac0, int ring_off, int ring_cap) {
	extern __shared__ float rsm[];
	float *tab = rsm;               // [table_len]
	float *x = rsm + p.table_len;   // [filt_len-1 + in_frames]
	const int N = p.filt_len;
	const int stream = blockIdx.x / p.nch, ch = blockIdx.x % p.nch;
	const short *gin = in + ((size_t)stream * in_stride) * p.nch + ch;
	short *gout = out + ((size_t)stream * out_stride) * p.nch + ch;
	short *h = hist + ((size_t)stream * p.nch + ch) * (N - 1);
	for (int i = threadIdx.x; i < p.table_len; i += blockDim.x) tab[i] = table[i];
	for (int i = threadIdx.x; i < N - 1; i += blockDim.x) x[i] = (float)h[i];
	for (int i = threadIdx.x; i < in_frames; i += blockDim.x) x[N - 1 + i] = (float)gin[(size_t)i * p.nch];
	__syncthreads();
	for (int k = threadIdx.x; k < out_frames; k += blockDim.x) {
		// closed form of the library's (last_sample, samp_frac_num) recurrence after k outputs
		long t = (long)samp_frac0 + (long)k * p.frac_advance;
		int last = last_sample0 + k * p.int_advance + (int)(t / p.den);
		int frac = (int)(t % p.den);
		const float *iptr = x + last;
		float y;
		if (p.use_direct) { // resampler_basic_direct_single
			const float *sinct = tab + (size_t)frac * N;
			float sum = 0.f;
#pragma unroll 8
			for (int j = 0; j < N; ++j) sum = __fadd_rn(sum, __fmul_rn(sinct[j], iptr[j]));
			y = sum;
		} else { // resampler_basic_interpolate_single
			const int offset = frac * p.oversample / p.den;
			const float fr = __fdiv_rn((float)((frac * p.oversample) % p.den), (float)p.den);
			float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
			for (int j = 0; j < N; ++j) {
				const float c = iptr[j];
				const float *tp = tab + 4 + (j + 1) * p.oversample - offset;
				a0 = __fadd_rn(a0, __fmul_rn(c, tp[-2]));
				a1 = __fadd_rn(a1, __fmul_rn(c, tp[-1]));
				a2 = __fadd_rn(a2, __fmul_rn(c, tp[0]));
				a3 = __fadd_rn(a3, __fmul_rn(c, tp[1]));
			}
			// cubic_coef(): evaluated left to right in float, interp[2] through double (the literal 1. is a double)
			float f2 = __fmul_rn(fr, fr), f3 = __fmul_rn(f2, fr);
			float i0 = __fadd_rn(__fmul_rn(-0.16667f, fr), __fmul_rn(__fmul_rn(__fmul_rn(0.16667f, fr), fr), fr));
			float i1 = __fsub_rn(__fadd_rn(fr, __fmul_rn(__fmul_rn(0.5f, fr), fr)), __fmul_rn(__fmul_rn(__fmul_rn(0.5f, fr), fr), fr));
			float i3 = __fsub_rn(__fadd_rn(__fmul_rn(-0.33333f, fr), __fmul_rn(__fmul_rn(0.5f, fr), fr)),
			                     __fmul_rn(__fmul_rn(__fmul_rn(0.16667f, fr), fr), fr));
			float i2 = (float)(1. - (double)i0 - (double)i1 - (double)i3);
			(void)f2; (void)f3;
			y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(i0, a0), __fmul_rn(i1, a1)), __fmul_rn(i2, a2)), __fmul_rn(i3, a3));
		}
		// ring_cap > 0: the output row is a circular buffer of ring_cap frames, written from ring_off (chain re-framing)
		const int ko = ring_cap > 0 ? (ring_off + k) % ring_cap : k;
		gout[(size_t)ko * p.nch] = word2int(y);
	}
	__syncthreads();
	// mem[j] = mem[j + in_frames] for j < N-1
	for (int i = threadIdx.x; i < N - 1; i += blockDim.x) h[i] = (short)x[i + in_frames];
}

// Integer-ratio up-sampling (16 kHz -> 48 kHz of BASELINE cfg2, 8 -> 16/48 kHz, ...): num = 1, den = R, 48 taps. Every
// input position m yields the R outputs R*m .. R*m+R-1 from the SAME 48-sample window with the R rows of the sinc table.
// One thread per input position: the window is read once into registers (48 LDS instead of 96 LDS per OUTPUT in the
// general kernel, whose shared-memory traffic is what bounds it) and the coefficients are compile-time-indexed kernel
// parameters, i.e. constant-bank operands of the multiplies — no loads at all. The products are accumulated in the
// library's order (j = 0..47, separate multiply and add), R independent chains per thread: bit-exact with the general
// kernel and the oracle. Requires phase (0, 0) at the start of the call: true for every block of a stream fed whole
// ticks (the phase returns to (0, 0) after each block).
#define RS_UP_TAPS 48
template <int R> struct ResampleUpTable { float t[R * RS_UP_TAPS]; };
template <int R>
__global__ void __launch_bounds__(256)
    resample_up_kernel(const short *__restrict__ in, int in_frames, int in_stride, short *__restrict__ out, int out_stride,
                       short *__restrict__ hist, const __grid_constant__ ResampleUpTable<R> tab, int nch, int ring_off, int ring_cap,
                       const short *__restrict__ in_b, short *__restrict__ out_b, short *__restrict__ hist_b) {
	extern __shared__ float rsm[];
	constexpr int N = RS_UP_TAPS;
	float *x = rsm; // [N-1 + in_frames]
	// gridDim.y == 2: two banks of one geometry and phase in ONE launch (the far-end and microphone resamplers of a chain
	// tick: two grids of small CTAs back to back left the chip half empty twice)
	if (blockIdx.y) {
		in = in_b;
		out = out_b;
		hist = hist_b;
	}
	const int stream = blockIdx.x / nch, ch = blockIdx.x % nch;
	const short *gin = in + ((size_t)stream * in_stride) * nch + ch;
	short *gout = out + ((size_t)stream * out_stride) * nch + ch;
	short *h = hist + ((size_t)stream * nch + ch) * (N - 1);
	for (int i = threadIdx.x; i < N - 1; i += blockDim.x) x[i] = (float)h[i];
	for (int i = threadIdx.x; i < in_frames; i += blockDim.x) x[N - 1 + i] = (float)gin[(size_t)i * nch];
	__syncthreads();
	for (int m = threadIdx.x; m < in_frames; m += blockDim.x) {
		float acc[R];
#pragma unroll
		for (int f = 0; f < R; ++f) acc[f] = 0.f;
		const float *w = x + m;
#pragma unroll
		for (int j = 0; j < N; ++j) {
			const float v = w[j];
#pragma unroll
			for (int f = 0; f < R; ++f) acc[f] = __fadd_rn(acc[f], __fmul_rn(tab.t[f * N + j], v));
		}
#pragma unroll
		for (int f = 0; f < R; ++f) {
			int ko = R * m + f;
			if (ring_cap > 0) {
				ko += ring_off;
				if (ko >= ring_cap) ko -= ring_cap;
			}
			gout[(size_t)ko * nch] = word2int(acc[f]);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < N - 1; i += blockDim.x) h[i] = (short)x[i + in_frames];
}

struct msb200_resample {
	msb200_ctx *ctx;
	int n, in_rate, out_rate, nch, max_in;
	int live; // streams [0, live) are processed (msb200_resample_set_live); == n by default
	ResampleDesign d;
	ResampleParams p;
	float *d_table;
	short *d_hist;
	int last_sample, samp_frac; // bank-wide phase (all streams are fed in lockstep)
	msb200_devbuf in, out;
};

// how many outputs the library loop produces for `in_frames` inputs from phase (last, frac), capped at out_cap;
// also returns the phase after the call (speex_resampler_process_native bookkeeping)
static int resample_count(const ResampleDesign &d, int in_frames, int out_cap, int &last, int &frac) {
	int out = 0;
	int l = last, f = frac;
	// closed form would do; the loop is <= out_cap iterations of integer adds per CALL (not per stream)
	while (!(l >= in_frames || out >= out_cap)) {
		out++;
		l += d.int_advance;
		f += d.frac_advance;
		if (f >= (int)d.den) {
			f -= (int)d.den;
			l++;
		}
	}
	int consumed = in_frames;
	if (l < in_frames) consumed = l;
	last = l - consumed;
	frac = f;
	return out;
}

extern "C" {

int msb200_resample_create(msb200_ctx *ctx, int n_streams, int in_rate, int out_rate, int nchannels, int max_in_frames,
                           msb200_resample **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && in_rate > 0 && out_rate > 0 && in_rate != out_rate);
	MSB200_CHECK_ARG(nchannels >= 1 && nchannels <= 8 && max_in_frames > 0 && max_in_frames <= 16384);
	msb200_resample *r = new msb200_resample();
	r->ctx = ctx;
	r->n = r->live = n_streams;
	r->in_rate = in_rate;
	r->out_rate = out_rate;
	r->nch = nchannels;
	r->max_in = max_in_frames;
	design((uint32_t)in_rate, (uint32_t)out_rate, r->d);
	r->p.filt_len = (int)r->d.filt_len;
	r->p.den = (int)r->d.den;
	r->p.oversample = (int)r->d.oversample;
	r->p.int_advance = r->d.int_advance;
	r->p.frac_advance = r->d.frac_advance;
	r->p.use_direct = r->d.use_direct;
	r->p.table_len = (int)r->d.table.size();
	r->p.nch = nchannels;
	r->last_sample = 0;
	r->samp_frac = 0;
	size_t smem = sizeof(float) * ((size_t)r->p.table_len + r->d.filt_len - 1 + (size_t)max_in_frames);
	if (smem > 200 * 1024) {
		msb200_set_error("resampler %d->%d with %d frames/call needs %zu B of shared memory", in_rate, out_rate, max_in_frames, smem);
		delete r;
		return MSB200_EINVAL;
	}
	MSB200_SMEM_OPTIN(resample_kernel, ctx, smem);
	size_t hist = (size_t)n_streams * nchannels * (r->d.filt_len - 1);
	MSB200_CUDA(cudaMalloc(&r->d_table, sizeof(float) * r->d.table.size()));
	MSB200_CUDA(cudaMalloc(&r->d_hist, sizeof(short) * hist));
	MSB200_CUDA(cudaMemcpy(r->d_table, r->d.table.data(), sizeof(float) * r->d.table.size(), cudaMemcpyHostToDevice));
	MSB200_CUDA(cudaMemset(r->d_hist, 0, sizeof(short) * hist));
	*out = r;
	return MSB200_OK;
}
void msb200_resample_destroy(msb200_resample *r) {
	if (!r) return;
	cudaStreamSynchronize(r->ctx->stream);
	cudaFree(r->d_table);
	cudaFree(r->d_hist);
	r->in.release();
	r->out.release();
	delete r;
}
int msb200_resample_max_out(msb200_resample *r, int in_frames) { // msresample.c:151-152
	if (!r || in_frames < 0) return MSB200_EINVAL;
	return (int)(((uint64_t)in_frames * (uint32_t)r->out_rate) / (uint32_t)r->in_rate) + 1;
}
int msb200_resample_reset(msb200_resample *r) {
	MSB200_CHECK_ARG(r);
	r->last_sample = r->samp_frac = 0;
	MSB200_CUDA(cudaMemsetAsync(r->d_hist, 0, sizeof(short) * (size_t)r->n * r->nch * (r->d.filt_len - 1), r->ctx->stream));
	return MSB200_OK;
}
int msb200_resample_reset_stream(msb200_resample *r, int stream) {
	MSB200_CHECK_ARG(r && stream >= 0 && stream < r->n);
	// the phase is bank-wide (lockstep streams); a stream that (re)joins starts from an empty history like a new handle
	const size_t per = sizeof(short) * (size_t)r->nch * (r->d.filt_len - 1);
	MSB200_CUDA(cudaMemsetAsync((char *)r->d_hist + per * (size_t)stream, 0, per, r->ctx->stream));
	return MSB200_OK;
}
int msb200_resample_process_dev(msb200_resample *r, const void *d_in, int in_frames, int in_stride, void *d_out,
                                int out_stride, int *out_frames) {
	return msb200i_resample_launch(r, d_in, in_frames, in_stride, d_out, out_stride, 0, 0, out_frames);
}
} // extern "C"
int msb200i_resample_launch(msb200_resample *r, const void *d_in, int in_frames, int in_stride, void *d_out,
                            int out_stride, int ring_off, int ring_cap, int *out_frames) {
	MSB200_CHECK_ARG(r && d_in && d_out && in_frames > 0 && in_frames <= r->max_in && in_stride >= in_frames);
	int cap = msb200_resample_max_out(r, in_frames);
	MSB200_CHECK_ARG(ring_cap > 0 ? out_stride >= ring_cap : out_stride >= cap);
	int last0 = r->last_sample, frac0 = r->samp_frac;
	int n_out = resample_count(r->d, in_frames, cap, r->last_sample, r->samp_frac);
	if (out_frames) *out_frames = n_out;
	// integer-ratio up-sampling from phase (0, 0): one thread per input position, coefficients in the constant bank
	if (r->live > 0 && r->p.use_direct && r->p.int_advance == 0 && r->p.frac_advance == 1 && r->p.filt_len == RS_UP_TAPS && last0 == 0 &&
	    frac0 == 0 && n_out == r->p.den * in_frames && (ring_cap == 0 || (ring_off + n_out <= 2 * ring_cap && n_out <= ring_cap)) &&
	    (r->p.den == 2 || r->p.den == 3 || r->p.den == 6)) {
		const size_t sm = sizeof(float) * (size_t)(RS_UP_TAPS - 1 + in_frames);
		int blk = in_frames >= 256 ? 256 : ((in_frames + 31) & ~31);
#define RS_UP_LAUNCH(R)                                                                                                \
	do {                                                                                                               \
		ResampleUpTable<R> tb;                                                                                         \
		memcpy(tb.t, r->d.table.data(), sizeof(tb.t));                                                                 \
		MSB200_SMEM_OPTIN(resample_up_kernel<R>, r->ctx, sm);                                                          \
		MSB200_LAUNCH(r->ctx, resample_up_kernel<R>, r->live * r->nch, blk, sm, (const short *)d_in, in_frames, in_stride, (short *)d_out, \
		              out_stride, r->d_hist, tb, r->nch, ring_off, ring_cap, (const short *)nullptr, (short *)nullptr, (short *)nullptr); \
	} while (0)
		if (r->p.den == 2) RS_UP_LAUNCH(2);
		else if (r->p.den == 3) RS_UP_LAUNCH(3);
		else RS_UP_LAUNCH(6);
#undef RS_UP_LAUNCH
		return MSB200_OK;
	}
	size_t smem = sizeof(float) * ((size_t)r->p.table_len + r->d.filt_len - 1 + (size_t)in_frames);
	int block = n_out >= 256 ? 256 : ((n_out + 31) & ~31);
	if (block < 32) block = 32;
	if (r->live > 0) MSB200_LAUNCH(r->ctx, resample_kernel, r->live * r->nch, block, smem, (const short *)d_in, in_frames, in_stride,
	              (short *)d_out, n_out, out_stride, r->d_hist, r->d_table, r->p, last0, frac0, ring_off, ring_cap);
	return MSB200_OK;
}
// Two banks of the same design, channel count, live range and phase in one launch (integer-ratio up-sampling from phase
// (0, 0), what a chain tick is); anything else: two launches. Same kernel, same arithmetic.
int msb200i_resample_launch_pair(msb200_resample *a, msb200_resample *b, const void *d_in_a, const void *d_in_b, int in_frames,
                                 int in_stride, void *d_out_a, void *d_out_b, int out_stride, int ring_off, int ring_cap, int *out_frames) {
	MSB200_CHECK_ARG(a && b && d_in_a && d_in_b && d_out_a && d_out_b);
	const ResampleParams &p = a->p;
	const int n_out_expected = (int)p.den * in_frames;
	const bool same = a->ctx == b->ctx && a->in_rate == b->in_rate && a->out_rate == b->out_rate && a->nch == b->nch && a->live == b->live &&
	                  a->last_sample == 0 && a->samp_frac == 0 && b->last_sample == 0 && b->samp_frac == 0;
	const bool up = a->live > 0 && p.use_direct && p.int_advance == 0 && p.frac_advance == 1 && p.filt_len == RS_UP_TAPS &&
	                (p.den == 2 || p.den == 3 || p.den == 6) && in_frames > 0 && in_frames <= a->max_in && in_frames <= b->max_in &&
	                in_stride >= in_frames && (ring_cap > 0 ? (out_stride >= ring_cap && ring_off + n_out_expected <= 2 * ring_cap && n_out_expected <= ring_cap)
	                                                        : out_stride >= n_out_expected);
	if (!(same && up)) {
		int r = msb200i_resample_launch(a, d_in_a, in_frames, in_stride, d_out_a, out_stride, ring_off, ring_cap, out_frames);
		if (r) return r;
		return msb200i_resample_launch(b, d_in_b, in_frames, in_stride, d_out_b, out_stride, ring_off, ring_cap, out_frames);
	}
	// the phase bookkeeping of both banks, as msb200i_resample_launch does it
	const int cap = msb200_resample_max_out(a, in_frames);
	const int n_out = resample_count(a->d, in_frames, cap, a->last_sample, a->samp_frac);
	const int n_out_b = resample_count(b->d, in_frames, cap, b->last_sample, b->samp_frac);
	if (n_out != n_out_expected || n_out_b != n_out_expected) {
		msb200_set_error("resample pair: %d / %d outputs for %d expected", n_out, n_out_b, n_out_expected);
		return MSB200_ESTATE;
	}
	if (out_frames) *out_frames = n_out;
	const size_t sm = sizeof(float) * (size_t)(RS_UP_TAPS - 1 + in_frames);
	const int blk = in_frames >= 256 ? 256 : ((in_frames + 31) & ~31);
	const dim3 grid((unsigned)(a->live * a->nch), 2, 1);
#define RS_UP_PAIR(R)                                                                                                  \
	do {                                                                                                               \
		ResampleUpTable<R> tb;                                                                                         \
		memcpy(tb.t, a->d.table.data(), sizeof(tb.t));                                                                 \
		MSB200_SMEM_OPTIN(resample_up_kernel<R>, a->ctx, sm);                                                          \
		MSB200_LAUNCH(a->ctx, resample_up_kernel<R>, grid, blk, sm, (const short *)d_in_a, in_frames, in_stride, (short *)d_out_a, out_stride, \
		              a->d_hist, tb, a->nch, ring_off, ring_cap, (const short *)d_in_b, (short *)d_out_b, b->d_hist); \
	} while (0)
	if (p.den == 2) RS_UP_PAIR(2);
	else if (p.den == 3) RS_UP_PAIR(3);
	else RS_UP_PAIR(6);
#undef RS_UP_PAIR
	return MSB200_OK;
}
extern "C" {
int msb200_resample_process(msb200_resample *r, const int16_t *in, int in_frames, int16_t *out, int out_stride,
                            int *out_frames) {
	MSB200_CHECK_ARG(r && in && out && in_frames > 0);
	int cap = msb200_resample_max_out(r, in_frames);
	MSB200_CHECK_ARG(out_stride >= cap);
	size_t in_bytes = (size_t)r->live * in_frames * r->nch * 2, out_bytes = (size_t)r->live * out_stride * r->nch * 2;
	int rc;
	if ((rc = r->in.reserve(in_bytes + 16)) || (rc = r->out.reserve(out_bytes + 16))) return rc;
	cudaStream_t s = r->ctx->stream;
	if (in_bytes) MSB200_CUDA(cudaMemcpyAsync(r->in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
	if ((rc = msb200_resample_process_dev(r, r->in.p, in_frames, in_frames, r->out.p, out_stride, out_frames))) return rc;
	if (out_bytes) MSB200_CUDA(cudaMemcpyAsync(out, r->out.p, out_bytes, cudaMemcpyDeviceToHost, s));
	MSB200_HOST_DONE(r->ctx);
	return MSB200_OK;
}
int msb200_resample_set_live(msb200_resample *r, int n_live) {
	MSB200_CHECK_ARG(r && n_live >= 0 && n_live <= r->n);
	r->live = n_live;
	return MSB200_OK;
}

} // extern "C"
