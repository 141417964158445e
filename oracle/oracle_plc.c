/* oracle_plc.c — TEST INFRASTRUCTURE (CPU checker; only tests/, smoke() and bench.py's CPU legs may load it).
 *
 * MSGenericPLC: packet-loss concealment by spectral stretching of the last 50 ms of signal.
 *   filter body      /root/reference/src/audiofilters/msgenericplc.c:61-157 (generic_plc_process)
 *   signal model     /root/reference/src/audiofilters/genericplc.c:28-63 (context, window), :74-110 (generic_plc_fftbf),
 *                    :112-200 (generic_plc_generate_samples), :202-231 (history / continuity buffers), :235-241 (cross-fade)
 *   concealer clock  /root/reference/src/base/mscommon.c:315-362 (MSConcealerContext)
 *   transform        ms_fft / ms_ifft, /root/reference/src/utils/dsptools.c:362-376 -> kiss_fftr2 / kiss_fftri2
 *                    (src/utils/kiss_fftr.c:204-296) over the mixed-radix complex transform of src/utils/kiss_fft.c
 *                    (float build: every fixed-point shift macro is the identity, include/mediastreamer2/dsptools.h:276-290)
 * Restated, not copied: the complex transform is ITERATIVE here (digit-reversal gather, then one pass per factor, each
 * pass a flat loop over independent butterflies — the shape the CUDA kernel uses), while each butterfly keeps the
 * reference's order of float operations so that the results are bit-identical (no FMA contraction: -ffp-contract=off).
 * Supported rates are those whose transform sizes factor into 2, 3, 4 and 5 (8 / 16 / 32 / 48 kHz ...).
 * One deliberate difference: resuming from comfort noise above 16 kHz the reference reads past its 80-sample stack buffer
 * (msgenericplc.c:79-86, undefined behaviour); here the missing samples are the zeros the code evidently means.
 * Pinned: bit-exact against the UNMODIFIED MSGenericPLC filter run in the reference's own MSTicker
 * (oracle/_ref/libms2ref.so) — tests/test_oracle_vs_reference.py::test_plc_*. */
#include "msb200_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float r, i; } cpx;

typedef struct kfft {
	int n, inverse, nf;
	int p[16], m[16], stride[16]; /* factor, sub-length and twiddle stride (= number of blocks) of each level */
	cpx *tw;                      /* n twiddles */
	int *perm;                    /* out[o] = in[perm[o]] */
} kfft;

/* kiss_fft_alloc (kiss_fft.c:438-472) + kf_factor (:405-428) */
static void kfft_init(kfft *k, int n, int inverse) {
	const double pi = 3.14159265358979323846264338327;
	k->n = n;
	k->inverse = inverse;
	k->tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
	k->perm = (int *)malloc(sizeof(int) * (size_t)n);
	for (int i = 0; i < n; ++i) {
		double phase = (-2 * pi / n) * i;
		if (inverse) phase *= -1;
		k->tw[i].r = (float)cos(phase);
		k->tw[i].i = (float)sin(phase);
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
		k->p[nf] = p;
		k->m[nf] = rem;
		k->stride[nf] = stride;
		stride *= p;
		++nf;
	} while (rem > 1);
	k->nf = nf;
	/* kf_shuffle (:276-298): output digit j_d (weight m_d) selects input offset j_d * stride_d */
	for (int o = 0; o < n; ++o) {
		int src = 0;
		for (int d = 0; d < nf; ++d) src += ((o / k->m[d]) % k->p[d]) * k->stride[d];
		k->perm[o] = src;
	}
}
static void kfft_free(kfft *k) {
	free(k->tw);
	free(k->perm);
}
int orc_plc_rate_supported(int rate) {
	int n = ((rate * 2 / 40) / 100) * 100;
	if (n < 4 || (n & 1)) return 0;
	for (int pass = 0; pass < 2; ++pass) {
		int v = pass ? n : n / 2;
		while (v % 2 == 0) v /= 2;
		while (v % 3 == 0) v /= 3;
		while (v % 5 == 0) v /= 5;
		if (v != 1) return 0;
	}
	return 1;
}

static inline cpx cmul(cpx a, cpx b) { /* C_MUL, _kiss_fft_guts.h:109-113 */
	cpx m;
	m.r = a.r * b.r - a.i * b.i;
	m.i = a.r * b.i + a.i * b.r;
	return m;
}
static inline cpx cadd(cpx a, cpx b) { cpx m = {a.r + b.r, a.i + b.i}; return m; }
static inline cpx csub(cpx a, cpx b) { cpx m = {a.r - b.r, a.i - b.i}; return m; }

/* one radix-p butterfly on F[0], F[m], ..., F[(p-1)m]; j = position inside the block, s = twiddle stride */
static void bfly(const kfft *k, cpx *F, int p, int m, int j, int s) {
	const cpx *tw = k->tw;
	if (p == 2) { /* kf_bfly2, kiss_fft.c:36-83 (both directions are the same arithmetic in the float build) */
		cpx t = cmul(F[m], tw[j * s]);
		F[m] = csub(F[0], t);
		F[0] = cadd(F[0], t);
	} else if (p == 4) { /* kf_bfly4 :85-149 */
		cpx s0 = cmul(F[m], tw[j * s]), s1 = cmul(F[2 * m], tw[2 * j * s]), s2 = cmul(F[3 * m], tw[3 * j * s]);
		cpx s5 = csub(F[0], s1);
		F[0] = cadd(F[0], s1);
		cpx s3 = cadd(s0, s2), s4 = csub(s0, s2);
		F[2 * m] = csub(F[0], s3);
		F[0] = cadd(F[0], s3);
		if (k->inverse) {
			F[m].r = s5.r - s4.i;
			F[m].i = s5.i + s4.r;
			F[3 * m].r = s5.r + s4.i;
			F[3 * m].i = s5.i - s4.r;
		} else {
			F[m].r = s5.r + s4.i;
			F[m].i = s5.i - s4.r;
			F[3 * m].r = s5.r - s4.i;
			F[3 * m].i = s5.i + s4.r;
		}
	} else if (p == 3) { /* kf_bfly3 :151-184 */
		const cpx epi3 = tw[s * m];
		cpx s1 = cmul(F[m], tw[j * s]), s2 = cmul(F[2 * m], tw[2 * j * s]);
		cpx s3 = cadd(s1, s2), s0 = csub(s1, s2);
		F[m].r = F[0].r - s3.r * .5f;
		F[m].i = F[0].i - s3.i * .5f;
		s0.r *= epi3.i;
		s0.i *= epi3.i;
		F[0] = cadd(F[0], s3);
		F[2 * m].r = F[m].r + s0.i;
		F[2 * m].i = F[m].i - s0.r;
		F[m].r -= s0.i;
		F[m].i += s0.r;
	} else { /* p == 5: kf_bfly5 :186-243 */
		const cpx ya = tw[s * m], yb = tw[s * 2 * m];
		cpx s0 = F[0];
		cpx s1 = cmul(F[m], tw[j * s]), s2 = cmul(F[2 * m], tw[2 * j * s]);
		cpx s3 = cmul(F[3 * m], tw[3 * j * s]), s4 = cmul(F[4 * m], tw[4 * j * s]);
		cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
		F[0].r += s7.r + s8.r;
		F[0].i += s7.i + s8.i;
		cpx s5, s6, s11, s12;
		s5.r = s0.r + s7.r * ya.r + s8.r * yb.r;
		s5.i = s0.i + s7.i * ya.r + s8.i * yb.r;
		s6.r = s10.i * ya.i + s9.i * yb.i;
		s6.i = -(s10.r * ya.i) - s9.r * yb.i;
		F[m] = csub(s5, s6);
		F[4 * m] = cadd(s5, s6);
		s11.r = s0.r + s7.r * yb.r + s8.r * ya.r;
		s11.i = s0.i + s7.i * yb.r + s8.i * ya.r;
		s12.r = -(s10.i * yb.i) + s9.i * ya.i;
		s12.i = s10.r * yb.i - s9.r * ya.i;
		F[2 * m] = cadd(s11, s12);
		F[3 * m] = csub(s11, s12);
	}
}

/* kiss_fft_stride (:474-484): gather in digit-reversed order, then the levels from the innermost out (kf_work :300-403) */
static void kfft_run(const kfft *k, const cpx *in, cpx *out) {
	for (int o = 0; o < k->n; ++o) out[o] = in[k->perm[o]];
	for (int d = k->nf - 1; d >= 0; --d) {
		const int p = k->p[d], m = k->m[d], nb = k->stride[d];
		for (int b = 0; b < nb * m; ++b) /* nb blocks of p*m points, m butterflies each: all independent */
			bfly(k, out + (b / m) * p * m + (b % m), p, m, b % m, nb);
	}
}

typedef struct rfft {
	kfft sub;
	cpx *super; /* kiss_fftr_alloc, kiss_fftr.c:39-77 */
	cpx *tmp;
	int ncfft;
} rfft;
static void rfft_init(rfft *r, int nfft, int inverse) {
	const double pi = 3.14159265358979323846264338327;
	const int n = nfft >> 1;
	r->ncfft = n;
	kfft_init(&r->sub, n, inverse);
	r->super = (cpx *)malloc(sizeof(cpx) * (size_t)n);
	r->tmp = (cpx *)malloc(sizeof(cpx) * (size_t)n);
	for (int i = 0; i < n; ++i) {
		double phase = pi * (((double)i) / n + .5);
		if (!inverse) phase = -phase;
		r->super[i].r = (float)cos(phase);
		r->super[i].i = (float)sin(phase);
	}
}
static void rfft_free(rfft *r) {
	kfft_free(&r->sub);
	free(r->super);
	free(r->tmp);
}
/* ms_fft (dsptools.c:362-370): kiss_fftr2 (kiss_fftr.c:204-259) then * (1/N); packed [r0, r1, i1, ..., r(N/2)] */
static void rfft_forward(rfft *r, const float *time, float *freq) {
	const int n = r->ncfft;
	kfft_run(&r->sub, (const cpx *)time, r->tmp);
	const cpx *t = r->tmp;
	freq[0] = t[0].r + t[0].i;
	freq[2 * n - 1] = t[0].r - t[0].i;
	for (int k = 1; k <= n / 2; ++k) {
		const float f2r = t[k].r - t[n - k].r, f2i = t[k].i + t[n - k].i;
		const float f1r = t[k].r + t[n - k].r, f1i = t[k].i - t[n - k].i;
		const float twr = f2r * r->super[k].r - f2i * r->super[k].i;
		const float twi = f2i * r->super[k].r + f2r * r->super[k].i;
		freq[2 * k - 1] = .5f * (f1r + twr);
		freq[2 * k] = .5f * (f1i + twi);
		freq[2 * (n - k) - 1] = .5f * (f1r - twr);
		freq[2 * (n - k)] = .5f * (twi - f1i);
	}
	const float scale = 1.f / (float)(2 * n);
	for (int i = 0; i < 2 * n; ++i) freq[i] *= scale;
}
/* ms_ifft (dsptools.c:373-376) = kiss_fftri2 (kiss_fftr.c:261-296), unnormalised */
static void rfft_inverse(rfft *r, const float *freq, float *time) {
	const int n = r->ncfft;
	cpx *t = r->tmp;
	t[0].r = freq[0] + freq[2 * n - 1];
	t[0].i = freq[0] - freq[2 * n - 1];
	for (int k = 1; k <= n / 2; ++k) {
		cpx fk = {freq[2 * k - 1], freq[2 * k]}, fnkc = {freq[2 * (n - k) - 1], -freq[2 * (n - k)]};
		cpx fek = cadd(fk, fnkc), d = csub(fk, fnkc), fok = cmul(d, r->super[k]);
		t[k] = cadd(fek, fok);
		t[n - k] = csub(fek, fok);
		t[n - k].i *= -1;
	}
	kfft_run(&r->sub, t, (cpx *)time);
}

/* ms_ifft for any even size whose half factors into 2, 3, 4, 5 (used by the equalizer oracle: 128 / 256 / 512 points) */
int orc_kiss_irfft(const float *spec, float *out, int nfft) {
	rfft r;
	int v = nfft / 2;
	if (nfft < 4 || (nfft & 1)) return -1;
	while (v % 2 == 0) v /= 2;
	while (v % 3 == 0) v /= 3;
	while (v % 5 == 0) v /= 5;
	if (v != 1) return -1;
	rfft_init(&r, nfft, 1);
	rfft_inverse(&r, spec, out);
	rfft_free(&r);
	return 0;
}
/* ms_fft (scaled by 1/N), same size rule */
int orc_kiss_rfft(const float *time, float *spec, int nfft) {
	rfft r;
	if (nfft < 4 || (nfft & 1)) return -1;
	rfft_init(&r, nfft, 0);
	rfft_forward(&r, time, spec);
	rfft_free(&r);
	return 0;
}

struct orc_plc {
	int rate, N, T;        /* history length (samples), transition length (samples) */
	int16_t *hist;         /* [N]   plc_buffer */
	int16_t *cont;         /* [2T]  continuity_buffer */
	int16_t *gen;          /* [2N]  plc_out_buffer */
	float *window;         /* [N] */
	uint16_t index, used;  /* plc_index, plc_samples_used (16-bit in the reference: they wrap) */
	rfft fwd, inv;
	float *td, *fd, *fdd, *tdd;
	/* concealer clock, mscommon.c:315-362 */
	int64_t sample_time, plc_start_time;
	int cng_set, cng_running;
};

orc_plc *orc_plc_create(int rate) {
	if (!orc_plc_rate_supported(rate)) return NULL;
	orc_plc *c = (orc_plc *)calloc(1, sizeof(*c));
	c->rate = rate;
	c->N = ((rate * 2 / 40) / 100) * 100; /* genericplc.c:42-44 with PLC_BUFFER_LEN = 2 / 40 (genericplc.h:31) */
	c->T = rate * 5 / 1000;               /* TRANSITION_DELAY 5 ms */
	c->hist = (int16_t *)calloc((size_t)c->N, 2);
	c->cont = (int16_t *)calloc((size_t)(2 * c->T), 2);
	c->gen = (int16_t *)calloc((size_t)(2 * c->N), 2);
	c->window = (float *)malloc(sizeof(float) * (size_t)c->N);
	for (int i = 0; i < c->N; ++i) c->window[i] = (float)(0.75 - 0.25 * cos(2 * 3.14159265 * i / c->N)); /* :60-62 */
	rfft_init(&c->fwd, c->N, 0);
	rfft_init(&c->inv, 2 * c->N, 1);
	c->td = (float *)malloc(sizeof(float) * (size_t)c->N);
	c->fd = (float *)malloc(sizeof(float) * (size_t)c->N);
	c->fdd = (float *)malloc(sizeof(float) * (size_t)(2 * c->N));
	c->tdd = (float *)malloc(sizeof(float) * (size_t)(2 * c->N));
	c->sample_time = -1;
	c->plc_start_time = -1;
	return c;
}
void orc_plc_destroy(orc_plc *c) {
	if (!c) return;
	rfft_free(&c->fwd);
	rfft_free(&c->inv);
	free(c->hist); free(c->cont); free(c->gen); free(c->window);
	free(c->td); free(c->fd); free(c->fdd); free(c->tdd);
	free(c);
}
int orc_plc_history_len(const orc_plc *c) { return c->N; }

/* genericplc.c:235-241 */
static void crossfade(int16_t *inout, const int16_t *from, int n) {
	for (int i = 0; i < n; ++i) {
		const float progress = ((float)i) / n;
		inout[i] = (int16_t)((float)from[i] * (1 - progress) + (float)inout[i] * progress);
	}
}
/* genericplc.c:74-110: window, N-point spectrum, every bin moved to twice its index (x 0.85), 2N-point inverse */
static void stretch(orc_plc *c, const int16_t *in, int16_t *out) {
	const int N = c->N;
	for (int i = 0; i < N; ++i) c->td[i] = (float)in[i] * c->window[i];
	rfft_forward(&c->fwd, c->td, c->fd);
	for (int i = 0; i < N; ++i) {
		c->fdd[2 * i] = c->fd[i] * 0.85f;
		c->fdd[2 * i + 1] = 0;
	}
	rfft_inverse(&c->inv, c->fdd, c->tdd);
	for (int i = 0; i < 2 * N; ++i) out[i] = (int16_t)c->tdd[i];
}
/* genericplc.c:202-213 */
static void push_history(orc_plc *c, const int16_t *data, int n) {
	if (n < c->N) {
		memmove(c->hist, c->hist + n, sizeof(int16_t) * (size_t)(c->N - n));
		memcpy(c->hist + c->N - n, data, sizeof(int16_t) * (size_t)n);
	} else memcpy(c->hist, data + n - c->N, sizeof(int16_t) * (size_t)c->N);
}

/* one received block, in place: msgenericplc.c:64-117 (after_cng = the filter was emitting comfort noise) */
void orc_plc_packet(orc_plc *c, int16_t *data, int n, int after_cng) {
	const int T = c->T;
	push_history(c, data, n);
	/* genericplc.c:215-231: delay the stream by T samples through the continuity buffer */
	const int tb = T > n ? n : T;
	int16_t *tail = (int16_t *)malloc(sizeof(int16_t) * (size_t)(tb > 0 ? tb : 1));
	memcpy(tail, data + n - tb, sizeof(int16_t) * (size_t)tb);
	memmove(data + tb, data, sizeof(int16_t) * (size_t)(n - tb));
	memcpy(data, c->cont, sizeof(int16_t) * (size_t)tb);
	memcpy(c->cont, tail, sizeof(int16_t) * (size_t)tb);
	free(tail);
	if (after_cng) { /* :77-88 without G.729B: T zeros, then fade in from zero */
		for (int i = 0; i < T && i < n; ++i) data[i] = 0;
		for (int i = 0; i < T && T + i < n; ++i) {
			const float progress = ((float)i) / T;
			data[T + i] = (int16_t)(0.f * (1 - progress) + (float)data[T + i] * progress);
		}
	}
	if (c->used != 0) { /* :90-112 */
		if (n >= 2 * T) crossfade(data + T, c->cont + T, T);
		else crossfade(c->cont, c->cont + T, T);
	}
	c->index = 0;
	c->used = 0;
}

/* one concealed block of n samples: genericplc.c:112-200, then the history update of msgenericplc.c:149-150 */
void orc_plc_conceal(orc_plc *c, int16_t *data, int n) {
	const int T = c->T, N = c->N, rate = c->rate;
	const int max_len = 150 * rate / 1000, dec_start = 100 * rate / 1000;
	if (c->used >= max_len) {
		c->used = (uint16_t)(c->used + n);
		memset(data, 0, sizeof(int16_t) * (size_t)n);
		memset(c->cont, 0, sizeof(int16_t) * (size_t)(2 * T));
	} else {
		if (c->used == 0) {
			stretch(c, c->hist, c->gen);
			crossfade(c->gen, c->cont, T);
		}
		if (c->index + n + 2 * T > 2 * N) {
			int ready = (uint16_t)(2 * N - c->index - T);
			if (ready > n) ready = n;
			memcpy(data, c->gen + c->index, sizeof(int16_t) * (size_t)ready);
			memcpy(c->cont, c->gen + c->index + ready, sizeof(int16_t) * (size_t)T);
			stretch(c, c->gen, c->gen);
			crossfade(c->gen, c->cont, T);
			if (n != ready) memcpy(data + ready, c->gen, sizeof(int16_t) * (size_t)(n - ready));
			c->index = (uint16_t)(n - ready);
			memcpy(c->cont, c->gen + c->index, sizeof(int16_t) * (size_t)(2 * T));
		} else {
			memcpy(data, c->gen + c->index, sizeof(int16_t) * (size_t)n);
			c->index = (uint16_t)(c->index + n);
			memcpy(c->cont, c->gen + c->index, sizeof(int16_t) * (size_t)(2 * T));
		}
		if (c->used + n > dec_start) { /* fade to silence between 100 and 150 ms */
			int i = dec_start - c->used;
			if (i < 0) i = 0;
			for (; i < n; ++i) {
				if (c->used + i >= max_len) data[i] = 0;
				else
					data[i] = (int16_t)((1.0 + ((float)(dec_start - (c->used + i)) / (float)(50 * rate / 1000))) *
					                    (float)data[i]);
			}
		}
		c->used = (uint16_t)(c->used + n);
	}
	push_history(c, data, n);
}

/* ---- filter level: the concealer clock decides when a block is missing (msgenericplc.c:61-157) ---- */
void orc_plc_filter_set_cn(orc_plc *c) { c->cng_set = 1; } /* MS_GENERIC_PLC_SET_CN :178-183 */

/* a block arrived during the tick at `now_ms` */
void orc_plc_filter_packet(orc_plc *c, uint64_t now_ms, int16_t *data, int n, int nchannels) {
	const unsigned int ms = (unsigned int)((1000 * (size_t)n * 2) / ((size_t)c->rate * 2 * (size_t)nchannels));
	if (c->sample_time == -1) c->sample_time = (int64_t)now_ms; /* ms_concealer_inc_sample_time, mscommon.c:328-343 */
	c->sample_time += ms;
	c->plc_start_time = -1;
	orc_plc_packet(c, data, n, c->cng_running);
	if (c->cng_running) {
		c->cng_running = 0;
		c->cng_set = 0;
	}
}
/* end of the tick: returns the number of samples written to `out` (0 = nothing to emit), *kind = 1 PLC, 2 comfort noise */
int orc_plc_filter_tick(orc_plc *c, uint64_t now_ms, int interval_ms, int nchannels, int16_t *out, int *kind) {
	*kind = 0;
	if (c->sample_time == -1) return 0; /* ms_concealer_context_is_concealement_required, mscommon.c:345-362 */
	if ((uint64_t)c->sample_time > now_ms) return 0;
	if (c->plc_start_time == -1) c->plc_start_time = c->sample_time;
	const uint32_t dur = (uint32_t)(now_ms - (uint64_t)c->plc_start_time);
	if (!(dur < UINT32_MAX)) {
		c->sample_time = -1;
		return 0;
	}
	const int n = c->rate * nchannels * interval_ms / 1000;
	if (c->cng_set) {
		c->cng_set = 0;
		c->cng_running = 1;
		memset(out, 0, sizeof(int16_t) * (size_t)n);
		*kind = 2;
	} else if (c->cng_running) {
		memset(out, 0, sizeof(int16_t) * (size_t)n);
		*kind = 2;
	} else {
		orc_plc_conceal(c, out, n);
		*kind = 1;
	}
	c->sample_time += interval_ms; /* ms_concealer_inc_sample_time(..., FALSE) :155 */
	return n;
}
