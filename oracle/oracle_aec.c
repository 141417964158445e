/* oracle/oracle_aec.c — TEST INFRASTRUCTURE (see msb200_oracle.h).
 *
 * CPU restatement of what MSSpeexEC computes per frame (/root/reference/src/audiofilters/speexec.c:297-298):
 *     speex_echo_cancellation(ecstate, mic, ref, out);  speex_preprocess_run(den, out);
 * configured as speex_ec_preprocess() does (:188-216): frame = adjust_framesize(64, rate) (:171-180), filter length =
 * tail_ms*rate/1000 (:194), SPEEX_ECHO_SET_SAMPLING_RATE (:202), SPEEX_PREPROCESS_SET_ECHO_STATE (:203).
 *
 * speexdsp (mdf.c, preprocess.c, filterbank.c, smallft.c) is an external UN-VENDORED dependency
 * (/root/reference/CMakeLists.txt:208, no version pin) and its source is not in this container. This file restates the
 * published speexdsp 1.2 float-build algorithm (TWO_PATH MDF canceller; preprocessor with denoise + residual-echo
 * suppression, AGC/VAD/dereverb off). PARITY UNPINNED: the reference's tests hold no vectors for MSSpeexEC
 * (SURVEY.md §8c). Known deliberate deviation: the real FFT is our own radix-2 Stockham transform (same packed output
 * format and 1/N forward scaling as spx_fft) instead of smallft's mixed-radix code; results agree to float rounding.
 * Behavioural checks (ERLE on synthetic echo, convergence) live in tests/test_oracle_aec.py.
 */
#include "msb200_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ real FFT
 * Packed format of spx_fft: out = [r0, r1, i1, r2, i2, ..., r(N/2-1), i(N/2-1), r(N/2)], forward scaled by 1/N,
 * inverse unscaled. Implemented as an N/2-point complex radix-2 Stockham FFT + split step. The SAME algorithm (same
 * butterfly order, same double-precision-derived float twiddles, separate multiplies and adds) is used by the CUDA
 * kernel so that the two agree to the last bit wherever the surrounding arithmetic does. */
typedef struct {
	int n;        /* real length N */
	float *twc;   /* [N/4] cos(2 pi j / (N/2)), complex-FFT twiddles */
	float *tws;   /* [N/4] sin(2 pi j / (N/2)) */
	float *splc;  /* [N/2+1] cos(2 pi k / N), split-step twiddles */
	float *spls;  /* [N/2+1] sin(2 pi k / N) */
	float *bufa, *bufb; /* complex scratch, N floats each */
} orc_fft;

static orc_fft *fft_new(int n) {
	orc_fft *t = (orc_fft *)calloc(1, sizeof(*t));
	int L = n / 2;
	t->n = n;
	t->twc = (float *)malloc(sizeof(float) * (size_t)(L / 2));
	t->tws = (float *)malloc(sizeof(float) * (size_t)(L / 2));
	t->splc = (float *)malloc(sizeof(float) * (size_t)(L + 1));
	t->spls = (float *)malloc(sizeof(float) * (size_t)(L + 1));
	t->bufa = (float *)malloc(sizeof(float) * (size_t)n);
	t->bufb = (float *)malloc(sizeof(float) * (size_t)n);
	for (int j = 0; j < L / 2; ++j) {
		t->twc[j] = (float)cos(2.0 * M_PI * j / L);
		t->tws[j] = (float)sin(2.0 * M_PI * j / L);
	}
	for (int k = 0; k <= L; ++k) {
		t->splc[k] = (float)cos(2.0 * M_PI * k / n);
		t->spls[k] = (float)sin(2.0 * M_PI * k / n);
	}
	return t;
}
static void fft_free(orc_fft *t) {
	if (!t) return;
	free(t->twc);
	free(t->tws);
	free(t->splc);
	free(t->spls);
	free(t->bufa);
	free(t->bufb);
	free(t);
}
/* complex Stockham radix-2 (decimation in frequency, autosort); sign=-1 forward, +1 inverse; returns result buffer */
static float *cfft(orc_fft *t, float *x, float *y, int sign) {
	const int L = t->n / 2;
	for (int n = L, s = 1; n > 1; n >>= 1, s <<= 1) {
		const int m = n >> 1;
		for (int o = 0; o < L; ++o) { /* one OUTPUT element per iteration (= one GPU thread) */
			const int q = o % s, tmp = o / s, r = tmp & 1, p = tmp >> 1;
			const float ar = x[2 * (q + s * p)], ai = x[2 * (q + s * p) + 1];
			const float br = x[2 * (q + s * (p + m))], bi = x[2 * (q + s * (p + m)) + 1];
			if (!r) {
				y[2 * o] = ar + br;
				y[2 * o + 1] = ai + bi;
			} else {
				const float wr = t->twc[p * s], wi = (sign < 0 ? -t->tws[p * s] : t->tws[p * s]);
				const float dr = ar - br, di = ai - bi;
				y[2 * o] = dr * wr - di * wi;
				y[2 * o + 1] = dr * wi + di * wr;
			}
		}
		float *sw = x;
		x = y;
		y = sw;
	}
	return x;
}
static void orc_rfft(orc_fft *t, const float *in, float *out) { /* spx_fft */
	const int N = t->n, L = N / 2;
	const float scale = (float)(1. / N);
	float *z = t->bufa;
	for (int i = 0; i < N; ++i)
		z[i] = scale * in[i];
	float *Z = cfft(t, z, t->bufb, -1);
	/* split: X_k = (Z_k + conj(Z_{L-k}))/2 - (i/2) e^{-2 pi i k/N} (Z_k - conj(Z_{L-k})) */
	out[0] = Z[0] + Z[1];
	out[N - 1] = Z[0] - Z[1];
	for (int k = 1; k < L; ++k) {
		const float zr = Z[2 * k], zi = Z[2 * k + 1], yr = Z[2 * (L - k)], yi = -Z[2 * (L - k) + 1];
		const float er = 0.5f * (zr + yr), ei = 0.5f * (zi + yi);
		const float dr = 0.5f * (zr - yr), di = 0.5f * (zi - yi);
		const float c = t->splc[k], s = t->spls[k];
		/* -i * (c - i s) * (dr + i di) = (-s - i c)(dr + i di) = (-s dr + c di) + i(-s di - c dr) */
		out[2 * k - 1] = er + (c * di - s * dr);
		out[2 * k] = ei - (s * di + c * dr);
	}
}
static void orc_irfft(orc_fft *t, const float *in, float *out) { /* spx_ifft */
	const int N = t->n, L = N / 2;
	float *Z = t->bufa;
	/* Z_k = E_k + i O_k, E_k = X_k + conj(X_{L-k}), O_k = (X_k - conj(X_{L-k})) e^{+2 pi i k/N} */
	Z[0] = in[0] + in[N - 1];
	Z[1] = in[0] - in[N - 1];
	for (int k = 1; k < L; ++k) {
		const float xr = in[2 * k - 1], xi = in[2 * k];
		const float yr = (L - k == 0) ? in[0] : in[2 * (L - k) - 1], yi = -in[2 * (L - k)];
		const float er = xr + yr, ei = xi + yi, dr = xr - yr, di = xi - yi;
		const float c = t->splc[k], s = t->spls[k];
		const float orr = dr * c - di * s, oi = dr * s + di * c; /* O_k */
		Z[2 * k] = er - oi;      /* E + iO : real = Er - Oi */
		Z[2 * k + 1] = ei + orr; /*          imag = Ei + Or */
	}
	float *z = cfft(t, Z, t->bufb, +1);
	for (int i = 0; i < N; ++i)
		out[i] = z[i];
}

/* ------------------------------------------------------------------------------------------------ filterbank.c */
#define NB_BANDS 24
static float to_bark(float n) {
	return 13.1f * (float)atan(.00074f * n) + 2.24f * (float)atan(n * n * 1.85e-8f) + 1e-4f * n;
}
typedef struct {
	int nb_banks, len;
	int *bank_left, *bank_right;
	float *filter_left, *filter_right;
} orc_filterbank;

static orc_filterbank *filterbank_new(int banks, float sampling, int len) {
	orc_filterbank *bank = (orc_filterbank *)calloc(1, sizeof(*bank));
	float df = sampling / (2.f * (float)len);
	float max_mel = to_bark(sampling / 2);
	float mel_interval = max_mel / (float)(banks - 1);
	bank->nb_banks = banks;
	bank->len = len;
	bank->bank_left = (int *)calloc((size_t)len, sizeof(int));
	bank->bank_right = (int *)calloc((size_t)len, sizeof(int));
	bank->filter_left = (float *)calloc((size_t)len, sizeof(float));
	bank->filter_right = (float *)calloc((size_t)len, sizeof(float));
	for (int i = 0; i < len; i++) {
		float curr_freq = (float)i * df;
		float mel = to_bark(curr_freq);
		float val;
		int id1;
		if (mel > max_mel) break;
		id1 = (int)(floor(mel / mel_interval));
		if (id1 > banks - 2) {
			id1 = banks - 2;
			val = 1.f;
		} else {
			val = (mel - (float)id1 * mel_interval) / mel_interval;
		}
		bank->bank_left[i] = id1;
		bank->filter_left[i] = 1.f - val;
		bank->bank_right[i] = id1 + 1;
		bank->filter_right[i] = val;
	}
	return bank;
}
static void filterbank_free(orc_filterbank *b) {
	if (!b) return;
	free(b->bank_left);
	free(b->bank_right);
	free(b->filter_left);
	free(b->filter_right);
	free(b);
}
static void filterbank_compute_bank32(const orc_filterbank *bank, const float *ps, float *mel) {
	for (int i = 0; i < bank->nb_banks; i++)
		mel[i] = 0;
	for (int i = 0; i < bank->len; i++) {
		mel[bank->bank_left[i]] += bank->filter_left[i] * ps[i];
		mel[bank->bank_right[i]] += bank->filter_right[i] * ps[i];
	}
}
static void filterbank_compute_psd16(const orc_filterbank *bank, const float *mel, float *ps) {
	for (int i = 0; i < bank->len; i++) {
		float tmp = mel[bank->bank_left[i]] * bank->filter_left[i];
		tmp += mel[bank->bank_right[i]] * bank->filter_right[i];
		ps[i] = tmp;
	}
}

/* ------------------------------------------------------------------------------------------------ state */
struct orc_aec {
	/* ---- mdf.c SpeexEchoState (C = K = 1) */
	int frame_size, window_size, M, cancel_count, adapted, saturated, screwed_up, sampling_rate;
	float spec_average, beta0, beta_max, sum_adapt, leak_estimate;
	float *e, *x, *X, *input, *y, *last_y, *Y, *E, *PHI, *W, *foreground;
	float Davg1, Davg2, Dvar1, Dvar2;
	float *power, *power_1, *wtmp, *Rf, *Yf, *Xf, *Eh, *Yh;
	float Pey, Pyy;
	float *window, *prop;
	float memX, memD, memE, preemph, notch_radius, notch_mem[2];
	orc_fft *fft;
	/* ---- preprocess.c SpeexPreprocessState (ps_size = frame_size) */
	int nbands, nb_adapt, min_count;
	int noise_suppress, echo_suppress, echo_suppress_active;
	orc_filterbank *bank;
	float *frame, *ft, *ps, *gain2, *gain_floor, *pwindow, *noise, *reverb_estimate, *old_ps, *gain, *prior, *post;
	float *S, *Smin, *Stmp, *zeta, *echo_noise, *residual_echo, *inbuf, *outbuf;
	int *update_prob;
	orc_fft *pfft;
};

int orc_aec_frame_size_for_rate(int sample_rate, int framesize_at_8000) { /* adjust_framesize, speexec.c:171-180 */
	int newsize = (framesize_at_8000 * sample_rate) / 8000;
	int n = 1, next;
	while ((next = n << 1) <= newsize)
		n = next;
	return n;
}

static float *falloc(size_t n) {
	return (float *)calloc(n ? n : 1, sizeof(float));
}

static void conj_window(float *w, int len) { /* preprocess.c conj_window, float build */
	for (int i = 0; i < len; i++) {
		float tmp;
		float x = (4.f * (float)i) / (float)len;
		int inv = 0;
		if (x < 1.f) {
		} else if (x < 2.f) {
			x = 2.f - x;
			inv = 1;
		} else if (x < 3.f) {
			x = x - 2.f;
			inv = 1;
		} else {
			x = 2.f - x + 2.f; /* 4 - x */
		}
		x = 1.271903f * x;
		tmp = .5f - .5f * (float)cos(.5 * M_PI * x);
		tmp = tmp * tmp;
		if (inv) tmp = 1.f - tmp;
		w[i] = (float)sqrt(tmp);
	}
}

static void echo_reset(orc_aec *st) { /* speex_echo_state_reset */
	int N = st->window_size, M = st->M;
	st->cancel_count = 0;
	st->screwed_up = 0;
	for (int i = 0; i < N * M; i++)
		st->W[i] = 0, st->foreground[i] = 0;
	for (int i = 0; i < N * (M + 1); i++)
		st->X[i] = 0;
	for (int i = 0; i <= st->frame_size; i++) {
		st->power[i] = 0;
		st->power_1[i] = 1.f;
		st->Eh[i] = 0;
		st->Yh[i] = 0;
	}
	for (int i = 0; i < st->frame_size; i++)
		st->last_y[i] = 0;
	for (int i = 0; i < N; i++) {
		st->E[i] = 0;
		st->x[i] = 0;
	}
	st->notch_mem[0] = st->notch_mem[1] = 0;
	st->memD = st->memE = st->memX = 0;
	st->saturated = 0;
	st->adapted = 0;
	st->sum_adapt = 0;
	st->Pey = st->Pyy = 1.f;
	st->Davg1 = st->Davg2 = 0;
	st->Dvar1 = st->Dvar2 = 0;
}

orc_aec *orc_aec_new(int sample_rate, int tail_length_ms, int framesize_at_8000) {
	orc_aec *st = (orc_aec *)calloc(1, sizeof(*st));
	const int frame_size = orc_aec_frame_size_for_rate(sample_rate, framesize_at_8000);
	const int filter_length = (tail_length_ms * sample_rate) / 1000; /* speexec.c:194 */
	int N, M;
	/* ---- speex_echo_state_init_mc(frame_size, filter_length, 1, 1) */
	st->frame_size = frame_size;
	st->window_size = N = 2 * frame_size;
	st->M = M = (filter_length + frame_size - 1) / frame_size;
	st->fft = fft_new(N);
	st->e = falloc((size_t)N);
	st->x = falloc((size_t)N);
	st->input = falloc((size_t)frame_size);
	st->y = falloc((size_t)N);
	st->last_y = falloc((size_t)N);
	st->Yf = falloc((size_t)frame_size + 1);
	st->Rf = falloc((size_t)frame_size + 1);
	st->Xf = falloc((size_t)frame_size + 1);
	st->Yh = falloc((size_t)frame_size + 1);
	st->Eh = falloc((size_t)frame_size + 1);
	st->X = falloc((size_t)(M + 1) * N);
	st->Y = falloc((size_t)N);
	st->E = falloc((size_t)N);
	st->W = falloc((size_t)M * N);
	st->foreground = falloc((size_t)M * N);
	st->PHI = falloc((size_t)N);
	st->power = falloc((size_t)frame_size + 1);
	st->power_1 = falloc((size_t)frame_size + 1);
	st->window = falloc((size_t)N);
	st->prop = falloc((size_t)M);
	st->wtmp = falloc((size_t)N);
	for (int i = 0; i < N; i++)
		st->window[i] = (float)(.5 - .5 * cos(2 * M_PI * i / N));
	for (int i = 0; i <= frame_size; i++)
		st->power_1[i] = 1.f;
	{
		float sum, decay = (float)exp(-2.4 / M);
		st->prop[0] = .7f;
		sum = st->prop[0];
		for (int i = 1; i < M; i++) {
			st->prop[i] = st->prop[i - 1] * decay;
			sum = sum + st->prop[i];
		}
		for (int i = M - 1; i >= 0; i--)
			st->prop[i] = (.8f * st->prop[i]) / sum;
	}
	st->preemph = .9f;
	st->Pey = st->Pyy = 1.f;
	/* ---- speex_echo_ctl(SPEEX_ECHO_SET_SAMPLING_RATE) */
	st->sampling_rate = sample_rate;
	st->spec_average = (float)frame_size / (float)sample_rate;
	st->beta0 = (2.0f * (float)frame_size) / (float)sample_rate;
	st->beta_max = (.5f * (float)frame_size) / (float)sample_rate;
	if (sample_rate < 12000) st->notch_radius = .9f;
	else if (sample_rate < 24000) st->notch_radius = .982f;
	else st->notch_radius = .992f;

	/* ---- speex_preprocess_state_init(frame_size, sample_rate): ps_size = frame_size, N3 = frame_size, N4 = 0 */
	{
		const int Np = frame_size, Mb = NB_BANDS;
		st->nbands = Mb;
		st->noise_suppress = -15;
		st->echo_suppress = -40;
		st->echo_suppress_active = -15;
		st->bank = filterbank_new(Mb, (float)sample_rate, Np);
		st->frame = falloc((size_t)2 * Np);
		st->pwindow = falloc((size_t)2 * Np);
		st->ft = falloc((size_t)2 * Np);
		st->ps = falloc((size_t)Np + Mb);
		st->noise = falloc((size_t)Np + Mb);
		st->echo_noise = falloc((size_t)Np + Mb);
		st->residual_echo = falloc((size_t)Np + Mb);
		st->reverb_estimate = falloc((size_t)Np + Mb);
		st->old_ps = falloc((size_t)Np + Mb);
		st->prior = falloc((size_t)Np + Mb);
		st->post = falloc((size_t)Np + Mb);
		st->gain = falloc((size_t)Np + Mb);
		st->gain2 = falloc((size_t)Np + Mb);
		st->gain_floor = falloc((size_t)Np + Mb);
		st->zeta = falloc((size_t)Np + Mb);
		st->S = falloc((size_t)Np);
		st->Smin = falloc((size_t)Np);
		st->Stmp = falloc((size_t)Np);
		st->update_prob = (int *)calloc((size_t)Np, sizeof(int));
		st->inbuf = falloc((size_t)Np);
		st->outbuf = falloc((size_t)Np);
		conj_window(st->pwindow, 2 * Np);
		for (int i = 0; i < Np + Mb; i++) {
			st->noise[i] = 1.f;
			st->reverb_estimate[i] = 0;
			st->old_ps[i] = 1.f;
			st->gain[i] = 1.f;
			st->post[i] = 1.f;
			st->prior[i] = 1.f;
		}
		for (int i = 0; i < Np; i++)
			st->update_prob[i] = 1;
		st->pfft = fft_new(2 * Np);
		st->nb_adapt = 0;
		st->min_count = 0;
	}
	return st;
}

void orc_aec_free(orc_aec *st) {
	if (!st) return;
	float *bufs[] = {st->e, st->x, st->X, st->input, st->y, st->last_y, st->Y, st->E, st->PHI, st->W, st->foreground,
	                 st->power, st->power_1, st->wtmp, st->Rf, st->Yf, st->Xf, st->Eh, st->Yh, st->window, st->prop,
	                 st->frame, st->ft, st->ps, st->gain2, st->gain_floor, st->pwindow, st->noise, st->reverb_estimate,
	                 st->old_ps, st->gain, st->prior, st->post, st->S, st->Smin, st->Stmp, st->zeta, st->echo_noise,
	                 st->residual_echo, st->inbuf, st->outbuf};
	for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); ++i)
		free(bufs[i]);
	free(st->update_prob);
	filterbank_free(st->bank);
	fft_free(st->fft);
	fft_free(st->pfft);
	free(st);
}
int orc_aec_frame_size(orc_aec *a) { return a->frame_size; }
int orc_aec_M(orc_aec *a) { return a->M; }

/* ------------------------------------------------------------------------------------------------ mdf.c helpers */
static float mdf_inner_prod(const float *x, const float *y, int len) {
	float sum = 0;
	len >>= 1;
	while (len--) {
		float part = 0;
		part = part + (*x++) * (*y++);
		part = part + (*x++) * (*y++);
		sum = sum + part;
	}
	return sum;
}
static void power_spectrum_accum(const float *X, float *ps, int N) {
	int i, j;
	ps[0] += X[0] * X[0];
	for (i = 1, j = 1; i < N - 1; i += 2, j++)
		ps[j] += X[i] * X[i] + X[i + 1] * X[i + 1];
	ps[j] += X[i] * X[i];
}
static void power_spectrum(const float *X, float *ps, int N) {
	int i, j;
	ps[0] = X[0] * X[0];
	for (i = 1, j = 1; i < N - 1; i += 2, j++)
		ps[j] = X[i] * X[i] + X[i + 1] * X[i + 1];
	ps[j] = X[i] * X[i];
}
static void spectral_mul_accum(const float *X, const float *Y, float *acc, int N, int M) {
	int i, j;
	for (i = 0; i < N; i++)
		acc[i] = 0;
	for (j = 0; j < M; j++) {
		acc[0] += X[0] * Y[0];
		for (i = 1; i < N - 1; i += 2) {
			acc[i] += (X[i] * Y[i] - X[i + 1] * Y[i + 1]);
			acc[i + 1] += (X[i + 1] * Y[i] + X[i] * Y[i + 1]);
		}
		acc[i] += X[i] * Y[i];
		X += N;
		Y += N;
	}
}
static void weighted_spectral_mul_conj(const float *w, float p, const float *X, const float *Y, float *prod, int N) {
	int i, j;
	float W = p * w[0];
	prod[0] = W * (X[0] * Y[0]);
	for (i = 1, j = 1; i < N - 1; i += 2, j++) {
		W = p * w[j];
		prod[i] = W * (X[i] * Y[i] + X[i + 1] * Y[i + 1]);
		prod[i + 1] = W * (-X[i + 1] * Y[i] + X[i] * Y[i + 1]);
	}
	W = p * w[j];
	prod[i] = W * (X[i] * Y[i]);
}
static void mdf_adjust_prop(const float *W, int N, int M, float *prop) {
	float max_sum = 1, prop_sum = 1;
	for (int i = 0; i < M; i++) {
		float tmp = 1;
		for (int j = 0; j < N; j++)
			tmp += W[i * N + j] * W[i * N + j];
		prop[i] = (float)sqrt(tmp);
		if (prop[i] > max_sum) max_sum = prop[i];
	}
	for (int i = 0; i < M; i++) {
		prop[i] += .1f * max_sum;
		prop_sum += prop[i];
	}
	for (int i = 0; i < M; i++)
		prop[i] = (.99f * prop[i]) / prop_sum;
}
static void filter_dc_notch16(const int16_t *in, float radius, float *out, int len, float *mem) {
	float den2 = radius * radius + .7f * (1 - radius) * (1 - radius);
	for (int i = 0; i < len; i++) {
		float vin = in[i];
		float vout = mem[0] + vin;
		mem[0] = mem[1] + 2 * (-vin + radius * vout);
		mem[1] = vin - den2 * vout;
		out[i] = radius * vout;
	}
}
static inline int16_t word2int(float x) {
	return (int16_t)(x < -32767.5f ? -32768 : (x > 32766.5f ? 32767 : (int)floor(.5 + x)));
}

/* ------------------------------------------------------------------------------------------------ speex_echo_cancellation */
void orc_aec_cancel_frame(orc_aec *st, const int16_t *in, const int16_t *far_end, int16_t *out) {
	const int N = st->window_size, M = st->M, F = st->frame_size;
	float Syy, See, Sxx, Sdd, Sff, Dbf, Sey, ss, ss_1, Pey = 1.f, Pyy = 1.f, alpha, alpha_1, RER, tmp32;
	int update_foreground;

	st->cancel_count++;
	ss = .35f / (float)M;
	ss_1 = 1 - ss;

	/* DC notch + pre-emphasis on the microphone signal */
	filter_dc_notch16(in, st->notch_radius, st->input, F, st->notch_mem);
	for (int i = 0; i < F; i++) {
		float t = st->input[i] - st->preemph * st->memD;
		st->memD = st->input[i];
		st->input[i] = t;
	}
	/* shift far-end window, pre-emphasis on the new half */
	for (int i = 0; i < F; i++) {
		float t;
		st->x[i] = st->x[i + F];
		t = (float)far_end[i] - st->preemph * st->memX;
		st->x[i + F] = t;
		st->memX = far_end[i];
	}
	/* shift the spectral history and transform the newest far-end window */
	for (int j = M - 1; j >= 0; j--)
		memcpy(&st->X[(j + 1) * N], &st->X[j * N], sizeof(float) * (size_t)N);
	orc_rfft(st->fft, st->x, &st->X[0]);

	Sxx = 0;
	Sxx += mdf_inner_prod(st->x + F, st->x + F, F);
	for (int i = 0; i <= F; i++)
		st->Xf[i] = 0; /* speex relies on the previous frame's zeroing; equivalent since Xf is re-zeroed below */
	power_spectrum_accum(st->X, st->Xf, N);

	/* foreground filter output */
	Sff = 0;
	spectral_mul_accum(st->X, st->foreground, st->Y, N, M);
	orc_irfft(st->fft, st->Y, st->e);
	for (int i = 0; i < F; i++)
		st->e[i] = st->input[i] - st->e[i + F];
	Sff += mdf_inner_prod(st->e, st->e, F);

	/* adjust proportional adaptation rate */
	if (st->adapted) mdf_adjust_prop(st->W, N, M, st->prop);
	/* weight gradient (uses the PREVIOUS frame's error spectrum E and step sizes power_1) */
	if (st->saturated == 0) {
		for (int j = M - 1; j >= 0; j--) {
			weighted_spectral_mul_conj(st->power_1, st->prop[j], &st->X[(j + 1) * N], st->E, st->PHI, N);
			for (int i = 0; i < N; i++)
				st->W[j * N + i] += st->PHI[i];
		}
	} else {
		st->saturated--;
	}
	/* AUMDF: constrain block 0 and one other block per frame to avoid circular convolution */
	for (int j = 0; j < M; j++) {
		if (j == 0 || st->cancel_count % (M - 1) == j - 1) {
			orc_irfft(st->fft, &st->W[j * N], st->wtmp);
			for (int i = F; i < N; i++)
				st->wtmp[i] = 0;
			orc_rfft(st->fft, st->wtmp, &st->W[j * N]);
		}
	}

	for (int i = 0; i <= F; i++)
		st->Rf[i] = st->Yf[i] = st->Xf[i] = 0;

	Dbf = 0;
	See = 0;
	/* background filter output */
	spectral_mul_accum(st->X, st->W, st->Y, N, M);
	orc_irfft(st->fft, st->Y, st->y);
	for (int i = 0; i < F; i++)
		st->e[i] = st->e[i + F] - st->y[i + F];
	Dbf += 10 + mdf_inner_prod(st->e, st->e, F);
	for (int i = 0; i < F; i++)
		st->e[i] = st->input[i] - st->y[i + F];
	See += mdf_inner_prod(st->e, st->e, F);

	/* two-path logic */
	st->Davg1 = .6f * st->Davg1 + .4f * (Sff - See);
	st->Davg2 = .85f * st->Davg2 + .15f * (Sff - See);
	st->Dvar1 = .36f * st->Dvar1 + .16f * Sff * Dbf;
	st->Dvar2 = .7225f * st->Dvar2 + .0225f * Sff * Dbf;

	update_foreground = 0;
	if ((Sff - See) * fabsf(Sff - See) > (Sff * Dbf)) update_foreground = 1;
	else if ((st->Davg1 * fabsf(st->Davg1)) > (.5f * st->Dvar1)) update_foreground = 1;
	else if ((st->Davg2 * fabsf(st->Davg2)) > (.25f * st->Dvar2)) update_foreground = 1;

	if (update_foreground) {
		st->Davg1 = st->Davg2 = 0;
		st->Dvar1 = st->Dvar2 = 0;
		memcpy(st->foreground, st->W, sizeof(float) * (size_t)N * M);
		/* smooth transition to avoid blocking artifacts */
		for (int i = 0; i < F; i++)
			st->e[i + F] = st->window[i + F] * st->e[i + F] + st->window[i] * st->y[i + F];
	} else {
		int reset_background = 0;
		if ((-(Sff - See) * fabsf(Sff - See)) > (4.f * (Sff * Dbf))) reset_background = 1;
		if ((-st->Davg1 * fabsf(st->Davg1)) > (4.f * st->Dvar1)) reset_background = 1;
		if ((-st->Davg2 * fabsf(st->Davg2)) > (4.f * st->Dvar2)) reset_background = 1;
		if (reset_background) {
			memcpy(st->W, st->foreground, sizeof(float) * (size_t)N * M);
			for (int i = 0; i < F; i++)
				st->y[i + F] = st->e[i + F];
			for (int i = 0; i < F; i++)
				st->e[i] = st->input[i] - st->y[i + F];
			See = Sff;
			st->Davg1 = st->Davg2 = 0;
			st->Dvar1 = st->Dvar2 = 0;
		}
	}

	Sey = Syy = Sdd = 0;
	/* output with de-emphasis */
	for (int i = 0; i < F; i++) {
		float tmp_out = st->input[i] - st->e[i + F];
		tmp_out = tmp_out + st->preemph * st->memE;
		if (in[i] <= -32000 || in[i] >= 32000) {
			if (st->saturated == 0) st->saturated = 1;
		}
		out[i] = word2int(tmp_out);
		st->memE = tmp_out;
	}
	/* error signal for the filter update */
	for (int i = 0; i < F; i++) {
		st->e[i + F] = st->e[i];
		st->e[i] = 0;
	}
	Sey += mdf_inner_prod(st->e + F, st->y + F, F);
	Syy += mdf_inner_prod(st->y + F, st->y + F, F);
	Sdd += mdf_inner_prod(st->input, st->input, F);
	orc_rfft(st->fft, st->e, st->E);
	for (int i = 0; i < F; i++)
		st->y[i] = 0;
	orc_rfft(st->fft, st->y, st->Y);
	power_spectrum_accum(st->E, st->Rf, N);
	power_spectrum_accum(st->Y, st->Yf, N);

	/* sanity checks */
	if (!(Syy >= 0 && Sxx >= 0 && See >= 0) || !(Sff < N * 1e9 && Syy < N * 1e9 && Sxx < N * 1e9)) {
		st->screwed_up += 50;
		for (int i = 0; i < F; i++)
			out[i] = 0;
	} else if (Sff > Sdd + (float)(N * 10000)) {
		st->screwed_up++;
	} else {
		st->screwed_up = 0;
	}
	if (st->screwed_up >= 50) {
		echo_reset(st);
		return;
	}

	See = See > (float)(N * 100) ? See : (float)(N * 100);

	Sxx += mdf_inner_prod(st->x + F, st->x + F, F);
	power_spectrum_accum(st->X, st->Xf, N);

	/* smooth far-end energy estimate */
	for (int j = 0; j <= F; j++)
		st->power[j] = ss_1 * st->power[j] + 1 + ss * st->Xf[j];

	/* filtered spectra and cross-correlations */
	for (int j = F; j >= 0; j--) {
		float Eh = st->Rf[j] - st->Eh[j];
		float Yh = st->Yf[j] - st->Yh[j];
		Pey = Pey + Eh * Yh;
		Pyy = Pyy + Yh * Yh;
		st->Eh[j] = (1 - st->spec_average) * st->Eh[j] + st->spec_average * st->Rf[j];
		st->Yh[j] = (1 - st->spec_average) * st->Yh[j] + st->spec_average * st->Yf[j];
	}
	Pyy = (float)sqrt(Pyy);
	Pey = Pey / Pyy;

	/* correlation update rate */
	tmp32 = st->beta0 * Syy;
	if (tmp32 > st->beta_max * See) tmp32 = st->beta_max * See;
	alpha = tmp32 / See;
	alpha_1 = 1.f - alpha;
	st->Pey = alpha_1 * st->Pey + alpha * Pey;
	st->Pyy = alpha_1 * st->Pyy + alpha * Pyy;
	if (st->Pyy < 1.f) st->Pyy = 1.f;
	if (st->Pey < .005f * st->Pyy) st->Pey = .005f * st->Pyy; /* MIN_LEAK */
	if (st->Pey > st->Pyy) st->Pey = st->Pyy;
	st->leak_estimate = st->Pey / st->Pyy;
	if (st->leak_estimate > 16383) st->leak_estimate = 32767;

	/* residual-to-error ratio */
	RER = (.0001f * Sxx + 3.f * st->leak_estimate * Syy) / See;
	if (RER < Sey * Sey / (1 + See * Syy)) RER = Sey * Sey / (1 + See * Syy);
	if (RER > .5f) RER = .5f;

	if (!st->adapted && st->sum_adapt > (float)M && st->leak_estimate * Syy > .03f * Syy) st->adapted = 1;

	if (st->adapted) {
		for (int i = 0; i <= F; i++) {
			float r, e;
			r = st->leak_estimate * st->Yf[i];
			e = st->Rf[i] + 1;
			if (r > .5f * e) r = .5f * e;
			r = .7f * r + .3f * (RER * e);
			st->power_1[i] = r / (e * (st->power[i] + 10));
		}
	} else {
		float adapt_rate = 0;
		if (Sxx > (float)(N * 1000)) {
			tmp32 = .25f * Sxx;
			if (tmp32 > .25f * See) tmp32 = .25f * See;
			adapt_rate = tmp32 / See;
		}
		for (int i = 0; i <= F; i++)
			st->power_1[i] = adapt_rate / (st->power[i] + 10);
		st->sum_adapt = st->sum_adapt + adapt_rate;
	}

	/* residual echo estimate input for the preprocessor */
	for (int i = 0; i < F; i++)
		st->last_y[i] = st->last_y[F + i];
	if (st->adapted) {
		for (int i = 0; i < F; i++)
			st->last_y[F + i] = (float)(in[i] - out[i]);
	}
}

/* speex_echo_get_residual */
static void echo_get_residual(orc_aec *st, float *residual_echo) {
	const int N = st->window_size, F = st->frame_size;
	float leak2;
	for (int i = 0; i < N; i++)
		st->y[i] = st->window[i] * st->last_y[i];
	orc_rfft(st->fft, st->y, st->Y);
	power_spectrum(st->Y, residual_echo, N);
	if (st->leak_estimate > .5f) leak2 = 1;
	else leak2 = 2 * st->leak_estimate;
	for (int i = 0; i <= F; i++)
		residual_echo[i] = (float)(int32_t)(leak2 * residual_echo[i]);
}

/* ------------------------------------------------------------------------------------------------ preprocess.c */
static float hypergeom_gain(float xx) {
	static const float table[21] = {0.82157f, 1.02017f, 1.20461f, 1.37534f, 1.53363f, 1.68092f, 1.81865f,
	                                1.94811f, 2.07038f, 2.18638f, 2.29688f, 2.40255f, 2.50391f, 2.60144f,
	                                2.69551f, 2.78647f, 2.87458f, 2.96015f, 3.04333f, 3.12431f, 3.20326f};
	float x = xx;
	float integer = (float)floor(2 * x);
	int ind = (int)integer;
	float frac;
	if (ind < 0) return 1.f;
	if (ind > 19) return (float)(1 + .1296 / x);
	frac = 2 * x - integer;
	return (float)(((1 - frac) * table[ind] + frac * table[ind + 1]) / sqrt(x + .0001f));
}
static float qcurve(float x) {
	return 1.f / (1.f + .15f / x);
}

void orc_aec_preprocess_frame(orc_aec *st, int16_t *x) { /* speex_preprocess_run with echo_state bound */
	const int N = st->frame_size, M = st->nbands;
	float *ps = st->ps;
	float Zframe, Pframe, beta, beta_1, effective_echo_suppress;

	st->nb_adapt++;
	if (st->nb_adapt > 20000) st->nb_adapt = 20000;
	st->min_count++;
	beta = 1.0f / (float)st->nb_adapt;
	if (beta < .03f) beta = .03f;
	beta_1 = 1.f - beta;

	/* residual echo from the canceller */
	echo_get_residual(st, st->residual_echo);
	if (!(st->residual_echo[0] >= 0 && st->residual_echo[0] < N * 1e9f)) {
		for (int i = 0; i < N; i++)
			st->residual_echo[i] = 0;
	}
	for (int i = 0; i < N; i++) {
		float a = .6f * st->echo_noise[i];
		st->echo_noise[i] = a > st->residual_echo[i] ? a : st->residual_echo[i];
	}
	filterbank_compute_bank32(st->bank, st->echo_noise, st->echo_noise + N);

	/* preprocess_analysis: N3 = N, N4 = 0 */
	for (int i = 0; i < N; i++)
		st->frame[i] = st->inbuf[i];
	for (int i = 0; i < N; i++)
		st->frame[N + i] = x[i];
	for (int i = 0; i < N; i++)
		st->inbuf[i] = x[i];
	for (int i = 0; i < 2 * N; i++)
		st->frame[i] = st->frame[i] * st->pwindow[i];
	orc_rfft(st->pfft, st->frame, st->ft);
	ps[0] = st->ft[0] * st->ft[0];
	for (int i = 1; i < N; i++)
		ps[i] = st->ft[2 * i - 1] * st->ft[2 * i - 1] + st->ft[2 * i] * st->ft[2 * i];
	filterbank_compute_bank32(st->bank, ps, ps + N);

	/* update_noise_prob */
	for (int i = 1; i < N - 1; i++)
		st->S[i] = .8f * st->S[i] + .05f * ps[i - 1] + .1f * ps[i] + .05f * ps[i + 1];
	st->S[0] = .8f * st->S[0] + .2f * ps[0];
	st->S[N - 1] = .8f * st->S[N - 1] + .2f * ps[N - 1];
	if (st->nb_adapt == 1) {
		for (int i = 0; i < N; i++)
			st->Smin[i] = st->Stmp[i] = 0;
	}
	{
		int min_range;
		if (st->nb_adapt < 100) min_range = 15;
		else if (st->nb_adapt < 1000) min_range = 50;
		else if (st->nb_adapt < 10000) min_range = 150;
		else min_range = 300;
		if (st->min_count > min_range) {
			st->min_count = 0;
			for (int i = 0; i < N; i++) {
				st->Smin[i] = st->Stmp[i] < st->S[i] ? st->Stmp[i] : st->S[i];
				st->Stmp[i] = st->S[i];
			}
		} else {
			for (int i = 0; i < N; i++) {
				st->Smin[i] = st->Smin[i] < st->S[i] ? st->Smin[i] : st->S[i];
				st->Stmp[i] = st->Stmp[i] < st->S[i] ? st->Stmp[i] : st->S[i];
			}
		}
	}
	for (int i = 0; i < N; i++)
		st->update_prob[i] = (.4f * st->S[i] > st->Smin[i]) ? 1 : 0;

	/* noise estimate */
	for (int i = 0; i < N; i++) {
		if (!st->update_prob[i] || ps[i] < st->noise[i]) {
			float v = beta_1 * st->noise[i] + beta * ps[i];
			st->noise[i] = v > 0 ? v : 0;
		}
	}
	filterbank_compute_bank32(st->bank, st->noise, st->noise + N);

	if (st->nb_adapt == 1)
		for (int i = 0; i < N + M; i++)
			st->old_ps[i] = ps[i];

	/* a posteriori / a priori SNR */
	for (int i = 0; i < N + M; i++) {
		float gamma;
		float tot_noise = 1.f + st->noise[i] + st->echo_noise[i] + st->reverb_estimate[i];
		st->post[i] = ps[i] / tot_noise - 1.f;
		if (st->post[i] > 100.f) st->post[i] = 100.f;
		{
			float r = st->old_ps[i] / (st->old_ps[i] + tot_noise);
			gamma = .1f + .89f * (r * r);
		}
		st->prior[i] = gamma * (st->post[i] > 0 ? st->post[i] : 0) + (1.f - gamma) * (st->old_ps[i] / tot_noise);
		if (st->prior[i] > 100.f) st->prior[i] = 100.f;
	}

	/* smoothed a priori SNR */
	st->zeta[0] = .7f * st->zeta[0] + .3f * st->prior[0];
	for (int i = 1; i < N - 1; i++)
		st->zeta[i] = .7f * st->zeta[i] + .15f * st->prior[i] + .075f * st->prior[i - 1] + .075f * st->prior[i + 1];
	for (int i = N - 1; i < N + M; i++)
		st->zeta[i] = .7f * st->zeta[i] + .3f * st->prior[i];

	Zframe = 0;
	for (int i = N; i < N + M; i++)
		Zframe = Zframe + st->zeta[i];
	Pframe = .1f + .899f * qcurve(Zframe / (float)st->nbands);

	effective_echo_suppress = (1.f - Pframe) * (float)st->echo_suppress + Pframe * (float)st->echo_suppress_active;
	{ /* compute_gain_floor on the Bark bands */
		float noise_floor = (float)exp(.2302585f * (float)st->noise_suppress);
		float echo_floor = (float)exp(.2302585f * effective_echo_suppress);
		for (int i = N; i < N + M; i++)
			st->gain_floor[i] = (float)(sqrt(noise_floor * st->noise[i] + echo_floor * st->echo_noise[i]) /
			                            sqrt(1 + st->noise[i] + st->echo_noise[i]));
	}

	/* Ephraim-Malah gain and speech presence probability per Bark band */
	for (int i = N; i < N + M; i++) {
		float theta, MM, prior_ratio, P1, q;
		prior_ratio = st->prior[i] / (st->prior[i] + 1.f);
		theta = prior_ratio * (1.f + st->post[i]);
		MM = hypergeom_gain(theta);
		st->gain[i] = prior_ratio * MM;
		if (st->gain[i] > 1.f) st->gain[i] = 1.f;
		st->old_ps[i] = .2f * st->old_ps[i] + (.8f * (st->gain[i] * st->gain[i])) * ps[i];
		P1 = .199f + .8f * qcurve(st->zeta[i]);
		q = 1.f - Pframe * P1;
		st->gain2[i] = (float)(1 / (1.f + (q / (1.f - q)) * (1 + st->prior[i]) * exp(-theta)));
	}
	filterbank_compute_psd16(st->bank, st->gain2 + N, st->gain2);
	filterbank_compute_psd16(st->bank, st->gain + N, st->gain);
	filterbank_compute_psd16(st->bank, st->gain_floor + N, st->gain_floor);

	/* linear-frequency gain */
	for (int i = 0; i < N; i++) {
		float MM, theta, prior_ratio, tmp, p, g;
		prior_ratio = st->prior[i] / (st->prior[i] + 1.f);
		theta = prior_ratio * (1.f + st->post[i]);
		MM = hypergeom_gain(theta);
		g = prior_ratio * MM;
		if (g > 1.f) g = 1.f;
		p = st->gain2[i];
		if (.333f * g > st->gain[i]) g = 3.f * st->gain[i];
		st->gain[i] = g;
		st->old_ps[i] = .2f * st->old_ps[i] + (.8f * (st->gain[i] * st->gain[i])) * ps[i];
		if (st->gain[i] < st->gain_floor[i]) st->gain[i] = st->gain_floor[i];
		tmp = p * (float)sqrt(st->gain[i]) + (1.f - p) * (float)sqrt(st->gain_floor[i]);
		st->gain2[i] = tmp * tmp;
	}

	/* apply gain */
	for (int i = 1; i < N; i++) {
		st->ft[2 * i - 1] = st->gain2[i] * st->ft[2 * i - 1];
		st->ft[2 * i] = st->gain2[i] * st->ft[2 * i];
	}
	st->ft[0] = st->gain2[0] * st->ft[0];
	st->ft[2 * N - 1] = st->gain2[N - 1] * st->ft[2 * N - 1];

	orc_irfft(st->pfft, st->ft, st->frame);
	for (int i = 0; i < 2 * N; i++)
		st->frame[i] = st->frame[i] * st->pwindow[i];
	for (int i = 0; i < N; i++)
		x[i] = word2int(st->outbuf[i] + st->frame[i]);
	for (int i = 0; i < N; i++)
		st->outbuf[i] = st->frame[N + i];
}

void orc_aec_process_frame(orc_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out) {
	orc_aec_cancel_frame(a, mic, ref, out);
	orc_aec_preprocess_frame(a, out);
}

/* ------------------------------------------------------------------------------------------------ probes */
int orc_aec_probe(orc_aec *st, const char *what, float *out, int max_floats) {
	const int N = st->window_size, M = st->M, F = st->frame_size;
	const float *src = NULL;
	int n = 0;
	float scal[16];
	if (!strcmp(what, "W")) src = st->W, n = M * N;
	else if (!strcmp(what, "foreground")) src = st->foreground, n = M * N;
	else if (!strcmp(what, "X")) src = st->X, n = (M + 1) * N;
	else if (!strcmp(what, "E")) src = st->E, n = N;
	else if (!strcmp(what, "power")) src = st->power, n = F + 1;
	else if (!strcmp(what, "power_1")) src = st->power_1, n = F + 1;
	else if (!strcmp(what, "prop")) src = st->prop, n = M;
	else if (!strcmp(what, "last_y")) src = st->last_y, n = N;
	else if (!strcmp(what, "noise")) src = st->noise, n = F + NB_BANDS;
	else if (!strcmp(what, "echo_noise")) src = st->echo_noise, n = F + NB_BANDS;
	else if (!strcmp(what, "gain2")) src = st->gain2, n = F;
	else if (!strcmp(what, "ps")) src = st->ps, n = F + NB_BANDS;
	else if (!strcmp(what, "old_ps")) src = st->old_ps, n = F + NB_BANDS;
	else if (!strcmp(what, "window")) src = st->window, n = N;
	else if (!strcmp(what, "pwindow")) src = st->pwindow, n = N;
	else if (!strcmp(what, "scalars")) {
		scal[0] = (float)st->adapted;
		scal[1] = st->sum_adapt;
		scal[2] = st->leak_estimate;
		scal[3] = st->Pey;
		scal[4] = st->Pyy;
		scal[5] = st->Davg1;
		scal[6] = st->Davg2;
		scal[7] = st->Dvar1;
		scal[8] = st->Dvar2;
		scal[9] = (float)st->saturated;
		scal[10] = (float)st->screwed_up;
		scal[11] = (float)st->cancel_count;
		scal[12] = st->memE;
		scal[13] = st->memD;
		scal[14] = st->memX;
		scal[15] = (float)st->nb_adapt;
		src = scal;
		n = 16;
	} else return -1;
	if (n > max_floats) n = max_floats;
	memcpy(out, src, sizeof(float) * (size_t)n);
	return n;
}

/* test hooks for the FFT itself (tests/test_oracle_aec.py checks it against numpy.fft.rfft) */
void orc_test_rfft(int n, const float *in, float *out) {
	orc_fft *t = fft_new(n);
	orc_rfft(t, in, out);
	fft_free(t);
}
void orc_test_irfft(int n, const float *in, float *out) {
	orc_fft *t = fft_new(n);
	orc_irfft(t, in, out);
	fft_free(t);
}
