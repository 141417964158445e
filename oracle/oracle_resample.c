/* oracle/oracle_resample.c — TEST INFRASTRUCTURE (see msb200_oracle.h).
 *
 * CPU restatement of the speexdsp 1.2 resampler, float build, as mediastreamer2 calls it:
 *   speex_resampler_init(nch, in, out, SPEEX_RESAMPLER_QUALITY_VOIP=3)   /root/reference/src/audiofilters/msresample.c:102-115
 *   speex_resampler_process_int / _process_interleaved_int                 :157-161
 * speexdsp is an external, UN-VENDORED dependency (find_package(SpeexDSP), /root/reference/CMakeLists.txt:208, no version
 * pin; upstream 1.2.x). Its source is not available in this container, so this file restates the published algorithm
 * (resample.c: update_filter, sinc, compute_func, resampler_basic_direct_single, resampler_basic_interpolate_single,
 * speex_resampler_process_native, speex_resampler_process_int). PARITY UNPINNED against the real library: every constant
 * below is to be re-verified when a speexdsp tree is available. Independent cross-check: scipy.signal.resample_poly
 * (tests/test_oracle_resample.py).
 *
 * Deliberate choice: the inner product is accumulated sequentially in float32 (the non-SSE upstream loop); the x86 SSE
 * build of the library sums in a different order, so even the real library is not bit-reproducible across builds.
 */
#include "msb200_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* resample.c: kaiser tables, 32x oversampled, cubic-interpolated by compute_func() */
static const double kaiser8_table[36] = {
    0.99537781, 1.00000000, 0.99537781, 0.98162644, 0.95908712, 0.92831446, 0.89005583, 0.84522401, 0.79486424,
    0.74011713, 0.68217934, 0.62226347, 0.56155915, 0.50119680, 0.44221549, 0.38553619, 0.33194107, 0.28205962,
    0.23636152, 0.19515633, 0.15859932, 0.12670280, 0.09935205, 0.07632451, 0.05731132, 0.04193980, 0.02979584,
    0.02044510, 0.01345224, 0.00839739, 0.00481569, 0.00247437, 0.00112393, 0.00042834, 0.00011921, 0.00000000};
static const double kaiser6_table[36] = {
    0.99733006, 1.00000000, 0.99733006, 0.98935595, 0.97618418, 0.95799003, 0.93501423, 0.90755855, 0.87598009,
    0.84068475, 0.80211977, 0.76076565, 0.71712752, 0.67172623, 0.62508937, 0.57774224, 0.53019925, 0.48295561,
    0.43647969, 0.39120616, 0.34752997, 0.30580127, 0.26632152, 0.22934058, 0.19505503, 0.16360756, 0.13508755,
    0.10953262, 0.08693120, 0.06722600, 0.05031820, 0.03607231, 0.02432151, 0.01487334, 0.00752000, 0.00000000};

struct func_def {
	const double *table;
	int oversample;
};
static const struct func_def KAISER8 = {kaiser8_table, 32};
static const struct func_def KAISER6 = {kaiser6_table, 32};

struct quality_mapping {
	int base_length;
	int oversample;
	float downsample_bandwidth;
	float upsample_bandwidth;
	const struct func_def *window_func;
};
/* only the entries mediastreamer2 can select: Q0 ("min", ARM without NEON) and Q3 ("voip") — msresample.c:104-110 */
static const struct quality_mapping quality_map[4] = {
    {8, 4, 0.830f, 0.860f, &KAISER6},
    {16, 4, 0.850f, 0.880f, &KAISER6},
    {32, 4, 0.882f, 0.910f, &KAISER6},
    {48, 8, 0.895f, 0.917f, &KAISER8},
};

struct orc_resampler {
	uint32_t in_rate, out_rate, num_rate, den_rate;
	int quality, nb_channels;
	uint32_t filt_len, mem_alloc_size, buffer_size;
	int int_advance, frac_advance;
	float cutoff;
	uint32_t oversample;
	int use_direct;
	int32_t *last_sample;
	uint32_t *samp_frac_num;
	float *mem;
	float *sinc_table;
	uint32_t sinc_table_length;
};

static double compute_func(float x, const struct func_def *func) {
	float y, frac;
	double interp[4];
	int ind;
	y = x * (float)func->oversample;
	ind = (int)floor(y);
	frac = (y - (float)ind);
	interp[3] = -0.1666666667 * frac + 0.1666666667 * (frac * frac * frac);
	interp[2] = frac + 0.5 * (frac * frac) - 0.5 * (frac * frac * frac);
	interp[0] = -0.3333333333 * frac + 0.5 * (frac * frac) - 0.1666666667 * (frac * frac * frac);
	interp[1] = 1.f - interp[3] - interp[2] - interp[0];
	return interp[0] * func->table[ind] + interp[1] * func->table[ind + 1] + interp[2] * func->table[ind + 2] +
	       interp[3] * func->table[ind + 3];
}

static float sinc(float cutoff, float x, int N, const struct func_def *window_func) {
	float xx = x * cutoff;
	if (fabs(x) < 1e-6) return cutoff;
	else if (fabs(x) > .5 * N) return 0;
	return (float)(cutoff * sin(M_PI * xx) / (M_PI * xx) * compute_func((float)fabs(2. * x / N), window_func));
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) {
	while (b) {
		uint32_t t = a % b;
		a = b;
		b = t;
	}
	return a;
}

static void update_filter(orc_resampler *st) {
	const struct quality_mapping *q = &quality_map[st->quality];
	st->int_advance = (int)(st->num_rate / st->den_rate);
	st->frac_advance = (int)(st->num_rate % st->den_rate);
	st->oversample = (uint32_t)q->oversample;
	st->filt_len = (uint32_t)q->base_length;
	if (st->num_rate > st->den_rate) {
		/* down-sampling */
		st->cutoff = q->downsample_bandwidth * (float)st->den_rate / (float)st->num_rate;
		st->filt_len = (uint32_t)(((uint64_t)st->filt_len * st->num_rate) / st->den_rate);
		/* round up to a multiple of 8 */
		st->filt_len = ((st->filt_len - 1) & (~0x7u)) + 8;
		if (2 * st->den_rate < st->num_rate) st->oversample >>= 1;
		if (4 * st->den_rate < st->num_rate) st->oversample >>= 1;
		if (8 * st->den_rate < st->num_rate) st->oversample >>= 1;
		if (16 * st->den_rate < st->num_rate) st->oversample >>= 1;
		if (st->oversample < 1) st->oversample = 1;
	} else {
		st->cutoff = q->upsample_bandwidth;
	}
	st->use_direct = st->filt_len * st->den_rate <= st->filt_len * st->oversample + 8;
	if (st->use_direct) {
		st->sinc_table_length = st->filt_len * st->den_rate;
		st->sinc_table = (float *)calloc(st->sinc_table_length, sizeof(float));
		for (uint32_t i = 0; i < st->den_rate; i++)
			for (int32_t j = 0; j < (int32_t)st->filt_len; j++)
				st->sinc_table[i * st->filt_len + (uint32_t)j] =
				    sinc(st->cutoff, ((float)(j - (int32_t)st->filt_len / 2 + 1) - ((float)i) / (float)st->den_rate),
				         (int)st->filt_len, q->window_func);
	} else {
		st->sinc_table_length = st->filt_len * st->oversample + 8;
		st->sinc_table = (float *)calloc(st->sinc_table_length, sizeof(float));
		for (int32_t i = -4; i < (int32_t)(st->oversample * st->filt_len + 4); i++)
			st->sinc_table[i + 4] = sinc(st->cutoff, ((float)i / (float)st->oversample - (float)(st->filt_len / 2)),
			                             (int)st->filt_len, q->window_func);
	}
	/* fresh state: filt_len-1 zeros of history, no zero skipping (mediastreamer2 never calls skip_zeros) */
	st->buffer_size = 160;
	st->mem_alloc_size = st->filt_len - 1 + st->buffer_size;
	st->mem = (float *)calloc((size_t)st->nb_channels * st->mem_alloc_size, sizeof(float));
}

orc_resampler *orc_resampler_new(int nch, int in_rate, int out_rate, int quality) {
	if (quality < 0 || quality > 3 || nch < 1 || in_rate <= 0 || out_rate <= 0) return NULL;
	orc_resampler *st = (orc_resampler *)calloc(1, sizeof(*st));
	uint32_t g = gcd_u32((uint32_t)in_rate, (uint32_t)out_rate);
	st->in_rate = (uint32_t)in_rate;
	st->out_rate = (uint32_t)out_rate;
	st->num_rate = (uint32_t)in_rate / g;
	st->den_rate = (uint32_t)out_rate / g;
	st->quality = quality;
	st->nb_channels = nch;
	st->last_sample = (int32_t *)calloc((size_t)nch, sizeof(int32_t));
	st->samp_frac_num = (uint32_t *)calloc((size_t)nch, sizeof(uint32_t));
	update_filter(st);
	return st;
}

void orc_resampler_free(orc_resampler *st) {
	if (!st) return;
	free(st->last_sample);
	free(st->samp_frac_num);
	free(st->mem);
	free(st->sinc_table);
	free(st);
}

int orc_resampler_filt_len(orc_resampler *r) { return (int)r->filt_len; }
int orc_resampler_den(orc_resampler *r) { return (int)r->den_rate; }
int orc_resampler_num(orc_resampler *r) { return (int)r->num_rate; }
int orc_resampler_use_direct(orc_resampler *r) { return r->use_direct; }
int orc_resampler_oversample(orc_resampler *r) { return (int)r->oversample; }
const float *orc_resampler_table(orc_resampler *r, int *len) {
	if (len) *len = (int)r->sinc_table_length;
	return r->sinc_table;
}

/* resampler_basic_direct_single */
static int basic_direct(orc_resampler *st, int ch, const float *in, uint32_t *in_len, float *out, uint32_t *out_len) {
	const int N = (int)st->filt_len;
	int out_sample = 0;
	int last_sample = st->last_sample[ch];
	uint32_t samp_frac_num = st->samp_frac_num[ch];
	while (!(last_sample >= (int32_t)*in_len || out_sample >= (int32_t)*out_len)) {
		const float *sinct = &st->sinc_table[samp_frac_num * (uint32_t)N];
		const float *iptr = &in[last_sample];
		float sum = 0;
		for (int j = 0; j < N; j++)
			sum += sinct[j] * iptr[j];
		out[out_sample++] = sum;
		last_sample += st->int_advance;
		samp_frac_num += (uint32_t)st->frac_advance;
		if (samp_frac_num >= st->den_rate) {
			samp_frac_num -= st->den_rate;
			last_sample++;
		}
	}
	st->last_sample[ch] = last_sample;
	st->samp_frac_num[ch] = samp_frac_num;
	return out_sample;
}

static void cubic_coef(float frac, float interp[4]) {
	interp[0] = -0.16667f * frac + 0.16667f * frac * frac * frac;
	interp[1] = frac + 0.5f * frac * frac - 0.5f * frac * frac * frac;
	interp[3] = -0.33333f * frac + 0.5f * frac * frac - 0.16667f * frac * frac * frac;
	interp[2] = (float)(1. - interp[0] - interp[1] - interp[3]);
}

/* resampler_basic_interpolate_single */
static int basic_interpolate(orc_resampler *st, int ch, const float *in, uint32_t *in_len, float *out,
                             uint32_t *out_len) {
	const int N = (int)st->filt_len;
	int out_sample = 0;
	int last_sample = st->last_sample[ch];
	uint32_t samp_frac_num = st->samp_frac_num[ch];
	while (!(last_sample >= (int32_t)*in_len || out_sample >= (int32_t)*out_len)) {
		const float *iptr = &in[last_sample];
		const int offset = (int)(samp_frac_num * st->oversample / st->den_rate);
		const float frac = ((float)((samp_frac_num * st->oversample) % st->den_rate)) / (float)st->den_rate;
		float interp[4];
		float accum[4] = {0, 0, 0, 0};
		for (int j = 0; j < N; j++) {
			const float curr_in = iptr[j];
			accum[0] += curr_in * st->sinc_table[4 + (j + 1) * (int)st->oversample - offset - 2];
			accum[1] += curr_in * st->sinc_table[4 + (j + 1) * (int)st->oversample - offset - 1];
			accum[2] += curr_in * st->sinc_table[4 + (j + 1) * (int)st->oversample - offset];
			accum[3] += curr_in * st->sinc_table[4 + (j + 1) * (int)st->oversample - offset + 1];
		}
		cubic_coef(frac, interp);
		out[out_sample++] = interp[0] * accum[0] + interp[1] * accum[1] + interp[2] * accum[2] + interp[3] * accum[3];
		last_sample += st->int_advance;
		samp_frac_num += (uint32_t)st->frac_advance;
		if (samp_frac_num >= st->den_rate) {
			samp_frac_num -= st->den_rate;
			last_sample++;
		}
	}
	st->last_sample[ch] = last_sample;
	st->samp_frac_num[ch] = samp_frac_num;
	return out_sample;
}

/* speex_resampler_process_native */
static void process_native(orc_resampler *st, int ch, uint32_t *in_len, float *out, uint32_t *out_len) {
	const int N = (int)st->filt_len;
	float *mem = st->mem + (size_t)ch * st->mem_alloc_size;
	int out_sample = st->use_direct ? basic_direct(st, ch, mem, in_len, out, out_len)
	                                : basic_interpolate(st, ch, mem, in_len, out, out_len);
	if (st->last_sample[ch] < (int32_t)*in_len) *in_len = (uint32_t)st->last_sample[ch];
	*out_len = (uint32_t)out_sample;
	st->last_sample[ch] -= (int32_t)*in_len;
	uint32_t ilen = *in_len;
	for (int j = 0; j < N - 1; ++j)
		mem[j] = mem[(uint32_t)j + ilen];
}

static inline int16_t word2int(float x) { /* arch.h WORD2INT, float build */
	return (int16_t)(x < -32767.5f ? -32768 : (x > 32766.5f ? 32767 : (int)floor(.5 + x)));
}

static int process_int_strided(orc_resampler *st, int ch, const int16_t *in, uint32_t *in_len, int16_t *out,
                               uint32_t *out_len, int istride, int ostride) {
	uint32_t ilen = *in_len, olen = *out_len;
	float *x = st->mem + (size_t)ch * st->mem_alloc_size;
	const uint32_t xlen = st->mem_alloc_size - (st->filt_len - 1);
	float ystack[1024];
	const uint32_t ylen = 1024;
	while (ilen && olen) {
		uint32_t ichunk = (ilen > xlen) ? xlen : ilen;
		uint32_t ochunk = (olen > ylen) ? ylen : olen;
		for (uint32_t j = 0; j < ichunk; ++j)
			x[j + st->filt_len - 1] = (float)in[j * (uint32_t)istride];
		process_native(st, ch, &ichunk, ystack, &ochunk);
		for (uint32_t j = 0; j < ochunk; ++j)
			out[j * (uint32_t)ostride] = word2int(ystack[j]);
		ilen -= ichunk;
		olen -= ochunk;
		out += ochunk * (uint32_t)ostride;
		in += ichunk * (uint32_t)istride;
	}
	*in_len -= ilen;
	*out_len -= olen;
	return 0;
}

int orc_resampler_process_int(orc_resampler *st, int channel, const int16_t *in, uint32_t *in_len, int16_t *out,
                              uint32_t *out_len) {
	return process_int_strided(st, channel, in, in_len, out, out_len, 1, 1);
}

int orc_resampler_process_interleaved_int(orc_resampler *st, const int16_t *in, uint32_t *in_len, int16_t *out,
                                          uint32_t *out_len) {
	uint32_t bak_out = *out_len, bak_in = *in_len;
	for (int i = 0; i < st->nb_channels; i++) {
		*out_len = bak_out;
		*in_len = bak_in;
		process_int_strided(st, i, in + i, in_len, out + i, out_len, st->nb_channels, st->nb_channels);
	}
	return 0;
}

int orc_msresample_block(orc_resampler *st, const int16_t *in, int in_frames, int16_t *out) { /* msresample.c:150-161 */
	uint32_t inlen = (uint32_t)in_frames;
	uint32_t outlen = (uint32_t)(((uint64_t)inlen * st->out_rate) / st->in_rate) + 1;
	if (st->nb_channels == 1) orc_resampler_process_int(st, 0, in, &inlen, out, &outlen);
	else orc_resampler_process_interleaved_int(st, in, &inlen, out, &outlen);
	return (int)outlen;
}
