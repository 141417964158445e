/* oracle/oracle_audio.c — TEST INFRASTRUCTURE (see msb200_oracle.h).
 * CPU restatements of the reference's in-tree integer / float32 audio arithmetic. Every function cites the reference
 * lines it follows; tests/test_oracle_vs_reference.py pins them against the unmodified reference (oracle/_ref). */
#include "msb200_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ mixer */
/* audiomixer.c:40-44 — note the asymmetric clamp: -32768 becomes -32767 */
static inline int16_t mix_saturate(int32_t s) {
	if (s > 32767) return 32767;
	if (s < -32767) return -32767;
	return (int16_t)s;
}

/* gained contribution of one pin: channel_process_in (:78-90) + apply_gain (:46-51).
 * Absent pin (bufferizer underrun) -> zeros (:88). Inactive pin: samples are read but neither gained nor summed
 * (:81-86), and channel_process_out does not subtract them (:124-128). */
static inline int16_t mix_contrib(int16_t s, float gain) {
	if (gain != 1.0f) return mix_saturate((int)(gain * (float)s));
	return s;
}

void orc_mixer_partial(int n_rooms, int n_pins, int nwords, const float *gain, const uint8_t *active, const int16_t *in,
                       const uint8_t *present, int32_t *sum) {
	for (int r = 0; r < n_rooms; ++r) {
		int32_t *acc = sum + (size_t)r * nwords;
		memset(acc, 0, sizeof(int32_t) * (size_t)nwords);
		for (int p = 0; p < n_pins; ++p) {
			size_t ch = (size_t)r * n_pins + p;
			const int16_t *x = in + ch * nwords;
			if (!present[ch] || !active[ch]) continue;
			for (int k = 0; k < nwords; ++k)
				acc[k] += mix_contrib(x[k], gain[ch]); /* accumulate :33-38 */
		}
	}
}

void orc_mixer_process(int n_rooms, int n_pins, int nwords, int conf_mode, const float *gain, const uint8_t *active,
                       const int16_t *in, const uint8_t *present, int16_t *out) {
	int32_t *sum = (int32_t *)malloc(sizeof(int32_t) * (size_t)nwords);
	for (int r = 0; r < n_rooms; ++r) {
		orc_mixer_partial(1, n_pins, nwords, gain + (size_t)r * n_pins, active + (size_t)r * n_pins,
		                  in + (size_t)r * n_pins * nwords, present + (size_t)r * n_pins, sum);
		if (!conf_mode) {
			int16_t *o = out + (size_t)r * nwords; /* make_output :210-217 */
			for (int k = 0; k < nwords; ++k)
				o[k] = mix_saturate(sum[k]);
			continue;
		}
		for (int p = 0; p < n_pins; ++p) { /* channel_process_out :113-130 */
			size_t ch = (size_t)r * n_pins + p;
			const int16_t *x = in + ch * nwords;
			int16_t *o = out + ch * nwords;
			if (active[ch]) {
				/* chan->input holds the post-gain block, or zeros when the pin was absent this tick */
				for (int k = 0; k < nwords; ++k) {
					int32_t own = present[ch] ? mix_contrib(x[k], gain[ch]) : 0;
					o[k] = mix_saturate(sum[k] - own);
				}
			} else {
				for (int k = 0; k < nwords; ++k)
					o[k] = mix_saturate(sum[k]);
			}
		}
	}
	free(sum);
}

/* ------------------------------------------------------------------------------------------------ volume */
static const float vol_max_e = (32768 * 0.7f); /* msvolume.c:37 */
static const float vol_coef = 0.2f;            /* :38 */

void orc_volume_init(orc_volume_state *v, int sample_rate) { /* volume_init :86-120 */
	memset(v, 0, sizeof(*v));
	v->static_gain = v->gain = v->target_gain = 1;
	v->ng_threshold = 0.1f;
	v->ng_floorgain = 0.005f;
	v->ng_gain = 1;
	v->sample_rate = sample_rate;
	v->ea_thres = 0.1f;
	v->ea_transmit_thres = 4;
	v->force = 4.0f;
	v->vol_upramp = 0.4f;
	v->sustain_time = 200;
	v->peer = -1;
}

static inline int16_t vol_saturate(int val) { /* :382-384 */
	return (int16_t)((val > 32767) ? 32767 : ((val < -32767) ? -32767 : val));
}

void orc_volume_process(orc_volume_state *v, int16_t *io, int n) {
	orc_volume_process_chunk(v, NULL, io, n);
}

void orc_volume_process_chunk(orc_volume_state *v, const float *peer_energy, int16_t *io, int n) {
	/* update_energy :388-407 */
	float acc = 0;
	int pk = 0;
	for (int i = 0; i < n; ++i) {
		int s = io[i];
		acc += (float)(s * s);
		int lp = abs(s);
		if (lp > pk) pk = lp;
	}
	float en = (float)((sqrt(acc / (float)n) + 1) / vol_max_e);
	v->energy = (en * vol_coef) + v->energy * (1.0f - vol_coef);
	v->level_pk = (float)pk / vol_max_e;
	v->instant_energy = en;
	float tgain = v->static_gain; /* volume_process :507 / :489 */
	if (peer_energy) {            /* volume_echo_avoider_process :201-238 (peer_e and peer_pk are both the peer's energy) */
		float peer_e = *peer_energy, peer_pk = *peer_energy, mic_spk_ratio;
		if (peer_pk > v->lt_speaker_en) v->lt_speaker_en = peer_pk;
		else v->lt_speaker_en = (0.005f * peer_pk) + (0.995f * v->lt_speaker_en);
		mic_spk_ratio = (v->energy / (v->lt_speaker_en + v->ea_thres));
		if (peer_e > v->ea_thres) {
			if (mic_spk_ratio > v->ea_transmit_thres) {
				v->target_gain = v->static_gain;
				v->fast_upramp = 1;
			} else {
				v->target_gain = v->static_gain / (1 + (peer_e * v->force)); /* compute_gain :186-189 */
				v->sustain_dur = v->sustain_time;
			}
		} else {
			if (v->sustain_dur > 0) v->sustain_dur -= (n * 1000) / v->sample_rate;
			else {
				v->target_gain = v->static_gain;
				v->fast_upramp = 1;
			}
		}
		tgain = v->target_gain;
	}
	if (v->agc_enabled) tgain /= (0.5f + v->level_pk) / 1; /* volume_agc_process :172-184 (non-speex build) */
	if (v->noise_gate_enabled) {  /* volume_noise_gate_process :240-260, called with instant_energy */
		float t = v->ng_floorgain;
		if (v->instant_energy > v->ng_threshold) {
			v->ng_noise_dur = 400; /* ng_cut_time :107 */
			t = 1.0f;
		} else if (v->ng_noise_dur > 0) {
			v->ng_noise_dur -= (n * 1000) / v->sample_rate;
			t = 1.0f;
		}
		v->ng_gain = v->ng_gain * 0.75f + t * 0.25f;
	}
	/* apply_gain :409-445 (vol_upramp .4, fast = 1.2, downramp .4) */
	if (v->gain < tgain) {
		if (v->gain < v->ng_floorgain) v->gain = v->ng_floorgain;
		v->gain *= 1 + (v->fast_upramp ? 0.4f * 3 : v->vol_upramp);
		if (v->gain > tgain) v->gain = tgain;
	} else if (v->gain > tgain) {
		v->gain *= 1 - 0.4f;
		if (v->gain < tgain) v->gain = tgain;
		v->fast_upramp = 0;
	}
	float gain = v->gain * v->ng_gain;
	int32_t intgain = (int32_t)(gain * 4096);
	if (v->remove_dc) {
		int dc = 0;
		for (int i = 0; i < n; ++i) {
			dc += io[i];
			io[i] = vol_saturate(((io[i] - v->dc_offset) * intgain) / 4096);
		}
		v->dc_offset = (v->dc_offset * 7 + dc * 2 / (n * 2)) / 8; /* divisor is the block size in BYTES :439 */
	} else if (gain != 1) {
		for (int i = 0; i < n; ++i)
			io[i] = vol_saturate((io[i] * intgain) / 4096);
	}
}

/* ------------------------------------------------------------------------------------------------ channel adapter */
void orc_chanadapt(int mode, int n_streams, int frames, const int16_t *in, const int16_t *in2, int16_t *out) {
	size_t total = (size_t)n_streams * frames;
	if (mode == 0) { /* chanadapt.c:112-120 */
		for (size_t i = 0; i < total; ++i)
			out[2 * i] = out[2 * i + 1] = in[i];
	} else if (mode == 1) { /* :121-129 */
		for (size_t i = 0; i < total; ++i)
			out[i] = in[2 * i];
	} else { /* :78-96, missing side zero-filled */
		for (size_t i = 0; i < total; ++i) {
			out[2 * i] = in ? in[i] : 0;
			out[2 * i + 1] = in2 ? in2[i] : 0;
		}
	}
}

/* ------------------------------------------------------------------------------------------------ equalizer */
/* dsptools.c:253-268 (float build): mem[0]=x; acc = mem[ord-1]*num[ord-1]; for j=ord-2..0 acc += num[j]*mem[j], shift */
void orc_fir_mem16(const float *x, const float *num, float *y, int N, int ord, float *mem) {
	for (int i = 0; i < N; ++i) {
		float xi = x[i];
		mem[0] = xi;
		float acc = mem[ord - 1] * num[ord - 1];
		for (int j = ord - 2; j >= 0; --j) {
			acc += num[j] * mem[j];
			mem[j + 1] = mem[j];
		}
		y[i] = acc;
	}
}

/* x86 semantics of (int16_t)float used by word16_to_int16 (equalizer.c:251-255): cvttss2si to int32, keep low 16 bits */
static inline int16_t c_cast_f32_to_s16(float f) {
	int32_t i;
	if (!(f > -2147483904.0f && f < 2147483648.0f)) i = (int32_t)0x80000000u; /* "integer indefinite" */
	else i = (int32_t)f;
	return (int16_t)(uint16_t)(uint32_t)i;
}

void orc_fir_s16(const float *taps, int ord, float *mem, int16_t *io, int n) { /* equalizer_state_run :263-269 */
	float *w = (float *)calloc((size_t)n + 1, sizeof(float));
	for (int i = 0; i < n; ++i)
		w[i] = (float)io[i];
	orc_fir_mem16(w, taps, w, n, ord, mem);
	for (int i = 0; i < n; ++i)
		io[i] = c_cast_f32_to_s16(w[i]);
	free(w);
}

static void eq_rate_update(orc_equalizer *s, int rate) { /* equalizer.c:57-79 */
	int nfft = rate < 16000 ? 128 : (rate < 32000 ? 256 : 512);
	s->rate = rate;
	s->nfft = nfft;
	free(s->fft_cpx);
	free(s->fir);
	free(s->mem);
	s->fft_cpx = (float *)calloc((size_t)nfft, sizeof(float));
	s->fir = (float *)calloc((size_t)nfft, sizeof(float));
	s->mem = (float *)calloc((size_t)nfft, sizeof(float));
	float val = 1.0f / (float)nfft; /* equalizer_state_flatten :49-55 */
	s->fft_cpx[0] = val;
	for (int i = 1; i < nfft; i += 2)
		s->fft_cpx[i] = val;
	s->needs_update = 1;
}

orc_equalizer *orc_equalizer_new(int rate) {
	orc_equalizer *s = (orc_equalizer *)calloc(1, sizeof(*s));
	eq_rate_update(s, rate);
	s->active = 1;
	return s;
}
void orc_equalizer_free(orc_equalizer *s) {
	if (!s) return;
	free(s->fft_cpx);
	free(s->fir);
	free(s->mem);
	free(s);
}
static int eq_hz_to_index(orc_equalizer *s, int hz) { /* :98-111 */
	if (hz < 0) return -1;
	if (hz > s->rate / 2) hz = s->rate / 2;
	int ret = ((hz * s->nfft) + (s->rate / 2)) / s->rate;
	if (ret == s->nfft / 2) ret = (s->nfft / 2) - 1;
	return ret;
}
static int eq_index2hz(orc_equalizer *s, int index) { /* :113-115 */
	return (index * s->rate + s->nfft / 2) / s->nfft;
}
float orc_equalizer_get_gain(orc_equalizer *s, float frequency) { /* :121-125 */
	int idx = eq_hz_to_index(s, (int)frequency);
	if (idx >= 0) return s->fft_cpx[idx * 2] * (float)s->nfft;
	return 0;
}
static float eq_gainpoint(int f, int freq_0, float sqrt_gain, int freq_bw) { /* :131-138 */
	float k1 = ((float)(f * f) - (float)(freq_0 * freq_0));
	k1 *= k1;
	float k2 = (float)(f * freq_bw);
	k2 *= k2;
	return (k1 + k2 * sqrt_gain) / (k1 + k2 / sqrt_gain);
}
static void eq_point_set(orc_equalizer *s, int i, float gain) { /* :140-148 */
	int index = 1 + ((i - 1) * 2);
	if (index >= 0 && index < s->nfft) s->fft_cpx[index] = (s->fft_cpx[index] * (float)(int)(gain * 32768)) / 32768;
}
void orc_equalizer_set_gain(orc_equalizer *s, float frequency, float gain, float width) { /* :150-177 */
	int freq_0 = (int)frequency, freq_bw = (int)width;
	int i, f;
	int delta_f = eq_index2hz(s, 1);
	float sqrt_gain = (float)sqrt(gain);
	int mid = eq_hz_to_index(s, freq_0);
	freq_bw -= delta_f / 2;
	if (freq_bw < delta_f / 2) freq_bw = delta_f / 2;
	i = mid;
	eq_point_set(s, i, gain);
	do {
		i++;
		f = eq_index2hz(s, i);
		gain = eq_gainpoint(f - delta_f, freq_0, sqrt_gain, freq_bw);
		eq_point_set(s, i, gain);
	} while (i < s->nfft / 2 && (gain > 1.1 || gain < 0.9));
	i = mid;
	do {
		i--;
		f = eq_index2hz(s, i);
		gain = eq_gainpoint(f + delta_f, freq_0, sqrt_gain, freq_bw);
		eq_point_set(s, i, gain);
	} while (i >= 0 && (gain > 1.1 || gain < 0.9));
	s->needs_update = 1;
}

/* ms_ifft (dsptools.c:373-376 -> kiss_fftri2, kiss_fftr.c:261-296): unnormalised inverse of the packed real spectrum
 * [r0, r1, i1, ..., r(n/2-1), i(n/2-1), r(n/2)] through the bit-exact restatement of the float kiss_fft in oracle_plc.c:
 * the taps, and with them every output sample, equal the reference's bit for bit. */
static void eq_packed_irfft(const float *spec, float *out, int n) {
	if (orc_kiss_irfft(spec, out, n) != 0) memset(out, 0, sizeof(float) * (size_t)n);
}
static void eq_compute_impulse_response(orc_equalizer *s) { /* :215-237 */
	int n = s->nfft, half = n / 2;
	eq_packed_irfft(s->fft_cpx, s->fir, n);
	for (int i = 0; i < half; ++i) { /* time_shift :184-193 */
		float tmp = s->fir[i];
		s->fir[i] = s->fir[i + half];
		s->fir[i + half] = tmp;
	}
	for (int i = 0; i < n; ++i) { /* norm_and_apodize :203-213 (Hamming) */
		float x = (float)((float)i * 2 * M_PI / (float)n);
		float w = (float)(0.54 - (0.46 * cos(x)));
		s->fir[i] = w * s->fir[i];
	}
	s->needs_update = 0;
}
const float *orc_equalizer_taps(orc_equalizer *s) {
	if (s->needs_update) eq_compute_impulse_response(s);
	return s->fir;
}
void orc_equalizer_process(orc_equalizer *s, int16_t *io, int n) { /* equalizer_process :279-288 */
	if (!s->active) return;
	if (s->needs_update) eq_compute_impulse_response(s);
	orc_fir_s16(s->fir, s->nfft, s->mem, io, n);
}

/* ---------------------------------------------------------------------------------------------------- MSAudioFlowControl
 * ms_audio_flow_controller_process() /root/reference/src/audiofilters/flowcontrol.c:110-150 with its helpers
 * discard_well_choosed_samples :58-92 (three-sample criterion, ties resolved towards the LAST position: `<=`) and
 * compute_frame_power :100-108 (sequential float32 sum, sqrtf). Restated iteratively. Pinned bit-exact against the
 * unmodified reference filter in an MSTicker (tests/test_oracle_vs_reference.py::test_flowcontrol_*). */
void orc_flowctl_init(orc_flowctl *c) {
	c->strategy = 1; /* MSAudioFlowControlSoft */
	c->silent_threshold = 0.02f;
	c->target_samples = c->total_samples = c->current_pos = c->current_dropped = 0;
}
void orc_flowctl_set_target(orc_flowctl *c, uint32_t samples_to_drop, uint32_t total_samples) {
	c->target_samples = samples_to_drop;
	c->total_samples = total_samples;
	c->current_pos = 0;
	c->current_dropped = 0;
}
/* processes one block in place; returns the number of samples that remain (0: the block is dropped) */
int orc_flowctl_process(orc_flowctl *c, int16_t *s, int nsamples) {
	uint32_t n = (uint32_t)nsamples;
	if (!(c->total_samples > 0 && c->target_samples > 0)) return nsamples;
	c->current_pos += n;
	if (c->strategy == 0) { /* basic: whole blocks while they fit in the target */
		if (c->current_dropped + n <= c->target_samples) {
			c->current_dropped += n;
			n = 0;
		}
	} else {
		const uint32_t th = (uint32_t)(((uint64_t)c->target_samples * (uint64_t)c->current_pos) / (uint64_t)c->total_samples);
		uint32_t todrop = th > c->current_dropped ? th - c->current_dropped : 0;
		if (todrop > 0) {
			int silent = 0;
			if (n <= c->target_samples) {
				float acc = 0;
				for (uint32_t i = 0; i < n; ++i) {
					const int v = s[i];
					acc += (float)(v * v);
				}
				silent = sqrtf(acc / (float)n) / (32768 * 0.7f) < c->silent_threshold;
			}
			if (silent) {
				todrop = n;
				n = 0;
			} else if (todrop * 8 < n) {
				for (uint32_t d = 0; d < todrop; ++d) { /* remove the middle sample of the flattest three-sample run */
					int best = 32768;
					uint32_t pos = 0;
					for (uint32_t i = 0; i + 2 < n; ++i) {
						const int t = abs((int)s[i] - (int)s[i + 1]) + abs((int)s[i + 1] - (int)s[i + 2]);
						if (t <= best) {
							best = t;
							pos = i;
						}
					}
					memmove(s + pos + 1, s + pos + 2, (size_t)(n - pos - 2) * sizeof(int16_t));
					--n;
				}
			} else {
				todrop = n;
				n = 0;
			}
			c->current_dropped += todrop;
		}
	}
	if (c->current_pos >= c->total_samples) c->target_samples = 0;
	return (int)n;
}
