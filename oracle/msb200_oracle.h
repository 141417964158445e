/* oracle/msb200_oracle.h — CPU restatements of the reference's per-tick DSP arithmetic.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (libmsb200dsp.so) never links, dlopens or calls anything declared here.
 *
 * Provenance of each restatement (see the .c files for file:line citations):
 *   mixer / volume / chanadapt / equalizer FIR / NV12->I420 : restated from the in-tree reference C and PINNED against
 *       the unmodified reference compiled into oracle/_ref/libms2ref.so (tests/test_oracle_vs_reference.py).
 *   resampler, MDF echo canceller, preprocessor : restated from the published speexdsp 1.2 algorithm (speexdsp is an
 *       external, un-vendored dependency: /root/reference/CMakeLists.txt:208). PARITY UNPINNED: the reference's own
 *       tests hold no sample-level vectors for these; only independent cross-checks exist.
 *   pixel conversion / bilinear scaling : restated from swscale's SWS_BILINEAR pipeline, PINNED against golden frames
 *       generated here from libswscale 9.1.100 (tests/golden/, tests/golden/make_swscale_golden.py).
 */
#ifndef MSB200_ORACLE_H
#define MSB200_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- mixer (audiomixer.c:33-51, 78-90, 113-130, 210-217) */
void orc_mixer_process(int n_rooms, int n_pins, int nwords, int conf_mode, const float *gain, const uint8_t *active,
                       const int16_t *in, const uint8_t *present, int16_t *out);
void orc_mixer_partial(int n_rooms, int n_pins, int nwords, const float *gain, const uint8_t *active, const int16_t *in,
                       const uint8_t *present, int32_t *sum);

/* ---- volume (msvolume.c:240-260, 388-445, 503-513); layout matches msb200_volume_state */
typedef struct orc_volume_state {
	float energy, level_pk, instant_energy, gain, static_gain, target_gain, ng_gain, ng_threshold, ng_floorgain;
	int32_t dc_offset, ng_noise_dur, noise_gate_enabled, remove_dc, sample_rate, fast_upramp;
	float lt_speaker_en, ea_thres, ea_transmit_thres, force, vol_upramp;
	int32_t sustain_time, sustain_dur, agc_enabled, peer;
} orc_volume_state;
void orc_volume_init(orc_volume_state *v, int sample_rate);
void orc_volume_process(orc_volume_state *v, int16_t *io, int nsamples);
/* chunked mode (msvolume.c:480-502): peer_energy = the peer MSVolume's smoothed energy, or NULL when there is no peer */
void orc_volume_process_chunk(orc_volume_state *v, const float *peer_energy, int16_t *io, int nsamples);

/* ---- channel adapter (chanadapt.c:68-131) */
void orc_chanadapt(int mode, int n_streams, int frames, const int16_t *in, const int16_t *in2, int16_t *out);

/* ---- equalizer (equalizer.c:147-172, 215-237, 263-269; dsptools.c:253-268) */
typedef struct orc_equalizer {
	int rate, nfft, needs_update, active;
	float *fft_cpx, *fir, *mem;
} orc_equalizer;
orc_equalizer *orc_equalizer_new(int rate);
void orc_equalizer_free(orc_equalizer *e);
void orc_equalizer_set_gain(orc_equalizer *e, float frequency, float gain, float width);
float orc_equalizer_get_gain(orc_equalizer *e, float frequency);
const float *orc_equalizer_taps(orc_equalizer *e);
void orc_equalizer_process(orc_equalizer *e, int16_t *io, int nsamples);
void orc_fir_mem16(const float *x, const float *num, float *y, int N, int ord, float *mem);
void orc_fir_s16(const float *taps, int ord, float *mem, int16_t *io, int nsamples);

/* ---- resampler (speexdsp resample.c, quality 3, float build) */
typedef struct orc_resampler orc_resampler;
orc_resampler *orc_resampler_new(int nch, int in_rate, int out_rate, int quality);
void orc_resampler_free(orc_resampler *r);
/* mirrors speex_resampler_process_int(st, ch, in, &in_len, out, &out_len) with st->in_stride = st->out_stride = nch */
int orc_resampler_process_int(orc_resampler *r, int channel, const int16_t *in, uint32_t *in_len, int16_t *out,
                              uint32_t *out_len);
int orc_resampler_process_interleaved_int(orc_resampler *r, const int16_t *in, uint32_t *in_len, int16_t *out,
                                          uint32_t *out_len);
int orc_resampler_filt_len(orc_resampler *r);
int orc_resampler_den(orc_resampler *r);
int orc_resampler_num(orc_resampler *r);
int orc_resampler_use_direct(orc_resampler *r);
int orc_resampler_oversample(orc_resampler *r);
const float *orc_resampler_table(orc_resampler *r, int *len);
/* the msresample.c:150-169 wrapper: outlen = inlen*out/in + 1, returns frames produced */
int orc_msresample_block(orc_resampler *r, const int16_t *in, int in_frames, int16_t *out);

/* ---- echo canceller + preprocessor (speexdsp mdf.c / preprocess.c / filterbank.c, float build, TWO_PATH) */
typedef struct orc_aec orc_aec;
int orc_aec_frame_size_for_rate(int sample_rate, int framesize_at_8000);
orc_aec *orc_aec_new(int sample_rate, int tail_length_ms, int framesize_at_8000);
void orc_aec_free(orc_aec *a);
int orc_aec_frame_size(orc_aec *a);
int orc_aec_M(orc_aec *a);
/* one frame: speex_echo_cancellation(mic, ref, out) ; speex_preprocess_run(out) */
void orc_aec_process_frame(orc_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out);
/* echo canceller only / preprocessor only, for staged parity tests */
void orc_aec_cancel_frame(orc_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out);
void orc_aec_preprocess_frame(orc_aec *a, int16_t *io);
int orc_aec_probe(orc_aec *a, const char *what, float *out, int max_floats);

/* ---- video */
int orc_nv12_to_i420(const uint8_t *y, const uint8_t *cbcr, int rotation, int w, int h, int y_stride, int cbcr_stride,
                     int u_first, int down_scale, uint8_t *out);
typedef struct orc_scaler orc_scaler;
orc_scaler *orc_scaler_new(int src_w, int src_h, int src_fmt, int dst_w, int dst_h, int dst_fmt);
void orc_scaler_free(orc_scaler *s);
size_t orc_scaler_src_bytes(orc_scaler *s);
size_t orc_scaler_dst_bytes(orc_scaler *s);
int orc_scaler_process(orc_scaler *s, const uint8_t *src, uint8_t *dst);
/* planar output only: 1 = round like the library's x86 SIMD vertical scaler (what a plain SWS_BILINEAR call returns on x86),
 * 0 (default) = its C reference arithmetic (SWS_BITEXACT), which is what the GPU kernels compute this round */
void orc_scaler_set_x86_vertical(orc_scaler *s, int on);
/* filter tables (shared verbatim with the product through tests only): returns filter size, fills pos/coef */
int orc_scaler_get_filter(orc_scaler *s, int which /*0 lumH,1 chrH,2 lumV,3 chrV*/, int32_t *pos, int16_t *coef,
                          int max_entries);

/* MSAudioFlowControl (oracle_audio.c) */
typedef struct orc_flowctl {
	int32_t strategy; /* 0 basic, 1 soft */
	float silent_threshold;
	uint32_t target_samples, total_samples, current_pos, current_dropped;
} orc_flowctl;
void orc_flowctl_init(orc_flowctl *c);
void orc_flowctl_set_target(orc_flowctl *c, uint32_t samples_to_drop, uint32_t total_samples);
int orc_flowctl_process(orc_flowctl *c, int16_t *samples, int nsamples);

/* ms_ifft / ms_fft (dsptools.c:362-376 over the float kiss_fft), restated bit-exactly in oracle_plc.c; 0 on success */
int orc_kiss_irfft(const float *spec, float *out, int nfft);
int orc_kiss_rfft(const float *time, float *spec, int nfft);

/* MSGenericPLC (oracle_plc.c): signal level (packet / conceal) and filter level (concealer clock) */
typedef struct orc_plc orc_plc;
int orc_plc_rate_supported(int rate);
orc_plc *orc_plc_create(int rate);
void orc_plc_destroy(orc_plc *c);
int orc_plc_history_len(const orc_plc *c);
void orc_plc_packet(orc_plc *c, int16_t *data, int n, int after_cng);
void orc_plc_conceal(orc_plc *c, int16_t *data, int n);
void orc_plc_filter_set_cn(orc_plc *c);
void orc_plc_filter_packet(orc_plc *c, uint64_t now_ms, int16_t *data, int n, int nchannels);
int orc_plc_filter_tick(orc_plc *c, uint64_t now_ms, int interval_ms, int nchannels, int16_t *out, int *kind);

/* G.711 (oracle_g711.c): law 0 = A-law, 1 = mu-law */
void orc_g711_encode(int law, const int16_t *pcm, uint8_t *code, size_t n);
void orc_g711_decode(int law, const uint8_t *code, int16_t *pcm, size_t n);

#ifdef __cplusplus
}
#endif
#endif
