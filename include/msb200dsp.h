/*
 * msb200dsp.h — C ABI of libmsb200dsp.so: mediastreamer2's per-tick DSP hot path on NVIDIA B200 (sm_100a).
 *
 * Plain C: pointers and sizes only, no CUDA / torch types. Device pointers cross the ABI as `void *`.
 * Every object is a BANK: the same reference filter instantiated for `n` independent call streams (or rooms, or
 * video streams) whose per-stream state lives in HBM in structure-of-arrays layout. One call = one MSTicker tick
 * (or one block / frame) of every stream in the bank = one or two kernel launches.
 *
 * The reference calls `MSFilterDesc.process(MSFilter*)` once per stream per tick on the ticker thread
 * (/root/reference/src/base/msticker.c:244-259). Each `msb200_<filter>_process*` below is the batched equivalent of
 * the cited process() body; `plugin/` wraps them back into MSFilterDesc objects with the reference's ids, names and
 * method tables.
 *
 * Error convention: functions return 0 on success, a negative MSB200_E* code otherwise; msb200_last_error() gives a
 * thread-local message. There is NO CPU fallback: without a CUDA device every create/process call fails with
 * MSB200_ENODEV.
 *
 * Two flavours of every process call:
 *   _process      host buffers (pageable or pinned): H2D copy + kernel(s) + D2H copy + stream sync ("e2e" path,
 *                 what an MSFilter.process() wrapper calls)
 *   _process_dev  device buffers, asynchronous on the context's stream (chaining filters without leaving HBM)
 */
#ifndef MSB200DSP_H
#define MSB200DSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSB200_API __attribute__((visibility("default")))

#define MSB200_OK 0
#define MSB200_EINVAL (-1) /* bad argument (the reference's methods return -1 likewise, src/base/msfilter.c:196) */
#define MSB200_ENODEV (-2) /* no CUDA device / driver: nothing runs, by design */
#define MSB200_ECUDA (-3)  /* a CUDA runtime call failed; see msb200_last_error() */
#define MSB200_ENOMEM (-4)
#define MSB200_ESTATE (-5) /* call not valid in the object's current state */

typedef struct msb200_ctx msb200_ctx;

/* ---------------------------------------------------------------------------------------------------- context */
MSB200_API int msb200_version(void);
MSB200_API const char *msb200_last_error(void);
/* One context per (process, GPU): owns a CUDA stream, timing events and a pinned staging arena. */
MSB200_API int msb200_ctx_create(int device_ordinal, msb200_ctx **out);
/* Same, but every launch goes to a stream owned by the caller (a cudaStream_t passed as void*; NULL = the legacy default
 * stream). Lets a host order our kernels with its own work on that stream — e.g. NCCL collectives issued by
 * torch.distributed between msb200_mixer_partial_dev and msb200_mixer_finish_dev — without host synchronisation. */
MSB200_API int msb200_ctx_create_on_stream(int device_ordinal, void *cuda_stream, msb200_ctx **out);
MSB200_API void msb200_ctx_destroy(msb200_ctx *ctx);
MSB200_API int msb200_ctx_sync(msb200_ctx *ctx);
/* Deferred synchronisation for hosts that flush several banks per tick (the plugin's lockstep groups of one MSTicker): while
 * on, the host-buffer entry points msb200_resample_process, msb200_aec_process_counts, msb200_volume_process_blocks,
 * msb200_mixer_process, msb200_plc_process_strided and msb200_g711_decode / _encode return as soon as their copies and
 * kernels are enqueued on the context's stream; their outputs are valid after the next msb200_ctx_sync(). The buffers must
 * be pinned (msb200_host_alloc_pinned) and stay untouched until then. Turning it off synchronises. */
MSB200_API int msb200_ctx_set_deferred_sync(msb200_ctx *ctx, int on);
/* cudaSetDevice(ctx's device) for the calling thread: call it first on every thread that uses ctx (or its banks) when
 * the process drives more than one GPU; a no-op cost otherwise */
MSB200_API int msb200_ctx_make_current(msb200_ctx *ctx);
/* Kernel launches issued through this context since creation (bench.py reports it as gpu_launches). */
MSB200_API uint64_t msb200_ctx_launch_count(msb200_ctx *ctx);
/* Raw device memory helpers so that a plain-C (or ctypes) caller can keep buffers resident in HBM. */
MSB200_API int msb200_dev_alloc(msb200_ctx *ctx, size_t bytes, void **dev_ptr);
MSB200_API int msb200_dev_free(msb200_ctx *ctx, void *dev_ptr);
MSB200_API int msb200_host_alloc_pinned(msb200_ctx *ctx, size_t bytes, void **host_ptr);
MSB200_API int msb200_host_free_pinned(msb200_ctx *ctx, void *host_ptr);
MSB200_API int msb200_memcpy_h2d(msb200_ctx *ctx, void *dev, const void *host, size_t bytes);
MSB200_API int msb200_memcpy_d2h(msb200_ctx *ctx, void *host, const void *dev, size_t bytes);
MSB200_API int msb200_memset_dev(msb200_ctx *ctx, void *dev, int value, size_t bytes);
/* Device-side stopwatch on the context's stream (CUDA events): start, run work, stop -> milliseconds. */
MSB200_API int msb200_timer_start(msb200_ctx *ctx);
MSB200_API int msb200_timer_stop_ms(msb200_ctx *ctx, float *ms);
/* Write `bytes` of scratch (> L2) so the next timed launch starts from HBM, not from the 126 MB L2. */
MSB200_API int msb200_flush_l2(msb200_ctx *ctx);

/* ---------------------------------------------------------------------------------------------------- MSAudioMixer
 * Replaces mixer_process() /root/reference/src/audiofilters/audiomixer.c:288-346 (accumulate :33-38, saturate
 * :40-44, apply_gain :46-51, channel_process_in :78-90, channel_process_out :113-130) for `n_rooms` mixers of
 * `n_pins` pins each. The per-pin bufferizer / bypass / flow-control decisions (:92-111, :244-286) depend on
 * ticker->time and stay on the host (plugin/); what arrives here is, per pin, either one tick of samples
 * (present=1) or nothing (present=0 -> contributes zeros, exactly as :88).
 * Layouts: in  [room][pin][nwords] s16;  present [room][pin] u8;
 *          out conference mode: [room][pin][nwords] s16 = sat(sum - own) (own = post-gain input if active and present)
 *              plain mode:      [room][nwords] s16 = sat(sum)
 * Bit-exact with the reference, including saturation to [-32767, 32767] and `(int)(gain*(float)s)` gain. */
typedef struct msb200_mixer msb200_mixer;
MSB200_API int msb200_mixer_create(msb200_ctx *ctx, int n_rooms, int n_pins, int nwords, int conf_mode,
                                   msb200_mixer **out);
MSB200_API void msb200_mixer_destroy(msb200_mixer *m);
MSB200_API int msb200_mixer_set_input_gain(msb200_mixer *m, int room, int pin, float gain); /* MS_AUDIO_MIXER_SET_INPUT_GAIN */
MSB200_API int msb200_mixer_set_active(msb200_mixer *m, int room, int pin, int active);     /* MS_AUDIO_MIXER_SET_ACTIVE */
MSB200_API int msb200_mixer_set_live(msb200_mixer *m, int n_live_rooms); /* see msb200_volume_set_live */
MSB200_API int msb200_mixer_process(msb200_mixer *m, const int16_t *in, const uint8_t *present, int16_t *out);
MSB200_API int msb200_mixer_process_dev(msb200_mixer *m, const void *d_in, const void *d_present, void *d_out);
/* Cross-GPU conference: phase 1 writes each room's int32 partial sum of the LOCAL pins to d_sum
 * [room][nwords] (the caller all-reduces it with ncclSum/ncclInt32 — integer, hence order-independent and
 * bit-exact), phase 2 emits the local pins' outputs from the reduced sum. */
MSB200_API int msb200_mixer_partial_dev(msb200_mixer *m, const void *d_in, const void *d_present, void *d_sum_i32);
MSB200_API int msb200_mixer_finish_dev(msb200_mixer *m, const void *d_in, const void *d_present, const void *d_sum_i32,
                                        void *d_out);
/* Device-memory sharing between the per-GPU processes of one node (cudaIpc): export a cudaMalloc'ed pointer as 64 opaque
 * bytes, import it in another process of the node, close it before the exporter frees the memory. */
#define MSB200_IPC_HANDLE_BYTES 64
#define MSB200_MAX_PEERS 8
MSB200_API int msb200_ipc_export(msb200_ctx *ctx, void *dev_ptr, uint8_t handle[MSB200_IPC_HANDLE_BYTES]);
MSB200_API int msb200_ipc_import(msb200_ctx *ctx, const uint8_t handle[MSB200_IPC_HANDLE_BYTES], void **dev_ptr);
MSB200_API int msb200_ipc_close(msb200_ctx *ctx, void *dev_ptr);

/* ---- the cross-GPU conference exchange behind the C ABI, two ways (csrc/mixer_xchg.cu) -------------------------------
 * Striped layout (BASELINE cfg3): rank r of `world` owns the pins {r, r+world, ...} of EVERY room; its msb200_mixer bank is
 * created with n_pins = the local pin count and conference mode on. The reference semantics are those of ONE mixer with
 * all the pins: sum :288-346, out_i = sat(sum - own_i) :113-130, +-32767 :40-44 — integer, so any reduction order is
 * bit-exact.
 *
 * (1) NCCL: a communicator bound to the context's stream. libnccl.so.2 is dlopen'ed on first use (search order: the
 *     MSB200_NCCL_LIB environment variable, a copy the process already loaded, the default library path); one rank
 *     calls msb200_comm_unique_id() and hands the 128 bytes to the others through any side channel.
 *     msb200_mixer_process_striped_dev = partial_dev -> ncclAllReduce(int32, SUM) in place -> finish_dev: 3 launches. */
#define MSB200_COMM_ID_BYTES 128
typedef struct msb200_comm msb200_comm;
MSB200_API int msb200_comm_available(void);    /* 1 when an NCCL library could be loaded */
MSB200_API int msb200_comm_nccl_version(void); /* e.g. 22809, 0 when unavailable */
MSB200_API int msb200_comm_unique_id(uint8_t id[MSB200_COMM_ID_BYTES]);
MSB200_API int msb200_comm_create(msb200_ctx *ctx, const uint8_t id[MSB200_COMM_ID_BYTES], int rank, int world,
                                  msb200_comm **out); /* collective: every rank of `world` calls it */
MSB200_API void msb200_comm_destroy(msb200_comm *c);
MSB200_API int msb200_comm_allreduce_sum_i32_dev(msb200_comm *c, void *d_buf_i32, size_t count);
MSB200_API int msb200_mixer_process_striped_dev(msb200_mixer *m, msb200_comm *c, const void *d_in, const void *d_present,
                                                 void *d_sum_i32, void *d_out);
/* (2) Fused over NVLink peer memory: ONE kernel per tick and rank, no NCCL on the data path. Each CTA pushes its int32
 *     partial sums into a receive slot of every rank (posted 16-byte stores), publishes a per-CTA epoch flag on every rank,
 *     waits on its LOCAL flags for the same CTA of every peer, then adds the world's slots from local memory and emits the
 *     local pins' outputs. Set-up: create on every rank; exchange the export() handles (ranks in different processes:
 *     connect(handles[world][64]); ranks that are contexts of ONE process, e.g. a host driving several GPUs:
 *     connect_local(all[world])); barrier; then process_dev once per tick on every rank, in the same order of ticks.
 *     A peer that stays silent for MSB200_XCHG_TIMEOUT_MS (default 2000) is counted in status() instead of hanging the
 *     GPU; that tick's outputs are then undefined. Barrier again before destroy (peers may still be reading). */
typedef struct msb200_mixer_xchg msb200_mixer_xchg;
MSB200_API int msb200_mixer_xchg_create(msb200_mixer *m, int rank, int world, msb200_mixer_xchg **out);
MSB200_API int msb200_mixer_xchg_export(msb200_mixer_xchg *x, uint8_t handle[MSB200_IPC_HANDLE_BYTES]);
MSB200_API int msb200_mixer_xchg_connect(msb200_mixer_xchg *x, const uint8_t *handles);
MSB200_API int msb200_mixer_xchg_connect_local(msb200_mixer_xchg *x, msb200_mixer_xchg *const *all);
MSB200_API int msb200_mixer_xchg_process_dev(msb200_mixer_xchg *x, const void *d_in, const void *d_present, void *d_out);
MSB200_API int msb200_mixer_xchg_status(msb200_mixer_xchg *x, uint32_t *timeouts);
MSB200_API size_t msb200_mixer_xchg_wire_bytes_per_tick(msb200_mixer_xchg *x); /* (world - 1) x rooms x nwords x 4 pushed */
MSB200_API void msb200_mixer_xchg_destroy(msb200_mixer_xchg *x);

/* ---------------------------------------------------------------------------------------------------- MSVolume
 * Replaces the light path of volume_process() /root/reference/src/audiofilters/msvolume.c:503-513:
 * update_energy :388-407, volume_noise_gate_process :240-260, apply_gain :409-445 (Q12 integer gain, truncating
 * division, saturation to +-32767, optional DC removal). One block of `nsamples` per stream per call, in place.
 * The float state machine is evaluated in the reference's operation order (no FMA contraction): bit-exact. */
typedef struct msb200_volume msb200_volume;
typedef struct msb200_volume_state {
	float energy, level_pk, instant_energy, gain, static_gain, target_gain, ng_gain, ng_threshold, ng_floorgain;
	int32_t dc_offset, ng_noise_dur, noise_gate_enabled, remove_dc, sample_rate, fast_upramp;
	/* chunked mode (AGC and/or echo-limiter peer, msvolume.c:480-502) */
	float lt_speaker_en, ea_thres, ea_transmit_thres, force, vol_upramp;
	int32_t sustain_time, sustain_dur, agc_enabled, peer; /* peer: stream index in the peer bank, -1 = none */
} msb200_volume_state;
MSB200_API int msb200_volume_create(msb200_ctx *ctx, int n_streams, int sample_rate, int max_block, msb200_volume **out);
MSB200_API void msb200_volume_destroy(msb200_volume *v);
MSB200_API int msb200_volume_reset_stream(msb200_volume *v, int stream); /* volume_init state, one stream */
/* Partially occupied banks (all four audio banks have this call): only streams (rooms) [0, n_live) are copied and
 * processed by the next process calls; the others keep their state and their host rows are not touched. */
MSB200_API int msb200_volume_set_live(msb200_volume *v, int n_live);
/* Which kernel serves the bank: 0 (default) = by size — one warp per stream for small banks (lowest latency), one LANE per
 * stream for the sequential energy sum from 256 live streams on (a warp per stream spends 31 of 32 issue slots idle there);
 * 1 / 2 force either (tests, A/B). Same bytes and state either way. */
MSB200_API int msb200_volume_set_kernel(msb200_volume *v, int choice);
MSB200_API int msb200_volume_set_gain(msb200_volume *v, int stream, float gain);           /* MS_VOLUME_SET_GAIN :270-276 */
MSB200_API int msb200_volume_set_db_gain(msb200_volume *v, int stream, float db);          /* MS_VOLUME_SET_DB_GAIN :262-268 */
MSB200_API int msb200_volume_enable_noise_gate(msb200_volume *v, int stream, int enabled); /* :352-359 */
MSB200_API int msb200_volume_set_noise_gate_threshold(msb200_volume *v, int stream, float thr);
MSB200_API int msb200_volume_set_noise_gate_floorgain(msb200_volume *v, int stream, float g);
MSB200_API int msb200_volume_remove_dc(msb200_volume *v, int stream, int enabled);
/* Chunked mode (msvolume.c:480-502): with AGC or an echo-limiter peer the reference re-frames its input to 10 ms chunks
 * and, per chunk, runs update_energy, volume_echo_avoider_process (:201-238, reads the PEER filter's smoothed energy),
 * volume_agc_process (:172-184), the noise gate and apply_gain. The peer is a stream of another (or the same) bank whose
 * state is read when this bank is processed: process the peer bank first, as the ticker does for volrecv -> volsend. */
MSB200_API int msb200_volume_enable_agc(msb200_volume *v, int stream, int enabled);                       /* MS_VOLUME_ENABLE_AGC */
MSB200_API int msb200_volume_set_peer(msb200_volume *v, int stream, msb200_volume *peer_bank, int peer_stream); /* MS_VOLUME_SET_PEER */
MSB200_API int msb200_volume_set_ea_threshold(msb200_volume *v, int stream, float thr);                   /* :305-314 */
MSB200_API int msb200_volume_set_ea_speed(msb200_volume *v, int stream, float speed);                     /* :324-333 */
MSB200_API int msb200_volume_set_ea_force(msb200_volume *v, int stream, float force);
MSB200_API int msb200_volume_set_ea_sustain(msb200_volume *v, int stream, int ms);
MSB200_API int msb200_volume_set_ea_transmit_threshold(msb200_volume *v, int stream, float thr);
MSB200_API int msb200_volume_get_state(msb200_volume *v, int stream, msb200_volume_state *st);
/* io: [stream][nsamples] s16, processed in place */
MSB200_API int msb200_volume_process(msb200_volume *v, int16_t *io, int nsamples);
MSB200_API int msb200_volume_process_dev(msb200_volume *v, void *d_io, int nsamples, int stride_samples);
/* up to `nblocks` consecutive blocks of `nsamples` per stream in one launch (the reference's per-mblk loop :505-512),
 * io laid out [stream][stride_samples]; counts (NULL = nblocks for all) gives each stream's own block count: a stream
 * that staged fewer blocks in this tick is left untouched beyond its count (the plugin's lockstep batch mode) */
MSB200_API int msb200_volume_process_blocks(msb200_volume *v, int16_t *io, int nsamples, int stride_samples, int nblocks,
                                            const int32_t *counts);

/* ---------------------------------------------------------------------------------------------------- MSChannelAdapter
 * Replaces adapter_process() /root/reference/src/audiofilters/chanadapt.c:99-132 (mono->stereo duplicate,
 * stereo->mono take-left) and the two-mono-pins -> interleaved-stereo path :68-97. Stateless, bit-exact. */
#define MSB200_CHAN_MONO_TO_STEREO 0
#define MSB200_CHAN_STEREO_TO_MONO 1
#define MSB200_CHAN_2MONO_TO_STEREO 2
/* mode 0: in [n][frames] -> out [n][frames][2];  mode 1: in [n][frames][2] -> out [n][frames];
 * mode 2: in = pin0 [n][frames], in2 = pin1 [n][frames] (NULL -> zeros) -> out [n][frames][2] */
MSB200_API int msb200_chanadapt_process(msb200_ctx *ctx, int mode, int n_streams, int frames, const int16_t *in,
                                        const int16_t *in2, int16_t *out);
MSB200_API int msb200_chanadapt_process_dev(msb200_ctx *ctx, int mode, int n_streams, int frames, const void *d_in,
                                            const void *d_in2, void *d_out);

/* ---------------------------------------------------------------------------------------------------- MSEqualizer
 * Replaces equalizer_process() /root/reference/src/audiofilters/equalizer.c:279-288 -> ms_fir_mem16()
 * /root/reference/src/utils/dsptools.c:253-268 (float build): nfft-tap direct-form FIR with a per-stream delay line,
 * accumulated from tap ord-1 down to 0 in float32 without FMA, output cast to s16 by C truncation — bit-exact given
 * equal taps. The taps (gain table -> IFFT -> time shift -> Hamming, equalizer.c:215-237) are computed on the host by
 * msb200_equalizer_set_gain(), mirroring MS_EQUALIZER_SET_GAIN (:147-172). */
typedef struct msb200_equalizer msb200_equalizer;
MSB200_API int msb200_equalizer_create(msb200_ctx *ctx, int n_streams, int sample_rate, int max_block,
                                       msb200_equalizer **out);
MSB200_API void msb200_equalizer_destroy(msb200_equalizer *e);
MSB200_API int msb200_equalizer_nfft(msb200_equalizer *e);
MSB200_API int msb200_equalizer_set_gain(msb200_equalizer *e, int stream, float frequency, float gain, float width);
MSB200_API int msb200_equalizer_get_gain(msb200_equalizer *e, int stream, float frequency, float *gain);
MSB200_API int msb200_equalizer_set_active(msb200_equalizer *e, int stream, int active);
/* direct tap access (tests, state restore): taps [nfft] float */
/* the design step alone (equalizer_state_compute_impulse_response, equalizer.c:215-237: ms_ifft of the packed gain table,
 * time shift, Hamming window), on the host, no device needed: nfft in {128, 256, 512}; taps equal the reference's bit for bit */
MSB200_API int msb200_equalizer_design(int nfft, const float *gain_table, float *taps);
MSB200_API int msb200_equalizer_set_taps(msb200_equalizer *e, int stream, const float *taps);
MSB200_API int msb200_equalizer_get_taps(msb200_equalizer *e, int stream, float *taps);
MSB200_API int msb200_equalizer_process(msb200_equalizer *e, int16_t *io, int nsamples);
MSB200_API int msb200_equalizer_process_dev(msb200_equalizer *e, void *d_io, int nsamples, int stride_samples);

/* ---------------------------------------------------------------------------------------------------- MSResample
 * Replaces resample_process_ms2() /root/reference/src/audiofilters/msresample.c:122-179, i.e. the call
 * speex_resampler_process_int(handle, 0, in, &inlen, out, &outlen) (:157) with quality
 * SPEEX_RESAMPLER_QUALITY_VOIP = 3 (:104). speexdsp is NOT under /root/reference; the algorithm (Kaiser-windowed sinc
 * polyphase, direct table when den<=8 else cubic-interpolated oversampled table) is restated in oracle/oracle_resample.c.
 * All streams of a bank share rates and are fed the same number of frames per call; per-stream filter phase is kept.
 * out must hold msb200_resample_max_out(r, in_frames) frames per stream; *out_frames receives the produced count
 * (identical for all streams fed in lockstep). Mono or interleaved multi-channel (channel 0..nch-1 independent). */
typedef struct msb200_resample msb200_resample;
MSB200_API int msb200_resample_create(msb200_ctx *ctx, int n_streams, int in_rate, int out_rate, int nchannels,
                                      int max_in_frames, msb200_resample **out);
MSB200_API void msb200_resample_destroy(msb200_resample *r);
MSB200_API int msb200_resample_max_out(msb200_resample *r, int in_frames); /* inlen*out/in + 1, msresample.c:151-152 */
MSB200_API int msb200_resample_reset(msb200_resample *r);
MSB200_API int msb200_resample_set_live(msb200_resample *r, int n_live); /* see msb200_volume_set_live */
/* one stream's filter memory only (a stream joining a running lockstep bank; the resampling phase stays bank-wide) */
MSB200_API int msb200_resample_reset_stream(msb200_resample *r, int stream);
MSB200_API int msb200_resample_process(msb200_resample *r, const int16_t *in, int in_frames, int16_t *out,
                                       int out_stride_frames, int *out_frames);
MSB200_API int msb200_resample_process_dev(msb200_resample *r, const void *d_in, int in_frames, int in_stride_frames,
                                           void *d_out, int out_stride_frames, int *out_frames);

/* ---------------------------------------------------------------------------------------------------- MSSpeexEC
 * Replaces the arithmetic of speex_ec_process() /root/reference/src/audiofilters/speexec.c:223-305:
 * per `framesize` block  speex_echo_cancellation(ec, mic, ref, out) + speex_preprocess_run(den, out)  (:297-298)
 * configured as speex_ec_preprocess() does (:188-216): framesize = largest pow2 <= 64*rate/8000 (:171-180),
 * filter length = tail_ms*rate/1000 (:194), SPEEX_ECHO_SET_SAMPLING_RATE, preprocessor bound to the echo state
 * (:203; denoise on, residual-echo suppression on, AGC/VAD/dereverb off).
 * speexdsp is NOT under /root/reference; the MDF two-path canceller and the preprocessor are restated in
 * oracle/oracle_aec.c (parity unpinned by the reference's own tests: SURVEY §8c).
 * One call processes `nframes` consecutive frames of every stream. Layouts: mic, ref, out [stream][nframes*framesize]. */
typedef struct msb200_aec msb200_aec;
typedef struct msb200_aec_info {
	int32_t frame_size, window_size, M, sample_rate, filter_length;
	size_t state_bytes_per_stream;
} msb200_aec_info;
MSB200_API int msb200_aec_frame_size_for_rate(int sample_rate, int framesize_at_8000); /* adjust_framesize :171-180 */
MSB200_API int msb200_aec_create(msb200_ctx *ctx, int n_streams, int sample_rate, int tail_length_ms,
                                 int framesize_at_8000, msb200_aec **out);
MSB200_API void msb200_aec_destroy(msb200_aec *a);
MSB200_API int msb200_aec_get_info(msb200_aec *a, msb200_aec_info *info);
MSB200_API int msb200_aec_reset(msb200_aec *a, int stream); /* stream < 0: all */
MSB200_API int msb200_aec_set_live(msb200_aec *a, int n_live); /* see msb200_volume_set_live */
/* Cross-check switch for the 48 kHz kernel builds (results are bit-identical, the tests assert it): 0 = default, 1 = the
 * 256-thread build (one of the per-bin threads runs the frame's sequential IIR filters, the others wait), 3 / 4 = builds
 * with a ninth "serial" warp that runs those filters beside the per-bin threads (3: at 3 CTAs per SM; 4: the serial warp
 * also feeds the block pass with cp.async.bulk row copies instead of per-thread cp.async); 5 = the default build with the
 * generic loop of the block pass for every frame (the default runs the common frame through a tighter form of that loop;
 * any frame size). */
MSB200_API int msb200_aec_set_path(msb200_aec *a, int path);
MSB200_API int msb200_aec_process(msb200_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out, int nframes);
MSB200_API int msb200_aec_process_dev(msb200_aec *a, const void *d_mic, const void *d_ref, void *d_out, int nframes,
                                      int stride_samples);
/* as msb200_aec_process, with host buffers laid out [stream][stride_samples] (a fixed staging arena whose frame count
 * varies from tick to tick: the plugin's lockstep batch mode) */
MSB200_API int msb200_aec_process_strided(msb200_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out,
                                          int nframes, int stride_samples);
/* ragged batches: counts[stream] (NULL = nframes for all) is the number of frames stream `stream` really staged in this
 * call; a stream with fewer frames than `nframes` runs only its own — its far-end history and adaptive filter never see
 * padding — and its output rows are left untouched beyond its count. Every stream keeps its own position in the far-end
 * ring, so streams whose 10 ms ticks fall differently against the frame grid can share a bank. */
MSB200_API int msb200_aec_process_counts(msb200_aec *a, const int16_t *mic, const int16_t *ref, int16_t *out, int nframes,
                                         int stride_samples, const int32_t *counts);
MSB200_API int msb200_aec_process_counts_dev(msb200_aec *a, const void *d_mic, const void *d_ref, void *d_out, int nframes,
                                             int stride_samples, const void *d_counts_i32);
/* MS_ECHO_CANCELLER_GET/SET_STATE_STRING (speexec.c:119-167, 361-374): the adaptive-filter weights of one stream as
 * an opaque blob (our format: header + W[M][N] float). */
MSB200_API size_t msb200_aec_state_blob_size(msb200_aec *a);
MSB200_API int msb200_aec_get_state_blob(msb200_aec *a, int stream, void *blob, size_t size);
MSB200_API int msb200_aec_set_state_blob(msb200_aec *a, int stream, const void *blob, size_t size);
/* Debug/parity probes: copies a named internal per-stream vector to host ("W","foreground","X","power","power_1",
 * "prop","noise","echo_noise","gain2", scalar pack "scalars"). Returns number of floats written or <0. */
MSB200_API int msb200_aec_probe(msb200_aec *a, int stream, const char *what, float *out, int max_floats);

/* ---------------------------------------------------------------------------------------------------- MSAudioFlowControl
 * ms_audio_flow_controller_process() /root/reference/src/audiofilters/flowcontrol.c:110-150 (+ :58-108): while a drop
 * target is armed (ms_audio_flow_controller_set_target :51-56, from MS_AUDIO_FLOW_CONTROL_DROP :196-207) a block passes,
 * is dropped whole, or loses a few samples at the flattest three-sample runs. One block of `nsamples` per stream per
 * call, in place; out_nsamples[stream] = samples that remain (0: block dropped). Bit-exact. */
#define MSB200_FLOWCONTROL_BASIC 0 /* MSAudioFlowControlBasic */
#define MSB200_FLOWCONTROL_SOFT 1  /* MSAudioFlowControlSoft */
typedef struct msb200_flowcontrol msb200_flowcontrol;
typedef struct msb200_flowcontrol_state { /* MSAudioFlowController, include/mediastreamer2/flowcontrol.h:37-43 */
	int32_t strategy;
	float silent_threshold;
	uint32_t target_samples, total_samples, current_pos, current_dropped;
} msb200_flowcontrol_state;
MSB200_API int msb200_flowcontrol_create(msb200_ctx *ctx, int n_streams, int max_block, msb200_flowcontrol **out);
MSB200_API void msb200_flowcontrol_destroy(msb200_flowcontrol *f);
MSB200_API int msb200_flowcontrol_set_config(msb200_flowcontrol *f, int stream, int strategy, float silent_threshold);
MSB200_API int msb200_flowcontrol_set_target(msb200_flowcontrol *f, int stream, uint32_t samples_to_drop, uint32_t total_samples);
MSB200_API int msb200_flowcontrol_reset(msb200_flowcontrol *f, int stream);
MSB200_API int msb200_flowcontrol_get_state(msb200_flowcontrol *f, int stream, msb200_flowcontrol_state *st);
MSB200_API int msb200_flowcontrol_process(msb200_flowcontrol *f, int16_t *io, int nsamples, int32_t *out_nsamples);
MSB200_API int msb200_flowcontrol_process_dev(msb200_flowcontrol *f, void *d_io, int nsamples, int stride_samples,
                                              void *d_out_nsamples);

/* ---------------------------------------------------------------------------------------------------- Generic PLC
 * MSGenericPLC (/root/reference/src/audiofilters/msgenericplc.c:61-157 over src/audiofilters/genericplc.c:74-241):
 * packet-loss concealment by spectral stretching of the last 50 ms (windowed N-point real FFT, packed bins moved to
 * twice their index x 0.85, 2N-point inverse; ms_fft / ms_ifft = float kiss_fft, src/utils/dsptools.c:362-376), a 5 ms
 * continuity delay with cross-fades in and out of the concealed stretch, fade to silence between 100 and 150 ms.
 * One bank = n_streams mono streams at one rate (8 / 16 / 32 / 48 kHz: transform sizes must factor into 2, 3, 4, 5).
 * The concealer clock (MSConcealerContext, src/base/mscommon.c:315-362) is the caller's: each tick it passes one mode
 * byte per stream. Concealed samples are bit-identical to the reference's. */
#define MSB200_PLC_IDLE 0     /* nothing for this stream in this call */
#define MSB200_PLC_PACKET 1   /* io[stream] holds a received block: delayed / cross-faded in place (:64-117) */
#define MSB200_PLC_CONCEAL 2  /* write nsamples concealed samples to io[stream] (:147-152) */
#define MSB200_PLC_AFTER_CNG 4 /* OR-ed with PACKET: the filter was emitting comfort noise before this block (:77-88) */
typedef struct msb200_plc msb200_plc;
MSB200_API int msb200_plc_create(msb200_ctx *ctx, int n_streams, int sample_rate, int max_block, msb200_plc **out);
MSB200_API void msb200_plc_destroy(msb200_plc *p);
MSB200_API int msb200_plc_history_samples(const msb200_plc *p);
MSB200_API int msb200_plc_reset_stream(msb200_plc *p, int stream);
/* io: [n_streams][nsamples] s16 (host), mode: [n_streams] */
MSB200_API int msb200_plc_process(msb200_plc *p, int16_t *io, int nsamples, const uint8_t *mode);
/* rows `stride_samples` apart (an arena with several units per stream); only streams [0, n_live) are copied and run */
MSB200_API int msb200_plc_process_strided(msb200_plc *p, int16_t *io, int nsamples, int stride_samples, const uint8_t *mode);
MSB200_API int msb200_plc_set_live(msb200_plc *p, int n_live);
MSB200_API int msb200_plc_process_dev(msb200_plc *p, void *d_io, int nsamples, int stride_samples, const void *d_mode);

/* ---------------------------------------------------------------------------------------------------- G.711
 * MSAlawDec / MSUlawDec (/root/reference/src/audiofilters/alaw.c:199-211, ulaw.c) and the arithmetic of MSAlawEnc /
 * MSUlawEnc (alaw.c:84-87): Snack_Alaw2Lin / Snack_Mulaw2Lin / Snack_Lin2Alaw / Snack_Lin2Mulaw
 * (src/audiofilters/g711.c:119-262), over a flat batch of n samples (any concatenation of RTP payloads: the codec is
 * stateless). Bit-exact. The encoders' re-framing to ptime stays on the host (plugin/msb200_filters.c). */
#define MSB200_G711_ALAW 0 /* PCMA */
#define MSB200_G711_ULAW 1 /* PCMU */
MSB200_API int msb200_g711_decode(msb200_ctx *ctx, int law, const uint8_t *code, int16_t *pcm, size_t n);
MSB200_API int msb200_g711_encode(msb200_ctx *ctx, int law, const int16_t *pcm, uint8_t *code, size_t n);
MSB200_API int msb200_g711_decode_dev(msb200_ctx *ctx, int law, const void *d_code, void *d_pcm, size_t n);
MSB200_API int msb200_g711_encode_dev(msb200_ctx *ctx, int law, const void *d_pcm, void *d_code, size_t n);

/* ---------------------------------------------------------------------------------------------------- RTP payload hand-off
 * The caller side of the audio path in a media server: what MSRtpRecv does to every received packet before
 * the decoder sees it — header fields into the block's meta data, b_rptr advanced to the payload
 * (/root/reference/src/otherfilters/msrtp.c:1050-1092) — and what MSRtpSend does to every block the encoder emits — a
 * header in front (:617-705) — for ALL streams of a ticker at once, fused with the G.711 banks: packets in, PCM out (or
 * left in HBM for the next bank), and PCM in, packets out. One payload copy per packet (into / out of a pinned arena),
 * one launch per direction per tick. Sockets, jitter buffer, RTCP, SRTP stay the host's (oRTP's). RFC 3550 §5.1 / §5.3.1. */
typedef struct msb200_rtp_meta { /* what receiver_process copies from the RTP header into the mblk_t (:1078-1080) */
	uint32_t timestamp;  /* mblk_set_timestamp_info(m, rtp_get_timestamp(m)) */
	uint32_t ssrc;
	int32_t payload_len; /* bytes; 0 = nothing for this stream in this tick */
	uint16_t seq;        /* mblk_set_cseq(m, rtp_get_seqnumber(m)) */
	uint8_t marker;      /* mblk_set_marker_info(m, rtp_get_markbit(m)) */
	uint8_t payload_type;
} msb200_rtp_meta;
/* header of one packet: fills *meta, *payload_offset = first payload byte (rtp_get_payload); CSRC list, header extension
 * and padding are skipped; MSB200_EINVAL for anything that is not a well-formed version 2 packet. Host only. */
MSB200_API int msb200_rtp_parse(const uint8_t *packet, size_t len, msb200_rtp_meta *meta, size_t *payload_offset);
typedef struct msb200_rtp_rx msb200_rtp_rx;
MSB200_API int msb200_rtp_rx_create(msb200_ctx *ctx, int n_streams, int law, int max_payload, msb200_rtp_rx **out);
MSB200_API void msb200_rtp_rx_destroy(msb200_rtp_rx *r);
MSB200_API int msb200_rtp_rx_row_samples(const msb200_rtp_rx *r); /* samples per stream row of the PCM output (>= max_payload) */
MSB200_API int msb200_rtp_rx_begin_tick(msb200_rtp_rx *r);
/* returns the payload length taken (0: empty payload or another payload type than expected_pt >= 0), < 0 on error */
MSB200_API int msb200_rtp_rx_push(msb200_rtp_rx *r, int stream, const uint8_t *packet, size_t len, int expected_pt);
MSB200_API int msb200_rtp_rx_push_payload(msb200_rtp_rx *r, int stream, const uint8_t *payload, int len, const msb200_rtp_meta *meta);
/* pcm: host [n_streams][row_samples] s16, the first meta[s].payload_len samples of a row are stream s's decoded packet */
MSB200_API int msb200_rtp_rx_decode(msb200_rtp_rx *r, int16_t *pcm, msb200_rtp_meta *meta);
MSB200_API int msb200_rtp_rx_decode_dev(msb200_rtp_rx *r, void **d_pcm, const msb200_rtp_meta **meta);
typedef struct msb200_rtp_tx msb200_rtp_tx;
MSB200_API int msb200_rtp_tx_create(msb200_ctx *ctx, int n_streams, int law, int samples_per_packet, msb200_rtp_tx **out);
MSB200_API void msb200_rtp_tx_destroy(msb200_rtp_tx *t);
MSB200_API int msb200_rtp_tx_set_stream(msb200_rtp_tx *t, int stream, uint32_t ssrc, int payload_type, uint16_t next_seq,
                                        uint32_t next_ts);
MSB200_API size_t msb200_rtp_tx_packet_bytes(const msb200_rtp_tx *t); /* 12 + samples_per_packet */
/* pcm [n_streams][samples_per_packet] -> *packets = pinned arena [n_streams][packet_bytes], valid until the next call.
 * marker / send: NULL or [n_streams] flags; a stream with send[s] == 0 keeps its sequence number and timestamp. */
MSB200_API int msb200_rtp_tx_encode(msb200_rtp_tx *t, const int16_t *pcm, const uint8_t *marker, const uint8_t *send,
                                    const uint8_t **packets);
MSB200_API int msb200_rtp_tx_encode_dev(msb200_rtp_tx *t, const void *d_pcm, const uint8_t *marker, const uint8_t *send,
                                        const uint8_t **packets);

/* ---------------------------------------------------------------------------------------------------- audio chain
 * The BASELINE cfg2 pipeline as one resident device-side graph, one call per 10 ms tick for `n_streams` streams:
 *   ref  [in_rate] -> MSResample -> \
 *                                    MSSpeexEC(rate, tail) -> MSVolume(gain) -> [optional MSAudioMixer rooms of P]
 *   mic  [in_rate] -> MSResample -> /
 * (graph shape: /root/reference/src/voip/audiostream.c:1798-1832; conference: src/voip/audioconference.c:209-257).
 * The EC re-frames 10 ms ticks into `framesize` blocks exactly as its MSBufferizer does (speexec.c:252-259): the
 * number of output samples per tick varies (0, 1 or 2 frames); the volume stage runs per EC output block, as the
 * reference's per-mblk loop does (msvolume.c:505-512); the mixer stage re-frames to 10 ms (audiomixer.c:78-90). */
typedef struct msb200_chain msb200_chain;
typedef struct msb200_chain_params {
	int32_t n_streams;
	int32_t in_rate;         /* e.g. 16000 */
	int32_t rate;            /* e.g. 48000 */
	int32_t tail_length_ms;  /* 250 */
	int32_t framesize_at_8000; /* 64 */
	float volume_gain;       /* 0.8 */
	int32_t mixer_pins;      /* 0 = no mixer stage; else streams are grouped in rooms of this many pins (conference mode) */
	int32_t use_cuda_graph;  /* reserved, ignored: a tick is 3 launches on a GPU-bound ~0.8 ms step, there is no launch gap to remove */
} msb200_chain_params;
MSB200_API int msb200_chain_create(msb200_ctx *ctx, const msb200_chain_params *p, msb200_chain **out);
MSB200_API void msb200_chain_destroy(msb200_chain *c);
/* samples per stream the next tick will produce on the (pre-mixer) EC/volume output, and after the mixer */
MSB200_API int msb200_chain_next_out_samples(msb200_chain *c);
MSB200_API int msb200_chain_max_out_samples(msb200_chain *c);
/* Host path: ref_in/mic_in [stream][in_rate/100] s16 -> out [stream][max_out_samples] s16 (first *out_samples valid).
 * H2D, all kernels, D2H inside. */
MSB200_API int msb200_chain_tick(msb200_chain *c, const int16_t *ref_in, const int16_t *mic_in, int16_t *out,
                                 int *out_samples);
/* Pipelined host path for free-running hosts: submit() enqueues tick T (H2D of its inputs, its kernels, D2H of its
 * output into `out`) on three streams and returns at once — *out_samples is known immediately, the samples are in `out`
 * after the matching wait(). At most two ticks in flight: tick T's input copy and tick T-1's output copy overlap the
 * kernels. ref_in / mic_in / out must be pinned (msb200_host_alloc_pinned) and stay untouched until that wait().
 * Results are identical to msb200_chain_tick(); do not mix the two styles while ticks are in flight. */
MSB200_API int msb200_chain_submit(msb200_chain *c, const int16_t *ref_in, const int16_t *mic_in, int16_t *out,
                                   int *out_samples);
MSB200_API int msb200_chain_wait(msb200_chain *c); /* waits for the OLDEST tick in flight (no-op when none) */
/* Device path: inputs/outputs already resident; asynchronous. */
MSB200_API int msb200_chain_tick_dev(msb200_chain *c, const void *d_ref_in, const void *d_mic_in, void *d_out,
                                     int *out_samples);
/* Overlap mode of msb200_chain_tick_dev (msb200_chain_submit: with MSB200_CHAIN_OVERLAP=1): the resamplers of tick T+1 and the
 * volume / hand-out of tick T-1 run on side streams beside the echo canceller of tick T, so the cancellers of consecutive
 * ticks run back to back. Same samples. What changes for the caller: d_ref_in / d_mic_in must be COMPLETE when the call is
 * made (not merely ordered on the context's stream), and d_out is complete in the context's stream order only after
 * msb200_chain_join(). Ignored for conference chains (mixer_pins > 0). */
MSB200_API int msb200_chain_set_overlap(msb200_chain *c, int enabled);
MSB200_API int msb200_chain_join(msb200_chain *c);
MSB200_API int msb200_chain_launches_per_tick(msb200_chain *c);
/* Per-kernel device timing of the dominant kernel (the echo canceller) with CUDA events recorded on the launching
 * stream around every AEC launch while enabled; get_ returns the accumulated milliseconds, launches and frames since
 * the last call and resets the counters (synchronises the stream). Used by bench.py for the roofline figure. */
MSB200_API int msb200_chain_enable_kernel_timing(msb200_chain *c, int enabled);
MSB200_API int msb200_chain_get_kernel_timing(msb200_chain *c, float *aec_ms, int *aec_launches, int *aec_frames);
MSB200_API msb200_aec *msb200_chain_aec(msb200_chain *c);

/* ---------------------------------------------------------------------------------------------------- video
 * Pixel formats: the reference's MSPixFmt values (include/mediastreamer2/msvideo.h:267-280) plus NV12/NV21 appended
 * (the reference has no such member; SURVEY §8a note). */
#define MSB200_PIX_YUV420P 0
#define MSB200_PIX_YUYV 1
#define MSB200_PIX_RGB24 2
#define MSB200_PIX_RGB24_REV 3 /* BGR24 bottom-up in the reference's world; here: BGR byte order */
#define MSB200_PIX_UYVY 5
#define MSB200_PIX_YUY2 6
#define MSB200_PIX_RGBA32 7      /* R G B A bytes */
#define MSB200_PIX_RGB565 8      /* MS_RGB565 -> AV_PIX_FMT_RGB565 little endian (msvideo.c:610-611): 16 bits per pixel, r5 g6 b5 */
#define MSB200_PIX_RGBA32_REV 11 /* B G R A bytes (MS_RGBA32_REV -> AV_PIX_FMT_BGRA, msvideo.c:600-601) */
#define MSB200_PIX_NV12 100
#define MSB200_PIX_NV21 101

/* NV12/NV21 -> I420 with rotation {0,90,180,270} and optional nearest 1/2 decimation: bit-exact replacement of
 * copy_ycbcrbiplanar_to_true_yuv_with_rotation_and_down_scale_by_2() /root/reference/src/voip/msvideo.c:787-919.
 * (w,h) is the DESTINATION size before rotation, as in the reference. Batched over n_frames frames laid out
 * back-to-back: src frame = y plane (y_stride*src_h) then cbcr plane (cbcr_stride*src_h/2); dst = tight I420. */
MSB200_API int msb200_nv12_to_i420(msb200_ctx *ctx, int n_frames, const uint8_t *src, size_t src_frame_bytes,
                                   size_t cbcr_offset, int rotation, int w, int h, int y_stride, int cbcr_stride,
                                   int u_first, int down_scale, uint8_t *dst);
MSB200_API int msb200_nv12_to_i420_dev(msb200_ctx *ctx, int n_frames, const void *d_src, size_t src_frame_bytes,
                                       size_t cbcr_offset, int rotation, int w, int h, int y_stride, int cbcr_stride,
                                       int u_first, int down_scale, void *d_dst);

/* ms_yuv_buf_copy_with_pix_strides() /root/reference/src/voip/msvideo.c:245-270 (plane_copy :204-227, row_copy :188-202),
 * batched over n_frames frames laid out frame_bytes apart: copies the region src_roi of a three-plane YUV picture to
 * dst_roi of another, every plane with its own row stride and PIXEL stride — the reference's way of converting between
 * planar I420 and semi-planar NV12 / NV21 (pixel stride 2 on the chroma planes, plane 2 one byte after plane 1) and of
 * placing a picture inside a larger one (a compositor's tile). For planes 1 and 2 every field of both rectangles is
 * halved. Bytes outside the destination region are left as they are. Bit-exact, including plane_copy's one-memcpy
 * shortcut for equal strides and rectangles. Pinned by the reference's own eight test patterns
 * (tester/mediastreamer2_framework_tester.c:393-500). */
typedef struct msb200_rect { /* MSRect, include/mediastreamer2/msvideo.h */
	int32_t x, y, w, h;
} msb200_rect;
typedef struct msb200_yuv_layout {
	size_t plane_offset[3]; /* first byte of each plane inside a frame */
	int32_t row_stride[3];
	int32_t pix_stride[3];
	size_t frame_bytes;     /* distance between consecutive frames of the batch */
} msb200_yuv_layout;
MSB200_API int msb200_yuv_copy_strided(msb200_ctx *ctx, int n_frames, const uint8_t *src, const msb200_yuv_layout *src_layout,
                                       msb200_rect src_roi, uint8_t *dst, const msb200_yuv_layout *dst_layout,
                                       msb200_rect dst_roi);
MSB200_API int msb200_yuv_copy_strided_dev(msb200_ctx *ctx, int n_frames, const void *d_src,
                                           const msb200_yuv_layout *src_layout, msb200_rect src_roi, void *d_dst,
                                           const msb200_yuv_layout *dst_layout, msb200_rect dst_roi);

/* MSScaler replacement: create_context / context_process / context_free of MSScalerDesc
 * (include/mediastreamer2/msvideo.h:473-479; backends src/voip/msvideo.c:517-691), used by MSPixConv
 * (src/videofilters/pixconv.c:62-94) and MSSizeConv (src/videofilters/sizeconv.c:97-184), batched over n_frames.
 * Arithmetic: the swscale SWS_BILINEAR pipeline the reference's ffmpeg back-end runs (msvideo.c:651-681) restated in
 * oracle/oracle_video.c and pinned against libswscale 9.1.100 golden frames (tests/golden/).
 * Format pairs: YUV420P / NV12 / NV21 -> YUV420P / RGB24 / RGB24_REV(BGR byte order) with bilinear scaling, any sizes
 * >= 8 (the TMA-tiled kernels take source widths % 16 == 0 and down-scale factors < 2; everything else runs the
 * tile-free direct kernel, same arithmetic); MSPixConv's same-size conversions YUYV / YUY2 / UYVY / RGB24 / RGB24_REV / RGBA32 /
 * RGBA32_REV / RGB565 -> YUV420P (w % 8 == 0 for 4:2:2, w % 4 == 0 for RGB; h even). Frames are tight (no row padding), back to back. */
typedef struct msb200_scaler msb200_scaler;
MSB200_API int msb200_scaler_create(msb200_ctx *ctx, int src_w, int src_h, int src_fmt, int dst_w, int dst_h,
                                    int dst_fmt, msb200_scaler **out);
MSB200_API void msb200_scaler_destroy(msb200_scaler *s);
MSB200_API size_t msb200_scaler_src_frame_bytes(msb200_scaler *s);
MSB200_API size_t msb200_scaler_dst_frame_bytes(msb200_scaler *s);
MSB200_API int msb200_scaler_process(msb200_scaler *s, int n_frames, const uint8_t *src, uint8_t *dst);
MSB200_API int msb200_scaler_process_dev(msb200_scaler *s, int n_frames, const void *d_src, void *d_dst);
/* as msb200_scaler_process for frames that do NOT lie back to back on the host: src_frames / dst_frames are arrays of
 * n_frames host pointers (tight frames; pinned memory makes the copies asynchronous). Runs of adjacent frames are copied
 * in one piece; the kernels run once over the whole batch. What the plugin's MSPixConv / MSSizeConv use: a ticker's worth
 * of frames staged in a pinned arena in, pinned arena slots lent to the output mblks out. */
MSB200_API int msb200_scaler_process_frames(msb200_scaler *s, int n_frames, const uint8_t *const *src_frames,
                                            uint8_t *const *dst_frames);
/* Vertical rounding of the scaled PLANAR output (MSSizeConv's I420 -> I420, row a8). libswscale has two: its C functions
 * (what SWS_BITEXACT selects; the default here) and, on x86, the SIMD vertical scaler a plain SWS_BILINEAR call — the call
 * the reference makes, src/voip/msvideo.c:660 — really runs, which differs by at most 1 (per-tap truncation in 16-bit
 * lanes, a compensating rounder, the last two luma rows and the last chroma row by the C functions). on = 1 reproduces the
 * x86 result bit for bit (pinned: the oracle's same mode equals the live library on random geometries). Also accepted by
 * MSPixConv's RGB inputs (RGB24 / RGBA / BGRA / RGB565 -> YUV420P), whose chroma rows the library filters 2:1 with the same
 * vertical scaler (folded taps on the top row); BGR24's special converter has no vertical filter and is unaffected. */
MSB200_API int msb200_scaler_set_x86_vertical(msb200_scaler *s, int on);

/* Mosaic / compositor (the reference's building blocks: ms_yuv_buf_copy_with_pix_strides src/voip/msvideo.c:
 * 245-270 places a picture in a region of a larger one, src/voip/layouts.c computes the rectangles): after set_canvas the
 * scaler writes frame k of a batch straight into rectangle tiles[k % n_tiles] of canvas k / n_tiles (I420, canvas_w x
 * canvas_h, tight planes) — N participants scaled AND composed in the same launches, with the scaler's bit-exact
 * arithmetic; bytes outside the tiles are not touched. Planar-output scalers only (YUV420P / NV12 / NV21 -> YUV420P);
 * every tile is dst_w x dst_h, tile x % 8 == 0, tile y % 2 == 0, canvas_w % 8 == 0. n_tiles == 0 restores tight frames.
 * process_dev: n_frames must be a multiple of n_tiles; d_dst holds n_frames / n_tiles canvases. */
MSB200_API int msb200_scaler_set_canvas(msb200_scaler *s, int canvas_w, int canvas_h, int n_tiles, const msb200_rect *tiles);
MSB200_API size_t msb200_scaler_canvas_bytes(msb200_scaler *s);

/* Kernel selection, for tests and profiling only (every path is bit-exact with the others): path 0 = best available,
 * 1 = persistent tile kernel, 2 = generic tile kernel, 3 = register-window strip kernel (the default where it applies),
 * 4 = per-warp streaming variant of the strip kernel (experimental: no vertical halo, but slower on B200 today).
 * get_path reports what process() will launch: 4 = streaming kernel, 3 = strip kernel, 2 = persistent tile kernel,
 * 1 = generic tile kernel, 0 = plane / packed-4:2:2 kernels, 5 = tile-free direct kernel (geometry outside the TMA limits). */
MSB200_API int msb200_scaler_set_path(msb200_scaler *s, int path);
MSB200_API int msb200_scaler_get_path(msb200_scaler *s);
/* Strip kernel only: index of the static row schedule its straight-line instantiation runs (0 = 3:2 down-scale,
 * 1 = 1:1), or -1 when the general loop handles every strip. *regular_strips / *strips (may be
 * NULL) receive how many strips of a frame column follow the schedule. Environment MSB200_SCALER_NO_SCHED=1 at create
 * time forces the general loop (A/B runs). */
MSB200_API int msb200_scaler_get_schedule(msb200_scaler *s, int *regular_strips, int *strips);

#ifdef __cplusplus
}
#endif
#endif /* MSB200DSP_H */
