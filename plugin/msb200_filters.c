/*
 * plugin/msb200_filters.c — libmsb200filters.so: the B200 DSP hot path packaged as a mediastreamer2 plugin.
 *
 * Loaded by an UNMODIFIED mediastreamer2 through its own plugin loader (src/base/msfactory.c:531-586: dlopen of
 * libms*.so, then `void <file-without-.so>_init(MSFactory*)`). libmsb200filters_init() registers MSFilterDesc objects
 * that reuse the built-in filters' ids, names, pin counts, flags and method tables, so that
 * ms_factory_create_filter(id) / _from_name(name) return these instead (registration prepends, :259-282 / :429-450).
 *
 * Every filter here is host-side control logic only (queues, bufferizers, flow control, method calls — the parts of
 * the reference filters that depend on ticker->time and on the mblk_t contract); all sample arithmetic happens in
 * libmsb200dsp.so's CUDA kernels through the C ABI of include/msb200dsp.h. There is no CPU fallback: if no GPU context
 * can be created the filters log an error and drop their input.
 *
 * Execution mode: synchronous — each process() call makes one bank call for its own stream (exact reference
 * semantics, one tick latency-free). The batched path (thousands of streams per launch) is the msb200_chain / bank API
 * itself; DESIGN.md §7 describes the deferred-batch mode that connects the two.
 *
 * Compiled against the host's mediastreamer2 / oRTP / bctoolbox headers (here: /root/reference/include + compat/).
 */
#include "mediastreamer2/msaudiomixer.h"
#include "mediastreamer2/mschanadapter.h"
#include "mediastreamer2/msequalizer.h"
#include "mediastreamer2/msfactory.h"
#include "mediastreamer2/msfilter.h"
#include "mediastreamer2/msinterfaces.h"
#include "mediastreamer2/msticker.h"
#include "mediastreamer2/msvideo.h"
#include "mediastreamer2/msvolume.h"

#include "msb200dsp.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ device context */
static msb200_ctx *g_ctx = NULL;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER; /* the context's stream is shared by all filter instances */

static msb200_ctx *dsp_ctx(void) {
	if (!g_ctx) {
		const char *dev = getenv("MSB200_DEVICE");
		int rc = msb200_ctx_create(dev ? atoi(dev) : 0, &g_ctx);
		if (rc != MSB200_OK) {
			ms_error("msb200: cannot create the GPU context (%s); filters will drop audio/video", msb200_last_error());
			g_ctx = NULL;
		}
	}
	return g_ctx;
}
#define DSP_LOCK() pthread_mutex_lock(&g_mu)
#define DSP_UNLOCK() pthread_mutex_unlock(&g_mu)
#define DSP_CHECK(expr, what)                                                                                          \
	do {                                                                                                               \
		if ((expr) != MSB200_OK) ms_error("msb200: %s failed: %s", what, msb200_last_error());                         \
	} while (0)

/* ================================================================================================ MSAudioMixer
 * host logic restated from /root/reference/src/audiofilters/audiomixer.c: channel bufferizers :78-90, flow control
 * :92-111, bypass mode :219-286, output dispatch :288-346, methods :348-431 */
#define MIXER_MAX_CHANNELS 50
#define BYPASS_MODE_TIMEOUT 1000

typedef struct MixChannel {
	MSBufferizer bufferizer;
	float gain;
	int min_fullness;
	uint64_t last_flow_control;
	uint64_t last_activity;
	bool_t active;
	bool_t output_enabled;
} MixChannel;

typedef struct MixerState {
	int nchannels, rate, bytespertick, conf_mode, skip_threshold, master_channel;
	MixChannel channels[MIXER_MAX_CHANNELS];
	bool_t bypass_mode, single_output;
	msb200_mixer *bank; /* 1 room x 50 pins x nwords */
	int16_t *in;        /* [50][nwords] */
	uint8_t *present;   /* [50] */
	int16_t *out;       /* [50][nwords] (conference) or [nwords] */
} MixerState;

static void mixer_init(MSFilter *f) {
	MixerState *s = ms_new0(MixerState, 1);
	int i;
	s->conf_mode = FALSE;
	s->nchannels = 1;
	s->rate = 44100;
	s->master_channel = -1;
	for (i = 0; i < MIXER_MAX_CHANNELS; ++i) {
		ms_bufferizer_init(&s->channels[i].bufferizer);
		s->channels[i].gain = 1.0f;
		s->channels[i].active = TRUE;
		s->channels[i].output_enabled = TRUE;
	}
	f->data = s;
}
static void mixer_uninit(MSFilter *f) {
	MixerState *s = (MixerState *)f->data;
	int i;
	for (i = 0; i < MIXER_MAX_CHANNELS; ++i)
		ms_bufferizer_uninit(&s->channels[i].bufferizer);
	ms_free(s);
}
static bool_t mixer_has_single_output(MSFilter *f, MixerState *s) {
	int i, count = 0;
	for (i = 0; i < f->desc->noutputs; ++i)
		if (f->outputs[i] && s->channels[i].output_enabled) count++;
	return count == 1;
}
static void mixer_preprocess(MSFilter *f) {
	MixerState *s = (MixerState *)f->data;
	int i, nwords;
	s->bytespertick = (2 * s->nchannels * s->rate * f->ticker->interval) / 1000;
	nwords = s->bytespertick / 2;
	for (i = 0; i < MIXER_MAX_CHANNELS; ++i) {
		s->channels[i].last_flow_control = (uint64_t)-1;
		s->channels[i].last_activity = (uint64_t)-1;
	}
	s->skip_threshold = s->bytespertick * 2;
	s->bypass_mode = FALSE;
	s->single_output = mixer_has_single_output(f, s);
	s->in = (int16_t *)ms_malloc0(sizeof(int16_t) * MIXER_MAX_CHANNELS * (size_t)nwords);
	s->out = (int16_t *)ms_malloc0(sizeof(int16_t) * MIXER_MAX_CHANNELS * (size_t)nwords);
	s->present = (uint8_t *)ms_malloc0(MIXER_MAX_CHANNELS);
	DSP_LOCK();
	if (dsp_ctx()) {
		DSP_CHECK(msb200_mixer_create(g_ctx, 1, MIXER_MAX_CHANNELS, nwords, s->conf_mode, &s->bank), "mixer_create");
		for (i = 0; s->bank && i < MIXER_MAX_CHANNELS; ++i) {
			msb200_mixer_set_input_gain(s->bank, 0, i, s->channels[i].gain);
			msb200_mixer_set_active(s->bank, 0, i, s->channels[i].active);
		}
	}
	DSP_UNLOCK();
}
static void mixer_postprocess(MSFilter *f) {
	MixerState *s = (MixerState *)f->data;
	DSP_LOCK();
	msb200_mixer_destroy(s->bank);
	DSP_UNLOCK();
	s->bank = NULL;
	ms_free(s->in);
	ms_free(s->out);
	ms_free(s->present);
	s->in = s->out = NULL;
	s->present = NULL;
}
static void mixer_dispatch_output(MSFilter *f, MixerState *s, MSQueue *inq, int active_input) {
	int i;
	for (i = 0; i < f->desc->noutputs; i++) {
		MSQueue *outq = f->outputs[i];
		if (outq && s->channels[i].output_enabled && (active_input != i || s->conf_mode == 0)) {
			mblk_t *m;
			if (s->single_output) {
				while ((m = ms_queue_get(inq)) != NULL)
					ms_queue_put(outq, m);
				break;
			}
			for (m = ms_queue_peek_first(inq); !ms_queue_end(inq, m); m = ms_queue_next(inq, m))
				ms_queue_put(outq, dupmsg(m));
		}
	}
	ms_queue_flush(inq);
}
static bool_t mixer_check_bypass(MSFilter *f, MixerState *s) {
	int i, active_cnt = 0, active_input = -1;
	MSQueue *activeq = NULL;
	uint64_t curtime = f->ticker->time;
	for (i = 0; i < f->desc->ninputs; i++) {
		MSQueue *q = f->inputs[i];
		MixChannel *chan = &s->channels[i];
		if (!q) continue;
		if (!ms_queue_empty(q)) {
			chan->last_activity = curtime;
			activeq = q;
			active_cnt++;
			active_input = i;
		} else if (chan->last_activity == (uint64_t)-1) {
			chan->last_activity = curtime;
		} else if (curtime - chan->last_activity < BYPASS_MODE_TIMEOUT) {
			activeq = q;
			active_cnt++;
			active_input = i;
		}
	}
	if (active_cnt == 1) {
		if (!s->bypass_mode) {
			s->bypass_mode = TRUE;
			ms_message("MSAudioMixer(b200) [%p] is entering bypass mode.", f);
		}
		mixer_dispatch_output(f, s, activeq, active_input);
		return TRUE;
	} else if (active_cnt > 1) {
		if (s->bypass_mode) {
			s->bypass_mode = FALSE;
			ms_message("MSAudioMixer(b200) [%p] is leaving bypass mode.", f);
		}
		return FALSE;
	}
	return TRUE;
}
static void mixer_process(MSFilter *f) {
	MixerState *s = (MixerState *)f->data;
	int i, nwords = s->bytespertick / 2;
	ms_filter_lock(f);
	if (mixer_check_bypass(f, s)) {
		ms_filter_unlock(f);
		return;
	}
	memset(s->present, 0, MIXER_MAX_CHANNELS);
	for (i = 0; i < f->desc->ninputs; ++i) {
		MSQueue *q = f->inputs[i];
		MixChannel *chan = &s->channels[i];
		int size, skip = 0;
		if (!q) continue;
		ms_bufferizer_put_from_queue(&chan->bufferizer, q);
		if (ms_bufferizer_read(&chan->bufferizer, (uint8_t *)(s->in + (size_t)i * nwords), (size_t)nwords * 2) != 0)
			s->present[i] = 1;
		/* channel_flow_control */
		if (chan->last_flow_control == (uint64_t)-1) {
			chan->last_flow_control = f->ticker->time;
			chan->min_fullness = -1;
			continue;
		}
		size = (int)ms_bufferizer_get_avail(&chan->bufferizer);
		if (chan->min_fullness == -1 || size < chan->min_fullness) chan->min_fullness = size;
		if (f->ticker->time - chan->last_flow_control >= 5000) {
			if (chan->min_fullness >= s->skip_threshold) {
				skip = chan->min_fullness - (s->skip_threshold / 2);
				ms_bufferizer_skip_bytes(&chan->bufferizer, skip);
			}
			chan->last_flow_control = f->ticker->time;
			chan->min_fullness = -1;
		}
		if (skip > 0)
			ms_warning("Too much data in channel %i, %i ms in excess dropped", i, (skip * 1000) / (2 * s->nchannels * s->rate));
	}
	/* the arithmetic: one launch for the whole mixer (sum, gains, minus-own, saturation) */
	DSP_LOCK();
	if (s->bank) DSP_CHECK(msb200_mixer_process(s->bank, s->in, s->present, s->out), "mixer_process");
	DSP_UNLOCK();
	if (s->bank) {
		if (s->conf_mode == 0) {
			mblk_t *om = NULL;
			for (i = 0; i < MIXER_MAX_CHANNELS; ++i) {
				MSQueue *q = f->outputs[i];
				if (q && s->channels[i].output_enabled) {
					if (om == NULL) {
						om = allocb((size_t)nwords * 2, 0);
						memcpy(om->b_wptr, s->out, (size_t)nwords * 2);
						om->b_wptr += nwords * 2;
					} else {
						om = dupb(om);
					}
					ms_queue_put(q, om);
				}
			}
		} else {
			for (i = 0; i < MIXER_MAX_CHANNELS; ++i) {
				MSQueue *q = f->outputs[i];
				if (q && s->channels[i].output_enabled) {
					mblk_t *om = allocb((size_t)nwords * 2, 0);
					memcpy(om->b_wptr, s->out + (size_t)i * nwords, (size_t)nwords * 2);
					om->b_wptr += nwords * 2;
					ms_queue_put(q, om);
				}
			}
		}
	}
	ms_filter_unlock(f);
}
static int mixer_set_rate(MSFilter *f, void *data) {
	((MixerState *)f->data)->rate = *(int *)data;
	return 0;
}
static int mixer_get_rate(MSFilter *f, void *data) {
	*(int *)data = ((MixerState *)f->data)->rate;
	return 0;
}
static int mixer_set_nchannels(MSFilter *f, void *data) {
	((MixerState *)f->data)->nchannels = *(int *)data;
	return 0;
}
static int mixer_get_nchannels(MSFilter *f, void *data) {
	*(int *)data = ((MixerState *)f->data)->nchannels;
	return 0;
}
static int mixer_set_input_gain(MSFilter *f, void *data) {
	MixerState *s = (MixerState *)f->data;
	MSAudioMixerCtl *ctl = (MSAudioMixerCtl *)data;
	if (ctl->pin < 0 || ctl->pin >= MIXER_MAX_CHANNELS) {
		ms_warning("mixer_set_input_gain: invalid pin number %i", ctl->pin);
		return -1;
	}
	s->channels[ctl->pin].gain = ctl->param.gain;
	if (s->bank) {
		DSP_LOCK();
		msb200_mixer_set_input_gain(s->bank, 0, ctl->pin, ctl->param.gain);
		DSP_UNLOCK();
	}
	return 0;
}
static int mixer_set_active(MSFilter *f, void *data) {
	MixerState *s = (MixerState *)f->data;
	MSAudioMixerCtl *ctl = (MSAudioMixerCtl *)data;
	if (ctl->pin < 0 || ctl->pin >= MIXER_MAX_CHANNELS) {
		ms_warning("mixer_set_active_gain: invalid pin number %i", ctl->pin);
		return -1;
	}
	s->channels[ctl->pin].active = (bool_t)ctl->param.active;
	if (s->bank) {
		DSP_LOCK();
		msb200_mixer_set_active(s->bank, 0, ctl->pin, ctl->param.active);
		DSP_UNLOCK();
	}
	return 0;
}
static int mixer_enable_output(MSFilter *f, void *data) {
	MixerState *s = (MixerState *)f->data;
	MSAudioMixerCtl *ctl = (MSAudioMixerCtl *)data;
	if (ctl->pin < 0 || ctl->pin >= MIXER_MAX_CHANNELS) {
		ms_warning("mixer_enable_output: invalid pin number %i", ctl->pin);
		return -1;
	}
	ms_filter_lock(f);
	s->channels[ctl->pin].output_enabled = (bool_t)ctl->param.enabled;
	s->single_output = mixer_has_single_output(f, s);
	ms_filter_unlock(f);
	return 0;
}
static int mixer_set_conference_mode(MSFilter *f, void *data) {
	((MixerState *)f->data)->conf_mode = *(int *)data;
	return 0;
}
static int mixer_set_master_channel(MSFilter *f, void *data) {
	((MixerState *)f->data)->master_channel = *(int *)data;
	return 0;
}
static MSFilterMethod mixer_methods[] = {{MS_FILTER_SET_NCHANNELS, mixer_set_nchannels},
                                         {MS_FILTER_GET_NCHANNELS, mixer_get_nchannels},
                                         {MS_FILTER_SET_SAMPLE_RATE, mixer_set_rate},
                                         {MS_FILTER_GET_SAMPLE_RATE, mixer_get_rate},
                                         {MS_AUDIO_MIXER_SET_INPUT_GAIN, mixer_set_input_gain},
                                         {MS_AUDIO_MIXER_SET_ACTIVE, mixer_set_active},
                                         {MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE, mixer_set_conference_mode},
                                         {MS_AUDIO_MIXER_SET_MASTER_CHANNEL, mixer_set_master_channel},
                                         {MS_AUDIO_MIXER_ENABLE_OUTPUT, mixer_enable_output},
                                         {0, NULL}};
static MSFilterDesc b200_audio_mixer_desc = {.id = MS_AUDIO_MIXER_ID,
                                             .name = "MSAudioMixer",
                                             .text = "B200: mixes 16 bit sample audio streams (libmsb200dsp)",
                                             .category = MS_FILTER_OTHER,
                                             .ninputs = MIXER_MAX_CHANNELS,
                                             .noutputs = MIXER_MAX_CHANNELS,
                                             .init = mixer_init,
                                             .preprocess = mixer_preprocess,
                                             .process = mixer_process,
                                             .postprocess = mixer_postprocess,
                                             .uninit = mixer_uninit,
                                             .methods = mixer_methods,
                                             .flags = MS_FILTER_IS_PUMP};

/* ================================================================================================ MSVolume
 * /root/reference/src/audiofilters/msvolume.c:471-514: light path (:503-513) works in place per mblk; with AGC or an
 * echo-limiter peer (:480-502) the input is re-framed to 10 ms chunks (MSBufferizer) and the kernel additionally runs the
 * echo avoider against the peer filter's energy and the AGC gain reduction. */
typedef struct VolState {
	int rate;
	float static_gain;
	int noise_gate, remove_dc, agc;
	float ng_threshold, ng_floorgain;
	MSFilter *peer;
	msb200_volume *bank;
	int bank_rate;
	bool_t dirty, gain_dirty, peer_linked;
	MSBufferizer *buffer;
	float ea_thres, ea_speed, ea_force, ea_transmit;
	int ea_sustain;
} VolState;
static MSFilterDesc b200_volume_desc;
#define VOL_MAX_BLOCK 8192

static void vol_sync_config(VolState *v) { /* DSP lock held */
	if (!v->bank) return;
	if (v->peer && !v->peer_linked && v->peer->desc == &b200_volume_desc && ((VolState *)v->peer->data)->bank) {
		msb200_volume_set_peer(v->bank, 0, ((VolState *)v->peer->data)->bank, 0);
		v->peer_linked = TRUE;
	}
	if (v->gain_dirty) { /* MS_VOLUME_SET_GAIN resets the ramp (gain = target = static, msvolume.c:270-276): apply it once */
		msb200_volume_set_gain(v->bank, 0, v->static_gain);
		v->gain_dirty = FALSE;
	}
	if (!v->dirty) return;
	if (v->noise_gate) {
		msb200_volume_enable_noise_gate(v->bank, 0, 1);
		msb200_volume_set_noise_gate_threshold(v->bank, 0, v->ng_threshold);
		msb200_volume_set_noise_gate_floorgain(v->bank, 0, v->ng_floorgain);
	}
	msb200_volume_remove_dc(v->bank, 0, v->remove_dc);
	msb200_volume_enable_agc(v->bank, 0, v->agc);
	msb200_volume_set_ea_threshold(v->bank, 0, v->ea_thres);
	msb200_volume_set_ea_speed(v->bank, 0, v->ea_speed);
	msb200_volume_set_ea_force(v->bank, 0, v->ea_force);
	msb200_volume_set_ea_sustain(v->bank, 0, v->ea_sustain);
	msb200_volume_set_ea_transmit_threshold(v->bank, 0, v->ea_transmit);
	v->dirty = FALSE;
}
static void vol_init(MSFilter *f) {
	VolState *v = ms_new0(VolState, 1);
	v->rate = 8000;
	v->static_gain = 1.0f;
	v->ng_threshold = 0.1f;
	v->ng_floorgain = 0.005f;
	v->ea_thres = 0.1f;
	v->ea_speed = 0.4f;
	v->ea_force = 4.0f;
	v->ea_transmit = 4.0f;
	v->ea_sustain = 200;
	v->buffer = ms_bufferizer_new();
	v->dirty = TRUE;
	v->gain_dirty = FALSE; /* the bank starts at gain 1 like volume_init */
	f->data = v;
}
static void vol_uninit(MSFilter *f) {
	VolState *v = (VolState *)f->data;
	DSP_LOCK();
	msb200_volume_destroy(v->bank);
	DSP_UNLOCK();
	ms_bufferizer_destroy(v->buffer);
	ms_free(v);
}
static void vol_preprocess(MSFilter *f) {
	VolState *v = (VolState *)f->data;
	DSP_LOCK();
	if (dsp_ctx() && (!v->bank || v->bank_rate != v->rate)) {
		msb200_volume_destroy(v->bank);
		v->bank = NULL;
		DSP_CHECK(msb200_volume_create(g_ctx, 1, v->rate, VOL_MAX_BLOCK, &v->bank), "volume_create");
		v->bank_rate = v->rate;
		v->dirty = TRUE;
		v->gain_dirty = v->static_gain != 1.0f;
		v->peer_linked = FALSE;
	}
	vol_sync_config(v);
	DSP_UNLOCK();
	if (v->peer && v->peer->desc != &b200_volume_desc) ms_warning("MSVolume(b200): the echo-limiter peer is not a B200 MSVolume; ignored");
}
static void vol_process(MSFilter *f) {
	VolState *v = (VolState *)f->data;
	mblk_t *m;
	if (v->agc || v->peer != NULL) { /* chunked mode :480-502 */
		int nsamples = (int)(0.01 * (float)v->rate);
		size_t nbytes = (size_t)nsamples * 2;
		ms_bufferizer_put_from_queue(v->buffer, f->inputs[0]);
		while (ms_bufferizer_get_avail(v->buffer) >= nbytes) {
			m = allocb(nbytes, 0);
			ms_bufferizer_read(v->buffer, m->b_wptr, nbytes);
			m->b_wptr += nbytes;
			DSP_LOCK();
			vol_sync_config(v);
			if (v->bank) DSP_CHECK(msb200_volume_process(v->bank, (int16_t *)m->b_rptr, nsamples), "volume_process");
			DSP_UNLOCK();
			if (v->bank) ms_queue_put(f->outputs[0], m);
			else freemsg(m);
		}
		return;
	}
	while ((m = ms_queue_get(f->inputs[0])) != NULL) {
		int n = (int)((m->b_wptr - m->b_rptr) / 2);
		if (v->bank && n > 0 && n <= VOL_MAX_BLOCK) {
			DSP_LOCK();
			vol_sync_config(v);
			DSP_CHECK(msb200_volume_process(v->bank, (int16_t *)m->b_rptr, n), "volume_process");
			DSP_UNLOCK();
			ms_queue_put(f->outputs[0], m);
		} else {
			freemsg(m); /* no GPU: never forward unprocessed audio as if it had been processed */
		}
	}
}
static int vol_get_state(VolState *v, msb200_volume_state *st) {
	int rc = -1;
	if (!v->bank) return -1;
	DSP_LOCK();
	rc = msb200_volume_get_state(v->bank, 0, st) == MSB200_OK ? 0 : -1;
	DSP_UNLOCK();
	return rc;
}
static int vol_get(MSFilter *f, void *arg) { /* volume_get :121-127: energy in dBm0 */
	msb200_volume_state st;
	if (vol_get_state((VolState *)f->data, &st)) return -1;
	*(float *)arg = st.energy == 0 ? MS_VOLUME_DB_LOWEST : 10 * log10f(st.energy);
	return 0;
}
static int vol_get_linear(MSFilter *f, void *arg) {
	msb200_volume_state st;
	if (vol_get_state((VolState *)f->data, &st)) return -1;
	*(float *)arg = st.energy;
	return 0;
}
static int vol_set_gain(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->static_gain = *(float *)arg;
	v->gain_dirty = TRUE;
	return 0;
}
static int vol_set_db_gain(MSFilter *f, void *arg) { /* pow(10, db/10), sic: msvolume.c:262-268 */
	VolState *v = (VolState *)f->data;
	v->static_gain = (float)pow(10, (*(float *)arg) / 10);
	v->gain_dirty = TRUE;
	return 0;
}
static int vol_get_gain(MSFilter *f, void *arg) {
	*(float *)arg = ((VolState *)f->data)->static_gain;
	return 0;
}
static int vol_get_gain_db(MSFilter *f, void *arg) {
	float g = ((VolState *)f->data)->static_gain;
	*(float *)arg = g == 0 ? MS_VOLUME_DB_LOWEST : 10 * log10f(g);
	return 0;
}
static int vol_set_rate(MSFilter *f, void *arg) {
	((VolState *)f->data)->rate = *(int *)arg;
	return 0;
}
static int vol_set_peer(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->peer = (MSFilter *)arg;
	v->peer_linked = FALSE;
	return 0;
}
static int vol_set_agc(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->agc = *(int *)arg;
	v->dirty = TRUE;
	return 0;
}
static int vol_enable_ng(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->noise_gate = *(bool_t *)arg;
	v->dirty = TRUE;
	return 0;
}
static int vol_set_ng_threshold(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->ng_threshold = *(float *)arg;
	v->dirty = TRUE;
	return 0;
}
static int vol_set_ng_floorgain(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->ng_floorgain = *(float *)arg;
	v->dirty = TRUE;
	return 0;
}
static int vol_remove_dc(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->remove_dc = *(int *)arg;
	v->dirty = TRUE;
	return 0;
}
static int vol_set_ea_threshold(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	float val = *(float *)arg;
	if (val < 0 || val > 1) {
		ms_error("Error: threshold must be in range [0..1]");
		return -1;
	}
	v->ea_thres = val;
	v->dirty = TRUE;
	return 0;
}
static int vol_set_ea_speed(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	float val = *(float *)arg;
	if (val < 0 || val > .5) {
		ms_error("Error: speed must be in range [0..0.5]");
		return -1;
	}
	v->ea_speed = val;
	v->dirty = TRUE;
	return 0;
}
static int vol_set_ea_force(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->ea_force = *(float *)arg;
	v->dirty = TRUE;
	return 0;
}
static int vol_set_ea_sustain(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->ea_sustain = *(int *)arg;
	v->dirty = TRUE;
	return 0;
}
static int vol_set_ea_transmit(MSFilter *f, void *arg) {
	VolState *v = (VolState *)f->data;
	v->ea_transmit = *(float *)arg;
	v->dirty = TRUE;
	return 0;
}
static MSFilterMethod vol_methods[] = {{MS_VOLUME_GET, vol_get},
                                       {MS_VOLUME_GET_LINEAR, vol_get_linear},
                                       {MS_VOLUME_SET_GAIN, vol_set_gain},
                                       {MS_VOLUME_SET_PEER, vol_set_peer},
                                       {MS_VOLUME_SET_EA_THRESHOLD, vol_set_ea_threshold},
                                       {MS_VOLUME_SET_EA_SPEED, vol_set_ea_speed},
                                       {MS_VOLUME_SET_EA_FORCE, vol_set_ea_force},
                                       {MS_VOLUME_SET_EA_SUSTAIN, vol_set_ea_sustain},
                                       {MS_VOLUME_SET_EA_TRANSMIT_THRESHOLD, vol_set_ea_transmit},
                                       {MS_FILTER_SET_SAMPLE_RATE, vol_set_rate},
                                       {MS_VOLUME_ENABLE_AGC, vol_set_agc},
                                       {MS_VOLUME_ENABLE_NOISE_GATE, vol_enable_ng},
                                       {MS_VOLUME_SET_NOISE_GATE_THRESHOLD, vol_set_ng_threshold},
                                       {MS_VOLUME_SET_NOISE_GATE_FLOORGAIN, vol_set_ng_floorgain},
                                       {MS_VOLUME_SET_DB_GAIN, vol_set_db_gain},
                                       {MS_VOLUME_GET_GAIN, vol_get_gain},
                                       {MS_VOLUME_GET_GAIN_DB, vol_get_gain_db},
                                       {MS_VOLUME_REMOVE_DC, vol_remove_dc},
                                       {0, NULL}};
static MSFilterDesc b200_volume_desc = {.id = MS_VOLUME_ID,
                                        .name = "MSVolume",
                                        .text = "B200: controls and measures sound volume (libmsb200dsp)",
                                        .category = MS_FILTER_OTHER,
                                        .ninputs = 1,
                                        .noutputs = 1,
                                        .init = vol_init,
                                        .preprocess = vol_preprocess,
                                        .process = vol_process,
                                        .uninit = vol_uninit,
                                        .methods = vol_methods};

/* ================================================================================================ MSChannelAdapter
 * /root/reference/src/audiofilters/chanadapt.c:45-132 */
typedef struct AdaptState {
	int inputchans, outputchans, sample_rate;
	size_t buffer_size;
	uint8_t *buffer1, *buffer2;
	MSFlowControlledBufferizer input_buffer1, input_buffer2;
} AdaptState;
static void adapt_init(MSFilter *f) {
	AdaptState *s = ms_new0(AdaptState, 1);
	s->inputchans = s->outputchans = 1;
	s->sample_rate = 8000;
	f->data = s;
}
static void adapt_uninit(MSFilter *f) {
	ms_free(f->data);
}
static void adapt_preprocess(MSFilter *f) {
	AdaptState *s = (AdaptState *)f->data;
	DSP_LOCK();
	dsp_ctx();
	DSP_UNLOCK();
	if (s->inputchans == 2 && s->outputchans == 1) {
		s->buffer_size = ((f->ticker->interval * s->sample_rate) / 1000) * 2;
		s->buffer1 = ms_new(uint8_t, s->buffer_size);
		s->buffer2 = ms_new(uint8_t, s->buffer_size);
		ms_flow_controlled_bufferizer_init(&s->input_buffer1, f, s->sample_rate, 1);
		ms_flow_controlled_bufferizer_set_drop_method(&s->input_buffer1, MSFlowControlledBufferizerImmediateDrop);
		ms_flow_controlled_bufferizer_set_max_size_ms(&s->input_buffer1, f->ticker->interval * 2);
		ms_flow_controlled_bufferizer_init(&s->input_buffer2, f, s->sample_rate, 1);
		ms_flow_controlled_bufferizer_set_drop_method(&s->input_buffer2, MSFlowControlledBufferizerImmediateDrop);
		ms_flow_controlled_bufferizer_set_max_size_ms(&s->input_buffer2, f->ticker->interval * 2);
	}
}
static void adapt_postprocess(MSFilter *f) {
	AdaptState *s = (AdaptState *)f->data;
	if (s->inputchans == 2 && s->outputchans == 1) {
		ms_flow_controlled_bufferizer_uninit(&s->input_buffer1);
		ms_flow_controlled_bufferizer_uninit(&s->input_buffer2);
		ms_free(s->buffer1);
		ms_free(s->buffer2);
		s->buffer1 = s->buffer2 = NULL;
	}
}
static void adapt_process(MSFilter *f) {
	AdaptState *s = (AdaptState *)f->data;
	if (f->inputs[0] != NULL && f->inputs[1] != NULL) {
		size_t a1, a2;
		ms_flow_controlled_bufferizer_put_from_queue(&s->input_buffer1, f->inputs[0]);
		ms_flow_controlled_bufferizer_put_from_queue(&s->input_buffer2, f->inputs[1]);
		a1 = ms_flow_controlled_bufferizer_get_avail(&s->input_buffer1);
		a2 = ms_flow_controlled_bufferizer_get_avail(&s->input_buffer2);
		if (a1 >= s->buffer_size || a2 >= s->buffer_size) {
			mblk_t *om = allocb(s->buffer_size * 2, 0);
			int frames = (int)(s->buffer_size / 2);
			ms_flow_controlled_bufferizer_read(&s->input_buffer1, s->buffer1, s->buffer_size);
			ms_flow_controlled_bufferizer_read(&s->input_buffer2, s->buffer2, s->buffer_size);
			DSP_LOCK();
			if (g_ctx)
				DSP_CHECK(msb200_chanadapt_process(g_ctx, MSB200_CHAN_2MONO_TO_STEREO, 1, frames,
				                                   a1 >= s->buffer_size ? (int16_t *)s->buffer1 : NULL,
				                                   a2 >= s->buffer_size ? (int16_t *)s->buffer2 : NULL, (int16_t *)om->b_wptr),
				          "chanadapt");
			DSP_UNLOCK();
			om->b_wptr += s->buffer_size * 2;
			if (g_ctx) ms_queue_put(f->outputs[0], om);
			else freemsg(om);
		}
		return;
	}
	{
		mblk_t *im;
		while ((im = ms_queue_get(f->inputs[0])) != NULL) {
			if (s->inputchans == s->outputchans) {
				ms_queue_put(f->outputs[0], im);
			} else {
				int to_stereo = s->outputchans == 2;
				size_t insz = msgdsize(im);
				size_t outsz = to_stereo ? insz * 2 : insz / 2;
				int frames = (int)(to_stereo ? insz / 2 : insz / 4);
				mblk_t *om = allocb(outsz, 0);
				DSP_LOCK();
				if (g_ctx && frames > 0)
					DSP_CHECK(msb200_chanadapt_process(g_ctx, to_stereo ? MSB200_CHAN_MONO_TO_STEREO : MSB200_CHAN_STEREO_TO_MONO,
					                                   1, frames, (int16_t *)im->b_rptr, NULL, (int16_t *)om->b_wptr),
					          "chanadapt");
				DSP_UNLOCK();
				om->b_wptr += outsz;
				if (g_ctx) ms_queue_put(f->outputs[0], om);
				else freemsg(om);
				freemsg(im);
			}
		}
	}
}
static int adapt_set_sr(MSFilter *f, void *data) {
	((AdaptState *)f->data)->sample_rate = *(int *)data;
	return 0;
}
static int adapt_get_sr(MSFilter *f, void *data) {
	*(int *)data = ((AdaptState *)f->data)->sample_rate;
	return 0;
}
static int adapt_set_nchannels(MSFilter *f, void *data) {
	((AdaptState *)f->data)->inputchans = *(int *)data;
	return 0;
}
static int adapt_get_nchannels(MSFilter *f, void *data) {
	*(int *)data = ((AdaptState *)f->data)->inputchans;
	return 0;
}
static int adapt_set_out_nchannels(MSFilter *f, void *data) {
	((AdaptState *)f->data)->outputchans = *(int *)data;
	return 0;
}
static int adapt_get_out_nchannels(MSFilter *f, void *data) {
	*(int *)data = ((AdaptState *)f->data)->outputchans;
	return 0;
}
static MSFilterMethod adapt_methods[] = {{MS_FILTER_SET_SAMPLE_RATE, adapt_set_sr},
                                         {MS_FILTER_GET_SAMPLE_RATE, adapt_get_sr},
                                         {MS_FILTER_SET_NCHANNELS, adapt_set_nchannels},
                                         {MS_FILTER_GET_NCHANNELS, adapt_get_nchannels},
                                         {MS_CHANNEL_ADAPTER_SET_OUTPUT_NCHANNELS, adapt_set_out_nchannels},
                                         {MS_CHANNEL_ADAPTER_GET_OUTPUT_NCHANNELS, adapt_get_out_nchannels},
                                         {0, NULL}};
static MSFilterDesc b200_channel_adapter_desc = {.id = MS_CHANNEL_ADAPTER_ID,
                                                 .name = "MSChannelAdapter",
                                                 .text = "B200: mono/stereo channel adaptation (libmsb200dsp)",
                                                 .category = MS_FILTER_OTHER,
                                                 .ninputs = 2,
                                                 .noutputs = 1,
                                                 .init = adapt_init,
                                                 .preprocess = adapt_preprocess,
                                                 .process = adapt_process,
                                                 .postprocess = adapt_postprocess,
                                                 .uninit = adapt_uninit,
                                                 .methods = adapt_methods,
                                                 .flags = MS_FILTER_IS_PUMP};

/* ================================================================================================ MSEqualizer
 * /root/reference/src/audiofilters/equalizer.c:279-342 */
typedef struct EqCmd {
	float f, g, w;
} EqCmd;
typedef struct EqState {
	int rate;
	bool_t active;
	msb200_equalizer *bank;
	int bank_rate;
	EqCmd cmds[128]; /* gains set before the bank exists are replayed in order */
	int ncmds;
} EqState;
static void eq_ensure_bank(EqState *s) { /* DSP lock held */
	int i;
	if (!dsp_ctx()) return;
	if (s->bank && s->bank_rate == s->rate) return;
	msb200_equalizer_destroy(s->bank);
	s->bank = NULL;
	DSP_CHECK(msb200_equalizer_create(g_ctx, 1, s->rate, 8192, &s->bank), "equalizer_create");
	s->bank_rate = s->rate;
	for (i = 0; s->bank && i < s->ncmds; ++i)
		msb200_equalizer_set_gain(s->bank, 0, s->cmds[i].f, s->cmds[i].g, s->cmds[i].w);
	if (s->bank) msb200_equalizer_set_active(s->bank, 0, s->active);
}
static void eq_init(MSFilter *f) {
	EqState *s = ms_new0(EqState, 1);
	s->rate = 8000;
	s->active = TRUE;
	f->data = s;
}
static void eq_uninit(MSFilter *f) {
	EqState *s = (EqState *)f->data;
	DSP_LOCK();
	msb200_equalizer_destroy(s->bank);
	DSP_UNLOCK();
	ms_free(s);
}
static void eq_preprocess(MSFilter *f) {
	DSP_LOCK();
	eq_ensure_bank((EqState *)f->data);
	DSP_UNLOCK();
}
static void eq_process(MSFilter *f) {
	EqState *s = (EqState *)f->data;
	mblk_t *m;
	while ((m = ms_queue_get(f->inputs[0])) != NULL) {
		int n = (int)((m->b_wptr - m->b_rptr) / 2);
		if (s->active && n > 0) {
			DSP_LOCK();
			eq_ensure_bank(s);
			if (s->bank) DSP_CHECK(msb200_equalizer_process(s->bank, (int16_t *)m->b_rptr, n), "equalizer_process");
			DSP_UNLOCK();
			if (!s->bank) {
				freemsg(m);
				continue;
			}
		}
		ms_queue_put(f->outputs[0], m);
	}
}
static int eq_set_gain(MSFilter *f, void *data) {
	EqState *s = (EqState *)f->data;
	MSEqualizerGain *d = (MSEqualizerGain *)data;
	if (s->ncmds < 128) {
		s->cmds[s->ncmds].f = d->frequency;
		s->cmds[s->ncmds].g = d->gain;
		s->cmds[s->ncmds].w = d->width;
		s->ncmds++;
	}
	if (s->bank && s->bank_rate == s->rate) {
		DSP_LOCK();
		msb200_equalizer_set_gain(s->bank, 0, d->frequency, d->gain, d->width);
		DSP_UNLOCK();
	}
	return 0;
}
static int eq_get_gain(MSFilter *f, void *data) {
	EqState *s = (EqState *)f->data;
	MSEqualizerGain *d = (MSEqualizerGain *)data;
	DSP_LOCK();
	eq_ensure_bank(s);
	if (s->bank) msb200_equalizer_get_gain(s->bank, 0, d->frequency, &d->gain);
	DSP_UNLOCK();
	d->width = 0;
	return s->bank ? 0 : -1;
}
static int eq_set_rate(MSFilter *f, void *data) { /* equalizer_rate_update resets the gain table (:57-79) */
	EqState *s = (EqState *)f->data;
	s->rate = *(int *)data;
	s->ncmds = 0;
	return 0;
}
static int eq_set_active(MSFilter *f, void *data) {
	EqState *s = (EqState *)f->data;
	s->active = *(bool_t *)data;
	if (s->bank) {
		DSP_LOCK();
		msb200_equalizer_set_active(s->bank, 0, s->active);
		DSP_UNLOCK();
	}
	return 0;
}
static int eq_get_nfreqs(MSFilter *f, void *data) {
	int rate = ((EqState *)f->data)->rate;
	*(int *)data = (rate < 16000 ? 128 : (rate < 32000 ? 256 : 512)) / 2;
	return 0;
}
static MSFilterMethod eq_methods[] = {{MS_EQUALIZER_SET_GAIN, eq_set_gain},
                                      {MS_EQUALIZER_GET_GAIN, eq_get_gain},
                                      {MS_EQUALIZER_SET_ACTIVE, eq_set_active},
                                      {MS_FILTER_SET_SAMPLE_RATE, eq_set_rate},
                                      {MS_EQUALIZER_GET_NUM_FREQUENCIES, eq_get_nfreqs},
                                      {0, NULL}};
static MSFilterDesc b200_equalizer_desc = {.id = MS_EQUALIZER_ID,
                                           .name = "MSEqualizer",
                                           .text = "B200: parametric sound equalizer (libmsb200dsp)",
                                           .category = MS_FILTER_OTHER,
                                           .ninputs = 1,
                                           .noutputs = 1,
                                           .init = eq_init,
                                           .preprocess = eq_preprocess,
                                           .process = eq_process,
                                           .uninit = eq_uninit,
                                           .methods = eq_methods};

/* ================================================================================================ MSResample
 * /root/reference/src/audiofilters/msresample.c:122-233 */
typedef struct RsState {
	uint32_t ts, input_rate, output_rate;
	int in_nchannels, out_nchannels;
	msb200_resample *bank;
	uint32_t bank_in, bank_out;
	int bank_ch;
} RsState;
#define RS_MAX_FRAMES 8192
static void rs_init(MSFilter *f) {
	RsState *s = ms_new0(RsState, 1);
	s->input_rate = 8000;
	s->output_rate = 16000;
	s->in_nchannels = s->out_nchannels = 1;
	f->data = s;
}
static void rs_uninit(MSFilter *f) {
	RsState *s = (RsState *)f->data;
	DSP_LOCK();
	msb200_resample_destroy(s->bank);
	DSP_UNLOCK();
	ms_free(s);
}
static void rs_ensure_bank(RsState *s) { /* DSP lock held; mirrors the lazy (re)creation of the speex handle :138-148 */
	if (!dsp_ctx()) return;
	if (s->bank && s->bank_in == s->input_rate && s->bank_out == s->output_rate && s->bank_ch == s->in_nchannels) return;
	msb200_resample_destroy(s->bank);
	s->bank = NULL;
	if (s->input_rate == s->output_rate) return;
	DSP_CHECK(msb200_resample_create(g_ctx, 1, (int)s->input_rate, (int)s->output_rate, s->in_nchannels, RS_MAX_FRAMES, &s->bank),
	          "resample_create");
	s->bank_in = s->input_rate;
	s->bank_out = s->output_rate;
	s->bank_ch = s->in_nchannels;
}
static mblk_t *rs_channel_adapt(int in_ch, int out_ch, mblk_t *im) { /* resample_channel_adapt :87-100 */
	size_t msgsize = msgdsize(im) * (size_t)out_ch / (size_t)in_ch;
	mblk_t *om = allocb(msgsize, 0);
	int i;
	for (; im->b_rptr < im->b_wptr; im->b_rptr += sizeof(int16_t) * in_ch, om->b_wptr += sizeof(int16_t) * out_ch)
		for (i = 0; i < out_ch; ++i)
			((int16_t *)om->b_wptr)[i] = *(int16_t *)im->b_rptr;
	mblk_meta_copy(im, om);
	return om;
}
static void rs_process(MSFilter *f) {
	RsState *s = (RsState *)f->data;
	mblk_t *im;
	if (s->output_rate == s->input_rate) {
		while ((im = ms_queue_get(f->inputs[0])) != NULL) {
			if (s->out_nchannels == s->in_nchannels) {
				ms_queue_put(f->outputs[0], im);
			} else {
				ms_queue_put(f->outputs[0], rs_channel_adapt(s->in_nchannels, s->out_nchannels, im));
				freemsg(im);
			}
		}
		return;
	}
	ms_filter_lock(f);
	while ((im = ms_queue_get(f->inputs[0])) != NULL) {
		int inlen = (int)((im->b_wptr - im->b_rptr) / (2 * s->in_nchannels));
		int outcap = (int)(((uint32_t)inlen * s->output_rate) / s->input_rate) + 1;
		int outlen = 0;
		mblk_t *om = allocb((size_t)outcap * 2 * (size_t)s->in_nchannels, 0);
		mblk_meta_copy(im, om);
		DSP_LOCK();
		rs_ensure_bank(s);
		if (s->bank && inlen > 0 && inlen <= RS_MAX_FRAMES)
			DSP_CHECK(msb200_resample_process(s->bank, (const int16_t *)im->b_rptr, inlen, (int16_t *)om->b_wptr, outcap, &outlen),
			          "resample_process");
		DSP_UNLOCK();
		if (!s->bank) {
			freemsg(om);
			freemsg(im);
			continue;
		}
		om->b_wptr += (size_t)outlen * 2 * (size_t)s->in_nchannels;
		mblk_set_timestamp_info(om, s->ts);
		s->ts += (uint32_t)outlen;
		if (s->out_nchannels != s->in_nchannels) {
			ms_queue_put(f->outputs[0], rs_channel_adapt(s->in_nchannels, s->out_nchannels, om));
			freemsg(om);
		} else {
			ms_queue_put(f->outputs[0], om);
		}
		freemsg(im);
	}
	ms_filter_unlock(f);
}
static void rs_preprocess(MSFilter *f) {
	DSP_LOCK();
	rs_ensure_bank((RsState *)f->data);
	DSP_UNLOCK();
}
static int rs_set_sr(MSFilter *f, void *arg) {
	ms_filter_lock(f);
	((RsState *)f->data)->input_rate = *(unsigned int *)arg;
	ms_filter_unlock(f);
	return 0;
}
static int rs_set_out_sr(MSFilter *f, void *arg) {
	ms_filter_lock(f);
	((RsState *)f->data)->output_rate = *(unsigned int *)arg;
	ms_filter_unlock(f);
	return 0;
}
static int rs_set_in_nch(MSFilter *f, void *arg) {
	ms_filter_lock(f);
	((RsState *)f->data)->in_nchannels = *(int *)arg;
	ms_filter_unlock(f);
	return 0;
}
static int rs_set_out_nch(MSFilter *f, void *arg) {
	ms_filter_lock(f);
	((RsState *)f->data)->out_nchannels = *(int *)arg;
	ms_filter_unlock(f);
	return 0;
}
static MSFilterMethod rs_methods[] = {{MS_FILTER_SET_SAMPLE_RATE, rs_set_sr},
                                      {MS_FILTER_SET_OUTPUT_SAMPLE_RATE, rs_set_out_sr},
                                      {MS_FILTER_SET_NCHANNELS, rs_set_in_nch},
                                      {MS_FILTER_SET_OUTPUT_NCHANNELS, rs_set_out_nch},
                                      {0, NULL}};
static MSFilterDesc b200_resample_desc = {.id = MS_RESAMPLE_ID,
                                          .name = "MSResample",
                                          .text = "B200: audio resampler (libmsb200dsp)",
                                          .category = MS_FILTER_OTHER,
                                          .ninputs = 1,
                                          .noutputs = 1,
                                          .init = rs_init,
                                          .preprocess = rs_preprocess,
                                          .process = rs_process,
                                          .uninit = rs_uninit,
                                          .methods = rs_methods};

/* ================================================================================================ MSSpeexEC
 * host logic restated from /root/reference/src/audiofilters/speexec.c:171-216 (configuration), :223-305 (process:
 * reference/echo bufferizers, silence injection on underrun), :308-391 (methods) */
typedef struct EcState {
	msb200_aec *bank;
	MSBufferizer delayed_ref;
	MSFlowControlledBufferizer ref;
	MSBufferizer echo;
	int framesize, framesize_at_8000, samplerate, delay_ms, tail_length_ms, nominal_ref_samples;
	char *state_str;
	bool_t echostarted, bypass_mode, using_zeroes;
} EcState;
static void ec_configure_fcb(EcState *s) {
	ms_flow_controlled_bufferizer_set_samplerate(&s->ref, s->samplerate);
	ms_flow_controlled_bufferizer_set_max_size_ms(&s->ref, s->delay_ms);
	ms_flow_controlled_bufferizer_set_granularity_ms(&s->ref, (s->framesize * 1000) / s->samplerate);
}
static void ec_init(MSFilter *f) {
	EcState *s = ms_new0(EcState, 1);
	s->samplerate = 8000;
	ms_bufferizer_init(&s->delayed_ref);
	ms_bufferizer_init(&s->echo);
	ms_flow_controlled_bufferizer_init(&s->ref, f, s->samplerate, 1);
	s->delay_ms = 0;
	s->tail_length_ms = 250;
	s->framesize_at_8000 = 64;
	s->framesize = 64;
	f->data = s;
}
static void ec_uninit(MSFilter *f) {
	EcState *s = (EcState *)f->data;
	if (s->state_str) ms_free(s->state_str);
	ms_bufferizer_uninit(&s->delayed_ref);
	ms_bufferizer_uninit(&s->echo);
	ms_flow_controlled_bufferizer_uninit(&s->ref);
	ms_free(s);
}
static void ec_preprocess(MSFilter *f) {
	EcState *s = (EcState *)f->data;
	int delay_samples;
	mblk_t *m;
	s->echostarted = FALSE;
	s->framesize = msb200_aec_frame_size_for_rate(s->samplerate, s->framesize_at_8000);
	delay_samples = s->delay_ms * s->samplerate / 1000;
	ms_message("Initializing B200 echo canceler with framesize=%i, filterlength=%i, delay_samples=%i", s->framesize,
	           (s->tail_length_ms * s->samplerate) / 1000, delay_samples);
	DSP_LOCK();
	if (dsp_ctx()) DSP_CHECK(msb200_aec_create(g_ctx, 1, s->samplerate, s->tail_length_ms, s->framesize_at_8000, &s->bank), "aec_create");
	DSP_UNLOCK();
	m = allocb((size_t)delay_samples * 2, 0);
	m->b_wptr += delay_samples * 2;
	ms_bufferizer_put(&s->delayed_ref, m);
	s->nominal_ref_samples = delay_samples;
	ec_configure_fcb(s);
}
static void ec_postprocess(MSFilter *f) {
	EcState *s = (EcState *)f->data;
	ms_bufferizer_flush(&s->delayed_ref);
	ms_bufferizer_flush(&s->echo);
	ms_flow_controlled_bufferizer_flush(&s->ref);
	DSP_LOCK();
	msb200_aec_destroy(s->bank);
	DSP_UNLOCK();
	s->bank = NULL;
}
static void ec_process(MSFilter *f) {
	EcState *s = (EcState *)f->data;
	int nbytes = s->framesize * 2;
	mblk_t *refm;
	uint8_t *ref, *echo;
	if (s->bypass_mode) {
		while ((refm = ms_queue_get(f->inputs[0])) != NULL)
			ms_queue_put(f->outputs[0], refm);
		while ((refm = ms_queue_get(f->inputs[1])) != NULL)
			ms_queue_put(f->outputs[1], refm);
		return;
	}
	if (f->inputs[0] != NULL) {
		if (s->echostarted) {
			while ((refm = ms_queue_get(f->inputs[0])) != NULL) {
				mblk_t *cp = dupmsg(refm);
				ms_bufferizer_put(&s->delayed_ref, cp);
				ms_flow_controlled_bufferizer_put(&s->ref, refm);
			}
		} else {
			ms_warning("Getting reference signal but no echo to synchronize on.");
			ms_queue_flush(f->inputs[0]);
		}
	}
	ms_bufferizer_put_from_queue(&s->echo, f->inputs[1]);
	ref = (uint8_t *)alloca((size_t)nbytes);
	echo = (uint8_t *)alloca((size_t)nbytes);
	while ((int)ms_bufferizer_read(&s->echo, echo, (size_t)nbytes) == nbytes) {
		mblk_t *oecho = allocb((size_t)nbytes, 0);
		if (!s->echostarted) s->echostarted = TRUE;
		if ((int)ms_bufferizer_get_avail(&s->delayed_ref) < ((s->nominal_ref_samples * 2) + nbytes)) {
			refm = allocb((size_t)nbytes, 0);
			memset(refm->b_wptr, 0, (size_t)nbytes);
			refm->b_wptr += nbytes;
			ms_bufferizer_put(&s->delayed_ref, refm);
			ms_queue_put(f->outputs[0], dupmsg(refm));
			if (!s->using_zeroes) {
				ms_warning("Not enough ref samples, using zeroes");
				s->using_zeroes = TRUE;
			}
		} else {
			if (s->using_zeroes) {
				ms_message("Samples are back.");
				s->using_zeroes = FALSE;
			}
			refm = allocb((size_t)nbytes, 0);
			if (ms_flow_controlled_bufferizer_read(&s->ref, refm->b_wptr, (size_t)nbytes) == 0) ms_fatal("Should never happen");
			refm->b_wptr += nbytes;
			ms_queue_put(f->outputs[0], refm);
		}
		if (ms_bufferizer_read(&s->delayed_ref, ref, (size_t)nbytes) == 0) ms_fatal("Should never happen");
		/* speex_echo_cancellation + speex_preprocess_run for this frame, on the GPU */
		DSP_LOCK();
		if (s->bank) DSP_CHECK(msb200_aec_process(s->bank, (int16_t *)echo, (int16_t *)ref, (int16_t *)oecho->b_wptr, 1), "aec_process");
		DSP_UNLOCK();
		if (!s->bank) {
			freemsg(oecho);
			continue;
		}
		oecho->b_wptr += nbytes;
		ms_queue_put(f->outputs[1], oecho);
	}
}
static int ec_set_sr(MSFilter *f, void *arg) {
	EcState *s = (EcState *)f->data;
	s->samplerate = *(int *)arg;
	ec_configure_fcb(s);
	return 0;
}
static int ec_get_sr(MSFilter *f, void *arg) {
	*(int *)arg = ((EcState *)f->data)->samplerate;
	return 0;
}
static int ec_set_framesize(MSFilter *f, void *arg) {
	((EcState *)f->data)->framesize_at_8000 = *(int *)arg;
	return 0;
}
static int ec_set_delay(MSFilter *f, void *arg) {
	EcState *s = (EcState *)f->data;
	s->delay_ms = *(int *)arg;
	ec_configure_fcb(s);
	return 0;
}
static int ec_get_delay(MSFilter *f, void *arg) {
	*(int *)arg = ((EcState *)f->data)->delay_ms;
	return 0;
}
static int ec_set_tail_length(MSFilter *f, void *arg) {
	EcState *s = (EcState *)f->data;
	s->tail_length_ms = *(int *)arg;
	ec_configure_fcb(s);
	return 0;
}
static int ec_set_bypass(MSFilter *f, void *arg) {
	((EcState *)f->data)->bypass_mode = *(bool_t *)arg;
	return 0;
}
static int ec_get_bypass(MSFilter *f, void *arg) {
	*(bool_t *)arg = ((EcState *)f->data)->bypass_mode;
	return 0;
}
static int ec_set_state(MSFilter *f, void *arg) {
	EcState *s = (EcState *)f->data;
	if (s->state_str) ms_free(s->state_str);
	s->state_str = ms_strdup((const char *)arg);
	return 0;
}
static int ec_get_state(MSFilter *f, void *arg) {
	*(char **)arg = ((EcState *)f->data)->state_str;
	return 0;
}
static MSFilterMethod ec_methods[] = {{MS_FILTER_SET_SAMPLE_RATE, ec_set_sr},
                                      {MS_FILTER_GET_SAMPLE_RATE, ec_get_sr},
                                      {MS_ECHO_CANCELLER_SET_TAIL_LENGTH, ec_set_tail_length},
                                      {MS_ECHO_CANCELLER_SET_DELAY, ec_set_delay},
                                      {MS_ECHO_CANCELLER_SET_FRAMESIZE, ec_set_framesize},
                                      {MS_ECHO_CANCELLER_SET_BYPASS_MODE, ec_set_bypass},
                                      {MS_ECHO_CANCELLER_GET_BYPASS_MODE, ec_get_bypass},
                                      {MS_ECHO_CANCELLER_GET_STATE_STRING, ec_get_state},
                                      {MS_ECHO_CANCELLER_SET_STATE_STRING, ec_set_state},
                                      {MS_ECHO_CANCELLER_GET_DELAY, ec_get_delay},
                                      {0, NULL}};
static MSFilterDesc b200_speex_ec_desc = {.id = MS_SPEEX_EC_ID,
                                          .name = "MSSpeexEC",
                                          .text = "B200: MDF echo canceller + preprocessor (libmsb200dsp)",
                                          .category = MS_FILTER_OTHER,
                                          .ninputs = 2,
                                          .noutputs = 2,
                                          .init = ec_init,
                                          .preprocess = ec_preprocess,
                                          .process = ec_process,
                                          .postprocess = ec_postprocess,
                                          .uninit = ec_uninit,
                                          .methods = ec_methods};

/* ================================================================================================ MSScalerDesc
 * the second drop-in boundary (/root/reference/include/mediastreamer2/msvideo.h:473-492): installed with
 * ms_video_set_scaler_impl() so that the reference's own MSPixConv / MSSizeConv / display filters scale on the GPU. */
typedef struct B200ScalerCtx {
	msb200_scaler *sc;
	int src_w, src_h, dst_w, dst_h, src_fmt, dst_fmt;
	uint8_t *src_pack, *dst_pack;
} B200ScalerCtx;
static int pixfmt_to_b200(MSPixFmt fmt) {
	switch (fmt) {
		case MS_YUV420P: return MSB200_PIX_YUV420P;
		case MS_YUYV: return MSB200_PIX_YUYV;
		case MS_YUY2: return MSB200_PIX_YUY2;
		case MS_UYVY: return MSB200_PIX_UYVY;
		case MS_RGB24: return MSB200_PIX_RGB24;
		case MS_RGB24_REV: return MSB200_PIX_RGB24_REV;
		default: return -1;
	}
}
static MSScalerContext *b200_scaler_create(int src_w, int src_h, MSPixFmt src_fmt, int dst_w, int dst_h, MSPixFmt dst_fmt, int flags) {
	B200ScalerCtx *c;
	int sf = pixfmt_to_b200(src_fmt), df = pixfmt_to_b200(dst_fmt);
	(void)flags;
	if ((sf != MSB200_PIX_YUV420P && sf != MSB200_PIX_YUYV && sf != MSB200_PIX_YUY2 && sf != MSB200_PIX_UYVY) || df < 0) {
		ms_error("msb200 scaler: unsupported conversion %s -> %s", ms_pix_fmt_to_string(src_fmt), ms_pix_fmt_to_string(dst_fmt));
		return NULL;
	}
	c = ms_new0(B200ScalerCtx, 1);
	c->src_w = src_w; c->src_h = src_h; c->dst_w = dst_w; c->dst_h = dst_h; c->src_fmt = sf; c->dst_fmt = df;
	DSP_LOCK();
	if (dsp_ctx()) DSP_CHECK(msb200_scaler_create(g_ctx, src_w, src_h, sf, dst_w, dst_h, df, &c->sc), "scaler_create");
	DSP_UNLOCK();
	if (!c->sc) {
		ms_free(c);
		return NULL;
	}
	c->src_pack = (uint8_t *)ms_malloc(msb200_scaler_src_frame_bytes(c->sc));
	c->dst_pack = (uint8_t *)ms_malloc(msb200_scaler_dst_frame_bytes(c->sc));
	return (MSScalerContext *)c;
}
static void pack_plane(uint8_t *dst, const uint8_t *src, int stride, int w, int h) {
	int y;
	for (y = 0; y < h; ++y)
		memcpy(dst + (size_t)y * w, src + (size_t)y * stride, (size_t)w);
}
static int b200_scaler_process(MSScalerContext *ctx, uint8_t *src[], int src_strides[], uint8_t *dst[], int dst_strides[]) {
	B200ScalerCtx *c = (B200ScalerCtx *)ctx;
	int rc, cw = (c->src_w + 1) / 2, ch = (c->src_h + 1) / 2, y;
	uint8_t *p = c->src_pack;
	if (c->src_fmt != MSB200_PIX_YUV420P) { /* packed 4:2:2: one plane of 2 bytes per pixel (msvideo.c:120-156) */
		pack_plane(p, src[0], src_strides[0], c->src_w * 2, c->src_h);
	} else {
		pack_plane(p, src[0], src_strides[0], c->src_w, c->src_h);
		p += (size_t)c->src_w * c->src_h;
		pack_plane(p, src[1], src_strides[1], cw, ch);
		p += (size_t)cw * ch;
		pack_plane(p, src[2], src_strides[2], cw, ch);
	}
	DSP_LOCK();
	rc = msb200_scaler_process(c->sc, 1, c->src_pack, c->dst_pack);
	DSP_UNLOCK();
	if (rc != MSB200_OK) {
		ms_error("msb200 scaler: %s", msb200_last_error());
		return -1;
	}
	p = c->dst_pack;
	if (c->dst_fmt == MSB200_PIX_YUV420P) {
		int dcw = (c->dst_w + 1) / 2, dch = (c->dst_h + 1) / 2;
		for (y = 0; y < c->dst_h; ++y) memcpy(dst[0] + (size_t)y * dst_strides[0], p + (size_t)y * c->dst_w, (size_t)c->dst_w);
		p += (size_t)c->dst_w * c->dst_h;
		for (y = 0; y < dch; ++y) memcpy(dst[1] + (size_t)y * dst_strides[1], p + (size_t)y * dcw, (size_t)dcw);
		p += (size_t)dcw * dch;
		for (y = 0; y < dch; ++y) memcpy(dst[2] + (size_t)y * dst_strides[2], p + (size_t)y * dcw, (size_t)dcw);
	} else {
		for (y = 0; y < c->dst_h; ++y) memcpy(dst[0] + (size_t)y * dst_strides[0], p + (size_t)y * c->dst_w * 3, (size_t)c->dst_w * 3);
	}
	return 0;
}
static void b200_scaler_free(MSScalerContext *ctx) {
	B200ScalerCtx *c = (B200ScalerCtx *)ctx;
	DSP_LOCK();
	msb200_scaler_destroy(c->sc);
	DSP_UNLOCK();
	ms_free(c->src_pack);
	ms_free(c->dst_pack);
	ms_free(c);
}
static MSScalerDesc b200_scaler_desc = {b200_scaler_create, b200_scaler_process, b200_scaler_free};

/* ================================================================================================ entry point */
__attribute__((visibility("default"))) void libmsb200filters_init(MSFactory *factory) {
	ms_factory_register_filter(factory, &b200_audio_mixer_desc);
	ms_factory_register_filter(factory, &b200_volume_desc);
	ms_factory_register_filter(factory, &b200_channel_adapter_desc);
	ms_factory_register_filter(factory, &b200_equalizer_desc);
	ms_factory_register_filter(factory, &b200_resample_desc);
	ms_factory_register_filter(factory, &b200_speex_ec_desc);
	if (getenv("MSB200_INSTALL_SCALER")) ms_video_set_scaler_impl(&b200_scaler_desc);
	ms_message("libmsb200filters: B200 DSP filters registered (MSAudioMixer, MSVolume, MSChannelAdapter, MSEqualizer, "
	           "MSResample, MSSpeexEC%s)", getenv("MSB200_INSTALL_SCALER") ? ", MSScaler" : "");
}
/* also exported so that a host can install the scaler explicitly */
__attribute__((visibility("default"))) MSScalerDesc *msb200_ms_scaler_desc(void) {
	return &b200_scaler_desc;
}
