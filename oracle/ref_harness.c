/* oracle/ref_harness.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Links against the UNMODIFIED reference runtime and in-tree filters (compiled from /root/reference by
 * oracle/Makefile into oracle/_ref/libms2ref.so) and exposes a tiny scripting surface for ctypes:
 * create a factory (optionally loading the msb200dsp plugin through the reference's own dlopen loader),
 * create filters by name, call methods, link, and run a real MSTicker for an exact number of ticks with a
 * gated tick function (virtual time, no sleeping) so tests are deterministic.
 *
 * Two helper filters live here (ours, not the reference's): "HarnessSource" feeds scripted blocks per tick,
 * "HarnessSink" records every block it receives. They play the role of MSFilePlayer/MSFileRec in the
 * reference's own tests (tester/mediastreamer2_basic_audio_tester.c:546-661).
 */
#include "mediastreamer2/msaudiomixer.h"
#include "mediastreamer2/mschanadapter.h"
#include "mediastreamer2/msequalizer.h"
#include "mediastreamer2/flowcontrol.h"
#include "mediastreamer2/msgenericplc.h"
#include "mediastreamer2/msfactory.h"
#include "mediastreamer2/msfilter.h"
#include "mediastreamer2/msinterfaces.h"
#include "mediastreamer2/msticker.h"
#include "mediastreamer2/msvideo.h"
#include "mediastreamer2/msvolume.h"
#include "msb200_ms2.h" /* the plugin's extension methods (include/), so that tests can name them */

#include <pthread.h>
#include <string.h>

/* ---------------------------------------------------------------- HarnessSource / HarnessSink */
typedef struct HBlock {
	struct HBlock *next;
	int tick;
	int nbytes;
	int w, h;       /* > 0: an I420 frame, delivered the way cameras / decoders do: ms_yuv_buf_alloc (16-byte header + planes) */
	int has_ts;     /* set the mblk's timestamp */
	uint32_t ts;
	uint8_t data[1];
} HBlock;

typedef struct HSource {
	HBlock *head, *tail;
	int tick; /* ticks elapsed since preprocess */
} HSource;

static void hsrc_init(MSFilter *f) {
	f->data = ms_new0(HSource, 1);
}
static void hsrc_preprocess(MSFilter *f) {
	((HSource *)f->data)->tick = 0;
}
static void hsrc_process(MSFilter *f) {
	HSource *s = (HSource *)f->data;
	while (s->head && s->head->tick <= s->tick) {
		HBlock *b = s->head;
		mblk_t *m;
		if (b->w > 0) {
			YuvBuf yb;
			m = ms_yuv_buf_alloc(&yb, b->w, b->h); /* src/voip/msvideo.c:158-172 */
			memcpy(yb.planes[0], b->data, (size_t)b->nbytes);
		} else {
			m = allocb((size_t)b->nbytes, 0);
			memcpy(m->b_wptr, b->data, (size_t)b->nbytes);
			m->b_wptr += b->nbytes;
		}
		if (b->has_ts) mblk_set_timestamp_info(m, b->ts);
		if (f->outputs[0]) ms_queue_put(f->outputs[0], m);
		else freemsg(m);
		s->head = b->next;
		if (!s->head) s->tail = NULL;
		ms_free(b);
	}
	s->tick++;
}
static void hsrc_uninit(MSFilter *f) {
	HSource *s = (HSource *)f->data;
	while (s->head) {
		HBlock *b = s->head;
		s->head = b->next;
		ms_free(b);
	}
	ms_free(s);
}
static MSFilterDesc harness_source_desc = {.id = MS_FILTER_PLUGIN_ID,
                                           .name = "HarnessSource",
                                           .text = "scripted block source",
                                           .category = MS_FILTER_OTHER,
                                           .ninputs = 0,
                                           .noutputs = 1,
                                           .init = hsrc_init,
                                           .preprocess = hsrc_preprocess,
                                           .process = hsrc_process,
                                           .uninit = hsrc_uninit};

typedef struct HSink {
	uint8_t *buf;
	size_t len, cap;
	int *sizes; /* (tick, nbytes, timestamp) triples */
	int *dims;  /* (w, h) of the 16-byte video header below b_rptr (msvideo.c:79-83), (0, 0) when the block has none */
	int nblocks, blocks_cap;
	int tick;
	int discard; /* timing runs: count the blocks and bytes, keep nothing (no growing buffer inside the timed ticks) */
} HSink;

static void hsink_init(MSFilter *f) {
	f->data = ms_new0(HSink, 1);
}
static void hsink_preprocess(MSFilter *f) {
	((HSink *)f->data)->tick = 0;
}
static void hsink_process(MSFilter *f) {
	HSink *s = (HSink *)f->data;
	mblk_t *m;
	while ((m = ms_queue_get(f->inputs[0])) != NULL) {
		mblk_t *it;
		size_t n = msgdsize(m);
		if (s->discard) {
			s->len += n;
			s->nblocks++;
			freemsg(m);
			continue;
		}
		if (s->len + n > s->cap) {
			s->cap = (s->len + n) * 2 + 4096;
			s->buf = (uint8_t *)ms_realloc(s->buf, s->cap);
		}
		for (it = m; it; it = it->b_cont) {
			size_t k = (size_t)(it->b_wptr - it->b_rptr);
			memcpy(s->buf + s->len, it->b_rptr, k);
			s->len += k;
		}
		if (s->nblocks == s->blocks_cap) {
			s->blocks_cap = s->blocks_cap * 2 + 64;
			s->sizes = (int *)ms_realloc(s->sizes, sizeof(int) * 3 * (size_t)s->blocks_cap);
			s->dims = (int *)ms_realloc(s->dims, sizeof(int) * 2 * (size_t)s->blocks_cap);
		}
		s->dims[2 * s->nblocks] = s->dims[2 * s->nblocks + 1] = 0;
		if (m->b_rptr - dblk_base(m->b_datap) >= 16) {
			const uint16_t *hdr = (const uint16_t *)dblk_base(m->b_datap);
			s->dims[2 * s->nblocks] = hdr[0];
			s->dims[2 * s->nblocks + 1] = hdr[1];
		}
		s->sizes[3 * s->nblocks] = s->tick;
		s->sizes[3 * s->nblocks + 1] = (int)n;
		s->sizes[3 * s->nblocks + 2] = (int)mblk_get_timestamp_info(m);
		s->nblocks++;
		freemsg(m);
	}
	s->tick++;
}
static void hsink_uninit(MSFilter *f) {
	HSink *s = (HSink *)f->data;
	ms_free(s->buf);
	ms_free(s->sizes);
	ms_free(s->dims);
	ms_free(s);
}
static MSFilterDesc harness_sink_desc = {.id = MS_FILTER_PLUGIN_ID,
                                         .name = "HarnessSink",
                                         .text = "recording sink",
                                         .category = MS_FILTER_OTHER,
                                         .ninputs = 1,
                                         .noutputs = 0,
                                         .init = hsink_init,
                                         .preprocess = hsink_preprocess,
                                         .process = hsink_process,
                                         .uninit = hsink_uninit};

/* ---------------------------------------------------------------- exported scripting surface */
void *ref_factory_new(const char *plugins_dir) {
	MSFactory *fac = ms_factory_new(); /* registers compat/include/basedescs.h list (reference descs, verbatim) */
	ms_factory_register_filter(fac, &harness_source_desc);
	ms_factory_register_filter(fac, &harness_sink_desc);
	if (plugins_dir && plugins_dir[0]) {
		/* the reference's own loader: scans for libms*.so, dlopen, calls <name>_init(factory)
		 * (src/base/msfactory.c:531-586, 620-759) */
		ms_factory_load_plugins(fac, plugins_dir);
	}
	return fac;
}
void ref_factory_destroy(void *fac) {
	ms_factory_destroy((MSFactory *)fac);
}
void *ref_filter_new(void *fac, const char *name) {
	return ms_factory_create_filter_from_name((MSFactory *)fac, name);
}
const char *ref_filter_text(void *f) {
	return ((MSFilter *)f)->desc->text;
}
void ref_filter_destroy(void *f) {
	ms_filter_destroy((MSFilter *)f);
}
int ref_filter_call(void *f, unsigned int id, void *arg) {
	return ms_filter_call_method((MSFilter *)f, id, arg);
}
int ref_link(void *f1, int pin1, void *f2, int pin2) {
	return ms_filter_link((MSFilter *)f1, pin1, (MSFilter *)f2, pin2);
}
int ref_unlink(void *f1, int pin1, void *f2, int pin2) {
	return ms_filter_unlink((MSFilter *)f1, pin1, (MSFilter *)f2, pin2);
}
void ref_source_push(void *f, int tick, const void *data, int nbytes) {
	HSource *s = (HSource *)((MSFilter *)f)->data;
	HBlock *b = (HBlock *)ms_malloc0(sizeof(HBlock) + (size_t)nbytes);
	b->tick = tick;
	b->nbytes = nbytes;
	memcpy(b->data, data, (size_t)nbytes);
	if (s->tail) s->tail->next = b;
	else s->head = b;
	s->tail = b;
}
/* an I420 frame of w x h (tight planes, w * h * 3 / 2 bytes) or, with w == 0, a raw packed frame; timestamp set */
void ref_source_push_video(void *f, int tick, const void *data, int nbytes, int w, int h, unsigned int ts) {
	HSource *s = (HSource *)((MSFilter *)f)->data;
	ref_source_push(f, tick, data, nbytes);
	s->tail->w = w;
	s->tail->h = h;
	s->tail->has_ts = 1;
	s->tail->ts = ts;
}
void ref_sink_set_discard(void *f, int on) {
	((HSink *)((MSFilter *)f)->data)->discard = on;
}
void ref_sink_read_dims(void *f, int *pairs) {
	HSink *s = (HSink *)((MSFilter *)f)->data;
	memcpy(pairs, s->dims, sizeof(int) * 2 * (size_t)s->nblocks);
}
/* ms_video_set_scaler_impl (src/voip/msvideo.c:719-721) with a desc built from three callbacks (the test's oracle scaler)
 * or with a ready MSScalerDesc (the plugin's msb200_ms_scaler_desc()); NULL restores "no scaler" */
static MSScalerDesc g_cb_scaler;
void ref_set_scaler_callbacks(void *create, void *process, void *ctx_free) {
	g_cb_scaler.create_context = (MSScalerContext * (*)(int, int, MSPixFmt, int, int, MSPixFmt, int)) create;
	g_cb_scaler.context_process = (int (*)(MSScalerContext *, uint8_t *[], int[], uint8_t *[], int[]))process;
	g_cb_scaler.context_free = (void (*)(MSScalerContext *))ctx_free;
	ms_video_set_scaler_impl(&g_cb_scaler);
}
void ref_set_scaler_desc(void *desc) {
	ms_video_set_scaler_impl((MSScalerDesc *)desc);
}
/* push `nblocks` consecutive blocks of `block_bytes`, one per tick starting at tick0 */
void ref_source_push_stream(void *f, int tick0, const void *data, int block_bytes, int nblocks) {
	int i;
	for (i = 0; i < nblocks; ++i)
		ref_source_push(f, tick0 + i, (const uint8_t *)data + (size_t)i * (size_t)block_bytes, block_bytes);
}
long ref_sink_size(void *f) {
	return (long)((HSink *)((MSFilter *)f)->data)->len;
}
int ref_sink_nblocks(void *f) {
	return ((HSink *)((MSFilter *)f)->data)->nblocks;
}
void ref_sink_read(void *f, void *out, int *triples) {
	HSink *s = (HSink *)((MSFilter *)f)->data;
	if (out) memcpy(out, s->buf, s->len);
	if (triples) memcpy(triples, s->sizes, sizeof(int) * 3 * (size_t)s->nblocks);
}

/* ---------------------------------------------------------------- gated ticker */
typedef struct HTicker {
	MSTicker *ticker;
	pthread_mutex_t mu;
	pthread_cond_t cv;
	long allowed; /* number of tick-waits the ticker thread may pass */
	long passed;
	int parked;
} HTicker;

static int gated_tick(void *data, uint64_t virt_time) {
	HTicker *h = (HTicker *)data;
	(void)virt_time;
	pthread_mutex_lock(&h->mu);
	while (h->passed >= h->allowed) {
		h->parked = 1;
		pthread_cond_broadcast(&h->cv);
		pthread_cond_wait(&h->cv, &h->mu);
	}
	h->parked = 0;
	h->passed++;
	pthread_mutex_unlock(&h->mu);
	return 0; /* never late */
}
static void hticker_wait_parked(HTicker *h) {
	pthread_mutex_lock(&h->mu);
	while (!(h->parked && h->passed >= h->allowed))
		pthread_cond_wait(&h->cv, &h->mu);
	pthread_mutex_unlock(&h->mu);
}
void *ref_ticker_new(void) {
	HTicker *h = ms_new0(HTicker, 1);
	pthread_mutex_init(&h->mu, NULL);
	pthread_cond_init(&h->cv, NULL);
	h->ticker = ms_ticker_new();
	ms_ticker_set_tick_func(h->ticker, gated_tick, h);
	hticker_wait_parked(h); /* the thread is now parked inside the wait of an (empty) tick */
	return h;
}
int ref_ticker_attach(void *ht, void *f) {
	return ms_ticker_attach(((HTicker *)ht)->ticker, (MSFilter *)f);
}
int ref_ticker_detach(void *ht, void *f) {
	return ms_ticker_detach(((HTicker *)ht)->ticker, (MSFilter *)f);
}
/* run exactly n ticks of the attached graphs, then park again */
void ref_ticker_run(void *ht, int n) {
	HTicker *h = (HTicker *)ht;
	pthread_mutex_lock(&h->mu);
	h->allowed += n;
	pthread_cond_broadcast(&h->cv);
	pthread_mutex_unlock(&h->mu);
	hticker_wait_parked(h);
}
/* the same in two halves, so that several tickers can run their ticks concurrently: release all, then wait for all */
void ref_ticker_release(void *ht, int n) {
	HTicker *h = (HTicker *)ht;
	pthread_mutex_lock(&h->mu);
	h->allowed += n;
	pthread_cond_broadcast(&h->cv);
	pthread_mutex_unlock(&h->mu);
}
void ref_ticker_wait(void *ht) {
	hticker_wait_parked((HTicker *)ht);
}
unsigned long long ref_ticker_time(void *ht) {
	return (unsigned long long)((HTicker *)ht)->ticker->time;
}
void ref_ticker_destroy(void *ht) {
	HTicker *h = (HTicker *)ht;
	/* let the thread leave the gate so ms_ticker_destroy can join it */
	pthread_mutex_lock(&h->mu);
	h->allowed = 1L << 60;
	pthread_cond_broadcast(&h->cv);
	pthread_mutex_unlock(&h->mu);
	ms_ticker_destroy(h->ticker);
	pthread_mutex_destroy(&h->mu);
	pthread_cond_destroy(&h->cv);
	ms_free(h);
}

/* ---------------------------------------------------------------- method ids by name (the macros embed sizeof(arg)) */
#define MID(x)                                                                                                         \
	if (strcmp(name, #x) == 0) return (unsigned int)(x)
unsigned int ref_method_id(const char *name) {
	MID(MS_FILTER_SET_SAMPLE_RATE);
	MID(MS_FILTER_GET_SAMPLE_RATE);
	MID(MS_FILTER_SET_OUTPUT_SAMPLE_RATE);
	MID(MS_FILTER_SET_NCHANNELS);
	MID(MS_FILTER_GET_NCHANNELS);
	MID(MS_FILTER_SET_OUTPUT_NCHANNELS);
	MID(MS_FILTER_SET_VIDEO_SIZE);
	MID(MS_FILTER_GET_VIDEO_SIZE);
	MID(MS_FILTER_SET_PIX_FMT);
	MID(MS_FILTER_SET_FPS);
	MID(MSB200_PIX_CONV_SET_OUTPUT_FMT);
	MID(MSB200_PIX_CONV_SET_OUTPUT_SIZE);
	MID(MS_AUDIO_MIXER_SET_INPUT_GAIN);
	MID(MS_AUDIO_MIXER_SET_ACTIVE);
	MID(MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE);
	MID(MS_AUDIO_MIXER_SET_MASTER_CHANNEL);
	MID(MS_AUDIO_MIXER_ENABLE_OUTPUT);
	MID(MS_VOLUME_GET);
	MID(MS_VOLUME_GET_LINEAR);
	MID(MS_VOLUME_SET_GAIN);
	MID(MS_VOLUME_GET_GAIN);
	MID(MS_VOLUME_SET_DB_GAIN);
	MID(MS_VOLUME_GET_GAIN_DB);
	MID(MS_VOLUME_ENABLE_NOISE_GATE);
	MID(MS_VOLUME_SET_NOISE_GATE_THRESHOLD);
	MID(MS_VOLUME_SET_NOISE_GATE_FLOORGAIN);
	MID(MS_VOLUME_REMOVE_DC);
	MID(MS_VOLUME_ENABLE_AGC);
	MID(MS_VOLUME_SET_PEER);
	MID(MS_VOLUME_SET_EA_THRESHOLD);
	MID(MS_VOLUME_SET_EA_SPEED);
	MID(MS_VOLUME_SET_EA_FORCE);
	MID(MS_VOLUME_SET_EA_SUSTAIN);
	MID(MS_VOLUME_SET_EA_TRANSMIT_THRESHOLD);
	MID(MS_VOLUME_GET_MIN);
	MID(MS_VOLUME_GET_MAX);
	MID(MS_EQUALIZER_SET_GAIN);
	MID(MS_EQUALIZER_GET_GAIN);
	MID(MS_EQUALIZER_SET_ACTIVE);
	MID(MS_EQUALIZER_DUMP_STATE);
	MID(MS_EQUALIZER_GET_NUM_FREQUENCIES);
	MID(MS_CHANNEL_ADAPTER_SET_OUTPUT_NCHANNELS);
	MID(MS_CHANNEL_ADAPTER_GET_OUTPUT_NCHANNELS);
	MID(MS_ECHO_CANCELLER_SET_TAIL_LENGTH);
	MID(MS_ECHO_CANCELLER_SET_DELAY);
	MID(MS_ECHO_CANCELLER_SET_FRAMESIZE);
	MID(MS_ECHO_CANCELLER_SET_BYPASS_MODE);
	MID(MS_ECHO_CANCELLER_GET_BYPASS_MODE);
	MID(MS_ECHO_CANCELLER_GET_STATE_STRING);
	MID(MS_ECHO_CANCELLER_SET_STATE_STRING);
	MID(MS_FILTER_ADD_FMTP);
	MID(MS_FILTER_ADD_ATTR);
	MID(MS_AUDIO_ENCODER_GET_PTIME);
	MID(MS_DECODER_HAVE_PLC);
	MID(MS_AUDIO_FLOW_CONTROL_SET_CONFIG);
	MID(MS_AUDIO_FLOW_CONTROL_DROP);
	MID(MS_GENERIC_PLC_SET_CN);
	return 0;
}

/* ---------------------------------------------------------------- direct call: NV12 -> I420 (+rotate, +1/2 downscale) */
/* src/voip/msvideo.c:787-919; returns bytes written to out (w*h*3/2 of the DESTINATION geometry) or -1 */
int ref_nv12_to_i420(const uint8_t *y, const uint8_t *cbcr, int rotation, int w, int h, int y_stride, int cbcr_stride,
                     int u_first, int down_scale, uint8_t *out) {
	MSYuvBufAllocator *alloc = ms_yuv_buf_allocator_new();
	mblk_t *m = copy_ycbcrbiplanar_to_true_yuv_with_rotation_and_down_scale_by_2(
	    alloc, y, cbcr, rotation, w, h, y_stride, cbcr_stride, (bool_t)u_first, (bool_t)down_scale);
	int n = -1;
	if (m) {
		n = w * h * 3 / 2;
		memcpy(out, m->b_rptr, (size_t)n);
		freemsg(m);
	}
	ms_yuv_buf_allocator_free(alloc);
	return n;
}

/* ---------------------------------------------------------------- direct call: ms_fir_mem16 (float build) */
#include "mediastreamer2/dsptools.h"
/* src/utils/dsptools.c:253-268 */
void ref_fir_mem16(const float *x, const float *num, float *y, int N, int ord, float *mem) {
	ms_fir_mem16((const ms_word16_t *)x, (const ms_coef_t *)num, (ms_word16_t *)y, N, ord, (ms_mem_t *)mem);
}
