/*
 * plugin/msb200_video_filters.c — the video half of libmsb200filters.so: MSPixConv, MSSizeConv and the MSScalerDesc.
 *
 * Reference behaviour replaced (host logic here, every pixel in libmsb200dsp.so's kernels):
 *   MSPixConv   /root/reference/src/videofilters/pixconv.c:62-94 (process), :96-110 (methods), ids/names :114-139
 *   MSSizeConv  /root/reference/src/videofilters/sizeconv.c:97-184 (process: frame pacing :104-132, geometry :141-157),
 *               :186-217 (methods)
 *   MSScaler    /root/reference/include/mediastreamer2/msvideo.h:473-492, installed with ms_video_set_scaler_impl
 *   frame mblk  /root/reference/src/voip/msvideo.c:79-83, 283-304: a 16-byte {uint16 w, h} header BELOW b_rptr, found by
 *               consumers through dblk_base() (ms_yuv_buf_init_from_mblk :100-111)
 *
 * Design (not the reference's): frames do not go through a per-filter scaler call. Filters with the same conversion
 * geometry share a LANE: a pinned source arena, a ring of pinned destination slots and one batched msb200_scaler.
 *   stage    process() copies the input frame's payload into the next arena position (the one host copy: mblk memory
 *            is pageable) and remembers its timestamp
 *   flush    one H2D of the staged run, ONE kernel sequence over all staged frames, D2H straight into free ring slots
 *   emit     each result is handed downstream as an mblk that POINTS INTO its pinned slot (esballoc): no copy out; the
 *            slot returns to the ring when the last reference to the mblk is freed
 * Synchronous mode (default): a lane per filter, flushed at once — same tick, same order as the reference filter.
 * Lockstep batch mode (MSB200_BATCH > 0): one lane per (MSTicker, geometry); the first member called in a tick flushes
 * what all members staged during the previous tick, so a frame leaves one ticker interval later, as in the audio groups.
 */
#include "msb200_plugin.h"

#include "mediastreamer2/allfilters.h"
#include "msb200_ms2.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ================================================================================================ formats */
int msb200p_pixfmt_to_b200(MSPixFmt fmt) {
	if (fmt == MSB200_MS_NV12) return MSB200_PIX_NV12;
	if (fmt == MSB200_MS_NV21) return MSB200_PIX_NV21;
	switch (fmt) {
		case MS_YUV420P: return MSB200_PIX_YUV420P;
		case MS_YUYV: return MSB200_PIX_YUYV;
		case MS_YUY2: return MSB200_PIX_YUY2;
		case MS_UYVY: return MSB200_PIX_UYVY;
		case MS_RGB24: return MSB200_PIX_RGB24;
		case MS_RGB24_REV: return MSB200_PIX_RGB24_REV;
		case MS_RGBA32: return MSB200_PIX_RGBA32;
		case MS_RGBA32_REV: return MSB200_PIX_RGBA32_REV;
		case MS_RGB565: return MSB200_PIX_RGB565;
		default: return -1;
	}
}
/* bytes of one tight frame and of one of its rows when it is a single packed plane (0: planar) */
/* MSB200_SWS_X86=1: planar outputs rounded like libswscale's x86 SIMD vertical scaler — what the reference's plain
 * SWS_BILINEAR call (src/voip/msvideo.c:660) returns on an x86 host; default: the library's C arithmetic (differs by <= 1).
 * Formats without that stage (RGB outputs, YUY2 / UYVY / BGR24 inputs) refuse the switch and stay as they are. */
static void sws_rounding(msb200_scaler *sc) {
	const char *e = getenv("MSB200_SWS_X86");
	if (e && atoi(e) > 0) (void)msb200_scaler_set_x86_vertical(sc, 1);
}

static size_t frame_bytes(int b200_fmt, int w, int h, int *packed_row) {
	const int he = h & 1 ? h + 1 : h; /* the reference rounds odd heights up when it sizes a frame (msvideo.c:158) */
	int row = 0;
	size_t n;
	switch (b200_fmt) {
		case MSB200_PIX_YUYV: case MSB200_PIX_YUY2: case MSB200_PIX_UYVY: case MSB200_PIX_RGB565: row = w * 2; break;
		case MSB200_PIX_RGB24: case MSB200_PIX_RGB24_REV: row = w * 3; break;
		case MSB200_PIX_RGBA32: case MSB200_PIX_RGBA32_REV: row = w * 4; break;
		default: break;
	}
	n = row ? (size_t)row * h : (size_t)w * he * 3 / 2;
	if (packed_row) *packed_row = row;
	return n;
}

/* ================================================================================================ lanes */
#define VSLOT_PREFIX 16 /* our bookkeeping, below the reference's 16-byte video header */
#define VHDR 16         /* sizeof(mblk_video_header), msvideo.c:79-83 */
typedef struct VSlotPrefix {
	struct VLane *lane;
	int idx, pad;
} VSlotPrefix;

struct VMember;
typedef struct VLane {
	struct VLane *next;
	MSTicker *ticker; /* NULL: the private lane of one synchronous filter */
	int key[6];       /* src w, h, fmt (b200), dst w, h, fmt (b200) */
	int refs, lent, dead;
	msb200_ctx *ctx;
	int owns_ctx;
	msb200_scaler *sc;
	size_t src_bytes, dst_bytes, dst_stride;
	int cap, n_staged, n_slots, cursor;
	uint8_t *src, *dst, *busy;
	struct VMember **who;
	uint32_t *ts;
	const uint8_t **srcv;
	uint8_t **dstv;
	int *slotv;
	uint64_t seen_tick, flushes, frames;
} VLane;
typedef struct VMember {
	VLane *lane;
	queue_t ready; /* results waiting to be put on the filter's output */
} VMember;

static VLane *g_lanes = NULL;
static pthread_mutex_t g_vmu = PTHREAD_MUTEX_INITIALIZER; /* lane list, refs, slot ring bookkeeping */
static uint64_t g_vflushes = 0, g_vframes = 0;

static int video_batch_capacity(void) {
	const char *e = getenv("MSB200_VIDEO_BATCH");
	int cap = msb200p_batch_capacity(), v = e ? atoi(e) : 64;
	if (cap <= 0) return 0;
	if (v < 1) v = 1;
	return cap < v ? cap : v;
}
static void lane_destroy(VLane *l) { /* unlinked, no refs, nothing lent */
	if (l->owns_ctx) msb200_ctx_make_current(l->ctx);
	else msb200p_sync_lock();
	if (l->sc) msb200_scaler_destroy(l->sc);
	if (l->src) msb200_host_free_pinned(l->ctx, l->src);
	if (l->dst) msb200_host_free_pinned(l->ctx, l->dst);
	if (l->owns_ctx) msb200_ctx_destroy(l->ctx);
	else msb200p_sync_unlock();
	ms_free(l->busy);
	ms_free(l->who);
	ms_free(l->ts);
	ms_free((void *)l->srcv);
	ms_free(l->dstv);
	ms_free(l->slotv);
	ms_free(l);
}
/* esballoc's free callback: the last reference to an output frame is gone, its slot returns to the ring */
static void vslot_release(void *buf) {
	VSlotPrefix *p = (VSlotPrefix *)((uint8_t *)buf - VSLOT_PREFIX);
	VLane *l = p->lane;
	int destroy;
	pthread_mutex_lock(&g_vmu);
	l->busy[p->idx] = 0;
	l->lent--;
	destroy = l->dead && l->lent == 0;
	pthread_mutex_unlock(&g_vmu);
	if (destroy) lane_destroy(l);
}
static VLane *lane_join(MSTicker *ticker, const int key[6]) {
	VLane *l;
	int rc, i;
	const int batch = ticker ? video_batch_capacity() : 0;
	pthread_mutex_lock(&g_vmu);
	if (batch > 0) {
		for (l = g_lanes; l; l = l->next)
			if (l->ticker == ticker && memcmp(l->key, key, sizeof(l->key)) == 0) {
				l->refs++;
				pthread_mutex_unlock(&g_vmu);
				return l;
			}
	}
	l = ms_new0(VLane, 1);
	memcpy(l->key, key, sizeof(l->key));
	l->refs = 1;
	l->seen_tick = (uint64_t)-1;
	if (batch > 0) {
		l->ticker = ticker;
		l->cap = batch;
		l->owns_ctx = 1;
		rc = msb200_ctx_create(msb200p_device_of_ticker(ticker), &l->ctx);
	} else {
		l->cap = 1;
		l->ctx = msb200p_sync_ctx();
		rc = l->ctx ? MSB200_OK : MSB200_ENODEV;
		if (rc == MSB200_OK) msb200p_sync_lock();
	}
	if (rc == MSB200_OK) rc = msb200_scaler_create(l->ctx, key[0], key[1], key[2], key[3], key[4], key[5], &l->sc);
	if (rc == MSB200_OK) {
		sws_rounding(l->sc);
		l->src_bytes = msb200_scaler_src_frame_bytes(l->sc);
		l->dst_bytes = msb200_scaler_dst_frame_bytes(l->sc);
		l->dst_stride = (VSLOT_PREFIX + VHDR + l->dst_bytes + 16 + 255) & ~(size_t)255;
		l->n_slots = 3 * l->cap + 2; /* results of the last flush + frames still held downstream */
		rc = msb200_host_alloc_pinned(l->ctx, (size_t)l->cap * l->src_bytes, (void **)&l->src);
		if (rc == MSB200_OK) rc = msb200_host_alloc_pinned(l->ctx, (size_t)l->n_slots * l->dst_stride, (void **)&l->dst);
	}
	if (!l->owns_ctx && l->ctx) msb200p_sync_unlock();
	if (rc != MSB200_OK) {
		ms_error("msb200 video: cannot set up %dx%d fmt %d -> %dx%d fmt %d (%s)", key[0], key[1], key[2], key[3], key[4], key[5],
		         msb200_last_error());
		pthread_mutex_unlock(&g_vmu);
		if (l->ctx) lane_destroy(l);
		else ms_free(l);
		return NULL;
	}
	l->busy = (uint8_t *)ms_malloc0((size_t)l->n_slots);
	l->who = (VMember **)ms_malloc0(sizeof(VMember *) * (size_t)l->cap);
	l->ts = (uint32_t *)ms_malloc0(sizeof(uint32_t) * (size_t)l->cap);
	l->srcv = (const uint8_t **)ms_malloc0(sizeof(uint8_t *) * (size_t)l->cap);
	l->dstv = (uint8_t **)ms_malloc0(sizeof(uint8_t *) * (size_t)l->cap);
	l->slotv = (int *)ms_malloc0(sizeof(int) * (size_t)l->cap);
	for (i = 0; i < l->n_slots; ++i) {
		VSlotPrefix *p = (VSlotPrefix *)(l->dst + (size_t)i * l->dst_stride);
		p->lane = l;
		p->idx = i;
	}
	if (batch > 0) {
		l->next = g_lanes;
		g_lanes = l;
		ms_message("msb200 video: lane %p on ticker %p: %dx%d fmt %d -> %dx%d fmt %d, %d frames per flush", l, ticker, key[0], key[1],
		           key[2], key[3], key[4], key[5], l->cap);
	}
	pthread_mutex_unlock(&g_vmu);
	return l;
}
static void lane_leave(VLane *l, VMember *m) {
	VLane **pp;
	int k, destroy;
	if (!l) return;
	pthread_mutex_lock(&g_vmu);
	for (k = 0; k < l->n_staged; ++k)
		if (l->who[k] == m) l->who[k] = NULL; /* still converted with the batch, delivered to nobody */
	destroy = 0;
	if (--l->refs == 0) {
		for (pp = &g_lanes; *pp && *pp != l; pp = &(*pp)->next) {
		}
		if (*pp) *pp = l->next;
		l->dead = 1;
		l->n_staged = 0;
		destroy = l->lent == 0;
	}
	pthread_mutex_unlock(&g_vmu);
	if (destroy) lane_destroy(l);
}
/* run everything staged: one upload, one kernel sequence, downloads into ring slots, results queued on their members */
static void lane_flush(VLane *l) {
	int k, n, rc;
	mblk_t *drop;
	if (l->n_staged == 0) return;
	n = l->n_staged;
	pthread_mutex_lock(&g_vmu);
	for (k = 0; k < n; ++k) {
		int tries = 0, s = l->cursor;
		while (tries < l->n_slots && l->busy[s]) {
			s = s + 1 == l->n_slots ? 0 : s + 1;
			++tries;
		}
		if (tries == l->n_slots) break; /* downstream holds every slot: the rest of the batch is dropped */
		l->busy[s] = 1;
		l->lent++;
		l->slotv[k] = s;
		l->cursor = s + 1 == l->n_slots ? 0 : s + 1;
		l->srcv[k] = l->src + (size_t)k * l->src_bytes;
		l->dstv[k] = l->dst + (size_t)s * l->dst_stride + VSLOT_PREFIX + VHDR;
	}
	pthread_mutex_unlock(&g_vmu);
	if (k < n) {
		ms_warning("msb200 video: lane %p has no free output slot, %d frame(s) dropped", l, n - k);
		n = k;
	}
	rc = MSB200_OK;
	if (n > 0) {
		if (l->owns_ctx) msb200_ctx_make_current(l->ctx);
		else msb200p_sync_lock();
		rc = msb200_scaler_process_frames(l->sc, n, l->srcv, l->dstv);
		if (!l->owns_ctx) msb200p_sync_unlock();
		if (rc != MSB200_OK) ms_error("msb200 video: lane %p: %s", l, msb200_last_error());
	}
	/* delivery under the lane lock: a member that is leaving on another thread (postprocess) either still gets its frame
	 * here, before lane_leave() clears its entries and its queue is flushed, or is no longer seen at all; frames nobody
	 * takes are freed after the lock is dropped (freeing one takes the lock again to return its slot) */
	drop = NULL;
	pthread_mutex_lock(&g_vmu);
	for (k = 0; k < n; ++k) {
		uint8_t *base = l->dst + (size_t)l->slotv[k] * l->dst_stride + VSLOT_PREFIX;
		mblk_t *m = esballoc(base, VHDR + l->dst_bytes + 16, 0, vslot_release);
		uint16_t *hdr = (uint16_t *)base;
		memset(base, 0, VHDR);
		hdr[0] = (uint16_t)l->key[3];
		hdr[1] = (uint16_t)l->key[4];
		m->b_rptr = base + VHDR;
		m->b_wptr = m->b_rptr + l->dst_bytes;
		mblk_set_timestamp_info(m, l->ts[k]);
		if (rc == MSB200_OK && l->who[k]) putq(&l->who[k]->ready, m);
		else {
			m->b_next = drop;
			drop = m;
		}
	}
	pthread_mutex_unlock(&g_vmu);
	while (drop) { /* gives the slots back */
		mblk_t *m = drop;
		drop = m->b_next;
		m->b_next = NULL;
		freemsg(m);
	}
	l->flushes++;
	l->frames += (uint64_t)n;
	pthread_mutex_lock(&g_vmu);
	g_vflushes++;
	g_vframes += (uint64_t)n;
	pthread_mutex_unlock(&g_vmu);
	l->n_staged = 0;
}
/* batch mode, first thing in a member's process(): the first caller of a tick flushes the previous tick's frames */
static void lane_tick(VLane *l, uint64_t ticks) {
	if (!l || !l->ticker || l->seen_tick == ticks) return;
	l->seen_tick = ticks;
	lane_flush(l);
}
/* the next arena position (flushing early when a tick stages more frames than the arena holds) */
static uint8_t *lane_stage(VLane *l, VMember *m, uint32_t ts) {
	uint8_t *p;
	if (l->n_staged == l->cap) lane_flush(l);
	p = l->src + (size_t)l->n_staged * l->src_bytes;
	l->who[l->n_staged] = m;
	l->ts[l->n_staged] = ts;
	l->n_staged++;
	return p;
}
static void member_emit(VMember *m, MSQueue *out) {
	mblk_t *r;
	while ((r = getq(&m->ready)) != NULL) {
		if (out) ms_queue_put(out, r);
		else freemsg(r);
	}
}
static void copy_rows(uint8_t *dst, const uint8_t *src, ptrdiff_t src_stride, size_t row, int rows) {
	int y;
	if (src_stride == (ptrdiff_t)row) {
		memcpy(dst, src, row * (size_t)rows);
		return;
	}
	for (y = 0; y < rows; ++y) memcpy(dst + (size_t)y * row, src + (ptrdiff_t)y * src_stride, row);
}

/* ================================================================================================ MSPixConv */
typedef struct PixConv {
	VMember m;
	MSVideoSize in_size, out_size; /* out_size {0, 0}: same as the input */
	MSPixFmt in_fmt, out_fmt;
	int key[6];
} PixConv;

static void pixc_new(MSFilter *f) {
	PixConv *s = ms_new0(PixConv, 1);
	qinit(&s->m.ready);
	s->in_size.width = MS_VIDEO_SIZE_CIF_W; /* the reference's defaults, pixconv.c:37-46 */
	s->in_size.height = MS_VIDEO_SIZE_CIF_H;
	s->in_fmt = s->out_fmt = MS_YUV420P;
	f->data = s;
}
static void pixconv_drop_lane(PixConv *s) {
	lane_leave(s->m.lane, &s->m);
	s->m.lane = NULL;
}
static void pixc_free(MSFilter *f) {
	PixConv *s = (PixConv *)f->data;
	pixconv_drop_lane(s);
	flushq(&s->m.ready, 0);
	ms_free(s);
}
static void pixc_detach(MSFilter *f) {
	PixConv *s = (PixConv *)f->data;
	pixconv_drop_lane(s);
	flushq(&s->m.ready, 0);
}
static VLane *pixconv_lane(MSFilter *f, PixConv *s) {
	int key[6];
	key[0] = s->in_size.width;
	key[1] = s->in_size.height;
	key[2] = msb200p_pixfmt_to_b200(s->in_fmt);
	key[3] = s->out_size.width > 0 ? s->out_size.width : s->in_size.width;
	key[4] = s->out_size.height > 0 ? s->out_size.height : s->in_size.height;
	key[5] = msb200p_pixfmt_to_b200(s->out_fmt);
	if (s->m.lane && memcmp(key, s->key, sizeof(key)) == 0) return s->m.lane;
	pixconv_drop_lane(s);
	memcpy(s->key, key, sizeof(key));
	if (key[2] < 0 || key[5] < 0) {
		ms_error("MSPixConv(B200): unsupported conversion %s -> %s", ms_pix_fmt_to_string(s->in_fmt), ms_pix_fmt_to_string(s->out_fmt));
		return NULL;
	}
	s->m.lane = lane_join(f->ticker, key);
	return s->m.lane;
}
static void pixc_tick(MSFilter *f) {
	PixConv *s = (PixConv *)f->data;
	mblk_t *im;
	const int resize = s->out_size.width > 0 && (s->out_size.width != s->in_size.width || s->out_size.height != s->in_size.height);
	lane_tick(s->m.lane, f->ticker->ticks);
	member_emit(&s->m, f->outputs[0]);
	while ((im = ms_queue_get(f->inputs[0])) != NULL) {
		VLane *l;
		mblk_t *body;
		int row = 0;
		size_t need;
		if (s->in_fmt == s->out_fmt && !resize) { /* nothing to convert: the frame itself goes on (pixconv.c:68-69) */
			ms_queue_put(f->outputs[0], im);
			continue;
		}
		l = pixconv_lane(f, s);
		body = im->b_cont ? im->b_cont : im; /* a leading block may only carry the video header (msvideo.c:120-121) */
		need = frame_bytes(s->key[2], s->key[0], s->key[1], &row);
		if (l && (size_t)(body->b_wptr - body->b_rptr) >= need && need == l->src_bytes) {
			uint8_t *dst = lane_stage(l, &s->m, mblk_get_timestamp_info(im));
			if (s->in_fmt == MS_RGB24_REV) /* bottom-up DIB: last stored row first (pixconv.c:78-81) */
				copy_rows(dst, body->b_rptr + (size_t)row * (s->key[1] - 1), -(ptrdiff_t)row, (size_t)row, s->key[1]);
			else memcpy(dst, body->b_rptr, need);
			if (!l->ticker) { /* synchronous lane: convert now, emit in this tick */
				lane_flush(l);
				member_emit(&s->m, f->outputs[0]);
			}
		} else if (l) {
			ms_error("MSPixConv(B200): a %dx%d %s frame needs %zu bytes, the block holds %zu", s->key[0], s->key[1],
			         ms_pix_fmt_to_string(s->in_fmt), need, (size_t)(body->b_wptr - body->b_rptr));
		}
		freemsg(im);
	}
}
static int pixc_take_size(MSFilter *f, void *arg) {
	((PixConv *)f->data)->in_size = *(MSVideoSize *)arg;
	return 0;
}
static int pixc_take_format(MSFilter *f, void *arg) {
	((PixConv *)f->data)->in_fmt = *(MSPixFmt *)arg;
	return 0;
}
static int pixconv_set_out_fmt(MSFilter *f, void *arg) {
	const MSPixFmt fmt = *(MSPixFmt *)arg;
	if (fmt != MS_YUV420P && fmt != MS_RGB24 && fmt != MS_RGB24_REV) return -1;
	((PixConv *)f->data)->out_fmt = fmt;
	return 0;
}
static int pixconv_set_out_size(MSFilter *f, void *arg) {
	((PixConv *)f->data)->out_size = *(MSVideoSize *)arg;
	return 0;
}
static MSFilterMethod pixc_method_table[] = {{MS_FILTER_SET_VIDEO_SIZE, pixc_take_size},
                                           {MS_FILTER_SET_PIX_FMT, pixc_take_format},
                                           {MSB200_PIX_CONV_SET_OUTPUT_FMT, pixconv_set_out_fmt},
                                           {MSB200_PIX_CONV_SET_OUTPUT_SIZE, pixconv_set_out_size},
                                           {0, NULL}};
static MSFilterDesc b200_pix_conv_desc = {.id = MS_PIX_CONV_ID,
                                          .name = "MSPixConv",
                                          .text = "B200: pixel format converter (libmsb200dsp)",
                                          .category = MS_FILTER_OTHER,
                                          .ninputs = 1,
                                          .noutputs = 1,
                                          .init = pixc_new,
                                          .process = pixc_tick,
                                          .postprocess = pixc_detach,
                                          .uninit = pixc_free,
                                          .methods = pixc_method_table};

/* ================================================================================================ MSSizeConv */
typedef struct SizeConv {
	VMember m;
	MSVideoSize target;
	float fps;      /* < 0: every frame */
	float t0;       /* ticker time of the first tick after (re)start */
	int running;    /* t0 / emitted valid */
	int emitted;    /* frames let through since t0 */
	int waiting_for_host; /* the target was corrected (orientation / aspect) and the host told: frames of another size wait */
	queue_t pending;
	int key[6];
} SizeConv;

static void sizec_new(MSFilter *f) {
	SizeConv *s = ms_new0(SizeConv, 1);
	qinit(&s->m.ready);
	qinit(&s->pending);
	s->target.width = MS_VIDEO_SIZE_CIF_W; /* sizeconv.c:44-58 */
	s->target.height = MS_VIDEO_SIZE_CIF_H;
	s->fps = -1;
	f->data = s;
}
static void sizeconv_drop_lane(SizeConv *s) {
	lane_leave(s->m.lane, &s->m);
	s->m.lane = NULL;
}
static void sizec_detach(MSFilter *f) {
	SizeConv *s = (SizeConv *)f->data;
	sizeconv_drop_lane(s);
	flushq(&s->pending, 0);
	flushq(&s->m.ready, 0);
	s->running = 0;
}
static void sizec_free(MSFilter *f) {
	sizec_detach(f);
	ms_free(f->data);
}
/* The size a frame of in_w x in_h is scaled to: the configured target turned to the frame's orientation, then shrunk along
 * one axis so that the frame's aspect ratio survives (sizeconv.c:141-157, same integer arithmetic). */
static MSVideoSize sizeconv_fit(MSVideoSize target, int in_w, int in_h) {
	MSVideoSize t = target;
	if ((in_w >= in_h) != (t.width >= t.height)) {
		t.width = target.height;
		t.height = target.width;
	}
	if (in_w * t.height / t.width != in_h) {
		if (in_w > in_h) t.height = in_h * t.width / in_w;
		else t.width = in_w * t.height / in_h;
	}
	return t;
}
static VLane *sizeconv_lane(MSFilter *f, SizeConv *s, int in_w, int in_h) {
	int key[6];
	key[0] = in_w; key[1] = in_h; key[2] = MSB200_PIX_YUV420P;
	key[3] = s->target.width; key[4] = s->target.height; key[5] = MSB200_PIX_YUV420P;
	if (s->m.lane && memcmp(key, s->key, sizeof(key)) == 0) return s->m.lane;
	sizeconv_drop_lane(s);
	memcpy(s->key, key, sizeof(key));
	s->m.lane = lane_join(f->ticker, key);
	return s->m.lane;
}
static void sizeconv_one_frame(MSFilter *f, SizeConv *s, mblk_t *im) {
	YuvBuf in;
	MSVideoSize fit;
	VLane *l;
	ms_yuv_buf_init_from_mblk(&in, im); /* w, h from the header below b_rptr; tight planes */
	s->emitted++;
	if (in.w == s->target.width && in.h == s->target.height) {
		ms_queue_put(f->outputs[0], im);
		return;
	}
	fit = sizeconv_fit(s->target, in.w, in.h);
	if (fit.width != s->target.width || fit.height != s->target.height) {
		s->target = fit;
		s->waiting_for_host = 1;
		ms_filter_notify_no_arg(f, MS_FILTER_OUTPUT_FMT_CHANGED);
	} else if (s->waiting_for_host) {
		ms_warning("MSSizeConv(B200): output format changed, waiting");
	} else if ((l = sizeconv_lane(f, s, in.w, in.h)) != NULL) {
		const int cw = in.w / 2, ch = (in.h & 1 ? in.h + 1 : in.h) / 2;
		uint8_t *dst = lane_stage(l, &s->m, mblk_get_timestamp_info(im));
		copy_rows(dst, in.planes[0], in.strides[0], (size_t)in.w, in.h);
		copy_rows(dst + (size_t)in.w * in.h, in.planes[1], in.strides[1], (size_t)cw, ch);
		copy_rows(dst + (size_t)in.w * in.h + (size_t)cw * ch, in.planes[2], in.strides[2], (size_t)cw, ch);
		if (!l->ticker) {
			lane_flush(l);
			member_emit(&s->m, f->outputs[0]);
		}
	}
	freemsg(im);
}
static void sizec_tick(MSFilter *f) {
	SizeConv *s = (SizeConv *)f->data;
	mblk_t *im;
	ms_filter_lock(f);
	lane_tick(s->m.lane, f->ticker->ticks);
	member_emit(&s->m, f->outputs[0]);
	if (!s->running) {
		s->t0 = (float)f->ticker->time;
		s->emitted = 0;
		s->running = 1;
	}
	while ((im = ms_queue_get(f->inputs[0])) != NULL) putq(&s->pending, im);
	if (s->fps >= 0) {
		/* paced output: frame number `due` is owed at this instant; only the newest captured frame is ever a candidate */
		const int due = (int)((f->ticker->time - s->t0) * s->fps / 1000.0);
		while (s->pending.q_mcount > 1) freemsg(getq(&s->pending));
		if (due <= s->emitted) {
			ms_filter_unlock(f);
			return;
		}
	}
	while ((im = getq(&s->pending)) != NULL) sizeconv_one_frame(f, s, im);
	ms_filter_unlock(f);
}
static int sizec_take_size(MSFilter *f, void *arg) {
	SizeConv *s = (SizeConv *)f->data;
	ms_filter_lock(f);
	s->target = *(MSVideoSize *)arg;
	ms_filter_unlock(f);
	return 0;
}
static int sizec_tell_size(MSFilter *f, void *arg) {
	*(MSVideoSize *)arg = ((SizeConv *)f->data)->target;
	return 0;
}
static int sizec_take_fps(MSFilter *f, void *arg) {
	SizeConv *s = (SizeConv *)f->data;
	s->fps = *(float *)arg;
	s->running = 0; /* pacing restarts from the next tick */
	return 0;
}
static MSFilterMethod sizec_method_table[] = {{MS_FILTER_SET_FPS, sizec_take_fps},
                                            {MS_FILTER_SET_VIDEO_SIZE, sizec_take_size},
                                            {MS_FILTER_GET_VIDEO_SIZE, sizec_tell_size},
                                            {0, NULL}};
static MSFilterDesc b200_size_conv_desc = {.id = MS_SIZE_CONV_ID,
                                           .name = "MSSizeConv",
                                           .text = "B200: video size converter (libmsb200dsp)",
                                           .category = MS_FILTER_OTHER,
                                           .ninputs = 1,
                                           .noutputs = 1,
                                           .init = sizec_new,
                                           .process = sizec_tick,
                                           .postprocess = sizec_detach,
                                           .uninit = sizec_free,
                                           .methods = sizec_method_table};

/* ================================================================================================ MSScalerDesc
 * The second drop-in boundary: the reference's OWN MSPixConv / MSSizeConv / display filters scale on the GPU once this
 * desc is installed (ms_video_set_scaler_impl). One frame per synchronous call with caller-owned planes of any stride:
 * planes that already form one tight frame go to the device as they are, anything else is packed through pinned memory. */
typedef struct B200ScalerCtx {
	msb200_scaler *sc;
	int src_w, src_h, dst_w, dst_h, src_fmt, dst_fmt;
	size_t src_bytes, dst_bytes;
	uint8_t *src_pack, *dst_pack; /* pinned */
} B200ScalerCtx;

static MSScalerContext *b200_scaler_create(int src_w, int src_h, MSPixFmt src_fmt, int dst_w, int dst_h, MSPixFmt dst_fmt, int flags) {
	B200ScalerCtx *c;
	msb200_ctx *ctx;
	const int sf = msb200p_pixfmt_to_b200(src_fmt), df = msb200p_pixfmt_to_b200(dst_fmt);
	int rc;
	(void)flags; /* MS_SCALER_METHOD_*: the kernels implement the bilinear method, which is what both filters ask for */
	if (sf < 0 || (df != MSB200_PIX_YUV420P && df != MSB200_PIX_RGB24 && df != MSB200_PIX_RGB24_REV)) {
		ms_error("msb200 scaler: unsupported conversion %s -> %s", ms_pix_fmt_to_string(src_fmt), ms_pix_fmt_to_string(dst_fmt));
		return NULL;
	}
	if ((ctx = msb200p_sync_ctx()) == NULL) return NULL;
	c = ms_new0(B200ScalerCtx, 1);
	c->src_w = src_w; c->src_h = src_h; c->dst_w = dst_w; c->dst_h = dst_h; c->src_fmt = sf; c->dst_fmt = df;
	msb200p_sync_lock();
	rc = msb200_scaler_create(ctx, src_w, src_h, sf, dst_w, dst_h, df, &c->sc);
	if (rc == MSB200_OK) {
		sws_rounding(c->sc);
		c->src_bytes = msb200_scaler_src_frame_bytes(c->sc);
		c->dst_bytes = msb200_scaler_dst_frame_bytes(c->sc);
		rc = msb200_host_alloc_pinned(ctx, c->src_bytes, (void **)&c->src_pack);
		if (rc == MSB200_OK) rc = msb200_host_alloc_pinned(ctx, c->dst_bytes, (void **)&c->dst_pack);
	}
	msb200p_sync_unlock();
	if (rc != MSB200_OK) {
		ms_error("msb200 scaler: %s", msb200_last_error());
		if (c->sc) {
			msb200p_sync_lock();
			msb200_scaler_destroy(c->sc);
			if (c->src_pack) msb200_host_free_pinned(ctx, c->src_pack);
			msb200p_sync_unlock();
		}
		ms_free(c);
		return NULL;
	}
	return (MSScalerContext *)c;
}
static int b200_scaler_process(MSScalerContext *ctx, uint8_t *src[], int src_strides[], uint8_t *dst[], int dst_strides[]) {
	B200ScalerCtx *c = (B200ScalerCtx *)ctx;
	int rc, row = 0;
	uint8_t *p = c->src_pack;
	const int cw = (c->src_w + 1) / 2, ch = (c->src_h + 1) / 2;
	frame_bytes(c->src_fmt, c->src_w, c->src_h, &row);
	if (row) { /* one packed plane; MSPixConv hands MS_RGB24_REV over with a negative stride (pixconv.c:78-81) */
		copy_rows(p, src[0], src_strides[0], (size_t)row, c->src_h);
	} else if (c->src_fmt == MSB200_PIX_NV12 || c->src_fmt == MSB200_PIX_NV21) {
		copy_rows(p, src[0], src_strides[0], (size_t)c->src_w, c->src_h);
		copy_rows(p + (size_t)c->src_w * c->src_h, src[1], src_strides[1], (size_t)cw * 2, ch);
	} else {
		copy_rows(p, src[0], src_strides[0], (size_t)c->src_w, c->src_h);
		p += (size_t)c->src_w * c->src_h;
		copy_rows(p, src[1], src_strides[1], (size_t)cw, ch);
		copy_rows(p + (size_t)cw * ch, src[2], src_strides[2], (size_t)cw, ch);
	}
	msb200p_sync_lock();
	rc = msb200_scaler_process(c->sc, 1, c->src_pack, c->dst_pack);
	msb200p_sync_unlock();
	if (rc != MSB200_OK) {
		ms_error("msb200 scaler: %s", msb200_last_error());
		return -1;
	}
	p = c->dst_pack;
	if (c->dst_fmt == MSB200_PIX_YUV420P) {
		const int dcw = (c->dst_w + 1) / 2, dch = (c->dst_h + 1) / 2;
		int y;
		for (y = 0; y < c->dst_h; ++y) memcpy(dst[0] + (ptrdiff_t)y * dst_strides[0], p + (size_t)y * c->dst_w, (size_t)c->dst_w);
		p += (size_t)c->dst_w * c->dst_h;
		for (y = 0; y < dch; ++y) memcpy(dst[1] + (ptrdiff_t)y * dst_strides[1], p + (size_t)y * dcw, (size_t)dcw);
		p += (size_t)dcw * dch;
		for (y = 0; y < dch; ++y) memcpy(dst[2] + (ptrdiff_t)y * dst_strides[2], p + (size_t)y * dcw, (size_t)dcw);
	} else {
		int y;
		for (y = 0; y < c->dst_h; ++y)
			memcpy(dst[0] + (ptrdiff_t)y * dst_strides[0], p + (size_t)y * c->dst_w * 3, (size_t)c->dst_w * 3);
	}
	return 0;
}
static void b200_scaler_free(MSScalerContext *ctx) {
	B200ScalerCtx *c = (B200ScalerCtx *)ctx;
	msb200_ctx *dctx = msb200p_sync_ctx();
	msb200p_sync_lock();
	msb200_scaler_destroy(c->sc);
	if (dctx) {
		msb200_host_free_pinned(dctx, c->src_pack);
		msb200_host_free_pinned(dctx, c->dst_pack);
	}
	msb200p_sync_unlock();
	ms_free(c);
}
static MSScalerDesc b200_scaler_desc = {b200_scaler_create, b200_scaler_process, b200_scaler_free};

MSScalerDesc *msb200p_scaler_desc(void) {
	return &b200_scaler_desc;
}
void msb200p_register_video_filters(MSFactory *factory) {
	if (video_batch_capacity() > 0) { /* a member must run every tick to emit what it staged one tick earlier */
		b200_pix_conv_desc.flags |= MS_FILTER_IS_PUMP;
		b200_size_conv_desc.flags |= MS_FILTER_IS_PUMP;
	}
	ms_factory_register_filter(factory, &b200_pix_conv_desc);
	ms_factory_register_filter(factory, &b200_size_conv_desc);
}
/* lane statistics for benchmarks: flushes (= batched launches) and frames converted so far */
__attribute__((visibility("default"))) void msb200_filters_video_stats(unsigned long long *flushes, unsigned long long *frames) {
	pthread_mutex_lock(&g_vmu);
	if (flushes) *flushes = g_vflushes;
	if (frames) *frames = g_vframes;
	pthread_mutex_unlock(&g_vmu);
}
