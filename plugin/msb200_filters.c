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
 * Execution modes:
 *   synchronous (default) — each process() call makes one bank call for its own stream: exact reference semantics, no
 *     added latency, launch-bound (a few hundred streams per ticker thread at best);
 *   lockstep batch (MSB200_BATCH=<slots>) — the MSResample / MSSpeexEC / MSVolume / MSAudioMixer instances attached to
 *     one MSTicker with the same configuration share ONE bank ("batch group"). process() at tick T stages the
 *     filter's block into the group's pinned arena and emits the result of the block it staged at tick T-1; the first
 *     member called in a tick runs the whole group's previous tick in one H2D + one launch + one D2H. Every batched
 *     stage therefore adds one ticker interval (10 ms) of latency and nothing else: the samples are bit-identical to
 *     the synchronous mode's. See the "batch groups" section below and DESIGN.md §7.
 *
 * Compiled against the host's mediastreamer2 / oRTP / bctoolbox headers (here: /root/reference/include + compat/).
 */
#include "mediastreamer2/flowcontrol.h"
#include "mediastreamer2/msaudiomixer.h"
#include "mediastreamer2/mschanadapter.h"
#include "mediastreamer2/msequalizer.h"
#include "mediastreamer2/msfactory.h"
#include "mediastreamer2/msfilter.h"
#include "mediastreamer2/msgenericplc.h"
#include "mediastreamer2/msinterfaces.h"
#include "mediastreamer2/msticker.h"
#include "mediastreamer2/msvideo.h"
#include "mediastreamer2/msvolume.h"

#include "msb200dsp.h"
#include "msb200_plugin.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
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
/* taking the lock also makes the context's device current for the calling thread: process() runs on ticker threads that
 * never called cudaSetDevice (with MSB200_DEVICE != 0 their launches would otherwise target device 0) */
#define DSP_LOCK()                                                                                                     \
	do {                                                                                                               \
		pthread_mutex_lock(&g_mu);                                                                                     \
		if (g_ctx) msb200_ctx_make_current(g_ctx);                                                                     \
	} while (0)
#define DSP_UNLOCK() pthread_mutex_unlock(&g_mu)
#define DSP_CHECK(expr, what)                                                                                          \
	do {                                                                                                               \
		if ((expr) != MSB200_OK) ms_error("msb200: %s failed: %s", what, msb200_last_error());                         \
	} while (0)


/* ------------------------------------------------------------------------------------------------ batch groups
 * MSB200_BATCH=<slots> turns on the lockstep batch mode: filters of one kind, on one MSTicker, with one configuration
 * key share a bank of <slots> streams (rooms for the mixer). All member process() calls of a ticker come from that
 * ticker's thread (src/base/msticker.c:244-282), so a group's arenas and slot tables are touched by one thread at a
 * time; joining / leaving (preprocess / postprocess, on the attaching thread), control-thread methods that reach the
 * group's bank and the flush serialise on the GROUP's mutex. Every group owns its device context (its own CUDA
 * stream), so the groups of different tickers run concurrently on the GPU and never wait for each other on the host.
 *
 *   tick T, member i:  batch_tick()   -> first caller of the tick: run the bank over everything staged during T-1
 *                      emit           -> the member's results of T-1 go to its output queue
 *                      stage          -> the member's block(s) of T are copied into its arena slot
 *
 * Only slots [0, highest occupied + 1) are copied and processed (msb200_*_set_live). A slot that staged fewer units
 * than the group's maximum in a tick runs only its own units: MSVolume and MSSpeexEC take per-slot counts (streams
 * that joined at different ticks stage their 1-or-2 frames of 256 per 480-sample tick in different ticks). Only the
 * resampler bank is fed zeros for a starving stream's missing block, and that block's output is discarded.
 *
 * Several GPUs in one process: MSB200_DEVICES=<n> spreads the TICKERS over devices 0..n-1 in order of first use
 * (BASELINE cfg5: rooms are pinned to a GPU by giving their streams a ticker of that GPU); MSB200_DEVICE=<d> (default 0)
 * is the device of the synchronous filters and of every ticker when MSB200_DEVICES is unset.
 */
enum { BK_RESAMPLE, BK_EC, BK_VOLUME, BK_MIXER, BK_G711DEC, BK_G711ENC, BK_PLC }; /* the codecs are stateless: no bank, key[0] = law */
#define BATCH_MAX_SLOTS 4096
typedef struct Batch {
	struct Batch *next;
	int kind;
	MSTicker *ticker;
	int key[4];
	int cap, n_members, hi, live; /* hi: highest occupied slot + 1; live: what the bank was last told */
	void **owner;      /* [cap] filter state owning the slot, NULL when free */
	uint64_t seen_tick; /* ticker->ticks of the last flush */
	msb200_ctx *ctx;   /* the group's own device context (stream) */
	pthread_mutex_t mu; /* bank, slot tables */
	void *bank;
	int unit_in, unit_out, max_units; /* samples per unit per slot in / out; units a slot may stage per tick */
	int16_t *in[2], *out;             /* pinned arenas [cap][max_units * unit_*] */
	uint8_t *present;                 /* mixer: [cap][pins]; PLC: [cap][max_units] mode bytes (key[2] = max_units) */
	uint8_t *modes;                   /* PLC: [cap] the mode bytes of one unit, gathered for the launch */
	int *staged, *ready;              /* units per slot: staged this tick / ready from the last flush */
	int out_len;                      /* samples per unit the last run produced (resampler: frames out x channels) */
	/* moves a slot's ready results out of the arenas into mblks kept by the owner; the flush calls it for every member
	 * that has not been scheduled since the previous flush, so a skipped tick delays a stream's output but never loses it */
	void (*collect)(void *owner, struct Batch *b);
	uint64_t flushes, units_run;
} Batch;
static Batch *g_batches = NULL;
static pthread_mutex_t g_batch_mu = PTHREAD_MUTEX_INITIALIZER; /* the list of groups and the ticker -> device table */
static int g_batch_cap = -1;
#define GRP_LOCK(b) pthread_mutex_lock(&(b)->mu)
#define GRP_UNLOCK(b) pthread_mutex_unlock(&(b)->mu)

static int batch_capacity(void) {
	if (g_batch_cap < 0) {
		const char *e = getenv("MSB200_BATCH");
		int v = e ? atoi(e) : 0;
		g_batch_cap = v < 0 ? 0 : (v > BATCH_MAX_SLOTS ? BATCH_MAX_SLOTS : v);
	}
	return g_batch_cap;
}
/* device of a ticker's groups (g_batch_mu held) */
#define BATCH_MAX_TICKERS 256
static MSTicker *g_dev_tickers[BATCH_MAX_TICKERS];
static int g_n_dev_tickers = 0;
static int batch_device_of(MSTicker *t) {
	const char *e = getenv("MSB200_DEVICES"), *d = getenv("MSB200_DEVICE");
	const int ndev = e ? atoi(e) : 0;
	int i;
	if (ndev <= 1) return d ? atoi(d) : 0;
	for (i = 0; i < g_n_dev_tickers; ++i)
		if (g_dev_tickers[i] == t) return i % ndev;
	if (g_n_dev_tickers < BATCH_MAX_TICKERS) g_dev_tickers[g_n_dev_tickers++] = t;
	return (g_n_dev_tickers - 1) % ndev;
}
static void *batch_pinned(Batch *b, size_t bytes) {
	void *p = NULL;
	if (msb200_host_alloc_pinned(b->ctx, bytes ? bytes : 16, &p) != MSB200_OK) return NULL;
	memset(p, 0, bytes);
	return p;
}
static void batch_free(Batch *b) { /* unlinked, no members left */
	if (b->ctx) {
		msb200_ctx_make_current(b->ctx);
		switch (b->kind) {
			case BK_RESAMPLE: msb200_resample_destroy((msb200_resample *)b->bank); break;
			case BK_EC: msb200_aec_destroy((msb200_aec *)b->bank); break;
			case BK_VOLUME: msb200_volume_destroy((msb200_volume *)b->bank); break;
			case BK_MIXER: msb200_mixer_destroy((msb200_mixer *)b->bank); break;
			case BK_PLC: msb200_plc_destroy((msb200_plc *)b->bank); break;
		}
		if (b->in[0]) msb200_host_free_pinned(b->ctx, b->in[0]);
		if (b->in[1]) msb200_host_free_pinned(b->ctx, b->in[1]);
		if (b->out) msb200_host_free_pinned(b->ctx, b->out);
		msb200_ctx_destroy(b->ctx);
	}
	pthread_mutex_destroy(&b->mu);
	ms_free(b->present);
	ms_free(b->modes);
	ms_free(b->owner);
	ms_free(b->staged);
	ms_free(b->ready);
	ms_free(b);
}
/* find (or create) the group for (kind, ticker, key) and take a slot in it; NULL when batching is off or impossible */
static Batch *batch_join(int kind, MSTicker *ticker, const int key[4], int unit_in, int unit_out, int max_units, void *owner,
                         void (*collect)(void *, Batch *), int *slot) {
	Batch *b;
	int i, cap = batch_capacity();
	if (cap <= 0 || ticker == NULL) return NULL;
	pthread_mutex_lock(&g_batch_mu);
	for (b = g_batches; b; b = b->next)
		if (b->kind == kind && b->ticker == ticker && memcmp(b->key, key, sizeof(b->key)) == 0 && b->n_members < b->cap) break;
	if (!b) {
		int rc;
		b = ms_new0(Batch, 1);
		pthread_mutex_init(&b->mu, NULL);
		b->kind = kind;
		b->ticker = ticker;
		memcpy(b->key, key, sizeof(b->key));
		b->cap = cap;
		b->live = cap;
		b->unit_in = unit_in;
		b->unit_out = unit_out;
		b->max_units = max_units;
		b->collect = collect;
		b->seen_tick = (uint64_t)-1;
		rc = msb200_ctx_create(batch_device_of(ticker), &b->ctx);
		if (rc == MSB200_OK) {
			switch (kind) {
				case BK_RESAMPLE: rc = msb200_resample_create(b->ctx, cap, key[0], key[1], key[2], key[3], (msb200_resample **)&b->bank); break;
				case BK_EC: rc = msb200_aec_create(b->ctx, cap, key[0], key[1], key[2], (msb200_aec **)&b->bank); break;
				case BK_VOLUME: rc = msb200_volume_create(b->ctx, cap, key[0], key[1], (msb200_volume **)&b->bank); break;
				case BK_MIXER: rc = msb200_mixer_create(b->ctx, cap, key[2], key[0], key[1], (msb200_mixer **)&b->bank); break;
				case BK_PLC: rc = msb200_plc_create(b->ctx, cap, key[0], key[1], (msb200_plc **)&b->bank); break;
			}
		} else {
			b->ctx = NULL;
		}
		if (rc == MSB200_OK) {
			const size_t n_in = (size_t)cap * max_units * unit_in * sizeof(int16_t), n_out = (size_t)cap * max_units * unit_out * sizeof(int16_t);
			b->in[0] = (int16_t *)batch_pinned(b, n_in);
			if (kind == BK_EC) b->in[1] = (int16_t *)batch_pinned(b, n_in);
			b->out = (kind == BK_VOLUME || kind == BK_PLC) ? NULL : (int16_t *)batch_pinned(b, n_out); /* those two work in place */
			if (!b->in[0] || (kind == BK_EC && !b->in[1]) || (kind != BK_VOLUME && kind != BK_PLC && !b->out)) rc = MSB200_ENOMEM;
		}
		if (rc != MSB200_OK) {
			ms_error("msb200: cannot create a batch group (%s); the filter stays synchronous", msb200_last_error());
			batch_free(b);
			pthread_mutex_unlock(&g_batch_mu);
			return NULL;
		}
		b->owner = (void **)ms_new0(void *, cap);
		b->staged = ms_new0(int, cap);
		b->ready = ms_new0(int, cap);
		if (kind == BK_MIXER || kind == BK_PLC) b->present = (uint8_t *)ms_malloc0((size_t)cap * key[2]); /* PLC: [slot][unit] mode bytes */
		if (kind == BK_PLC) b->modes = (uint8_t *)ms_malloc0((size_t)cap);
		b->next = g_batches;
		g_batches = b;
		ms_message("msb200: batch group %p: kind %d, %d slots, key {%d,%d,%d,%d} on ticker %p", b, kind, cap, key[0], key[1], key[2], key[3], ticker);
	}
	GRP_LOCK(b);
	for (i = 0; i < b->cap && b->owner[i]; ++i) {
	}
	b->owner[i] = owner;
	b->staged[i] = b->ready[i] = 0;
	b->n_members++;
	if (i + 1 > b->hi) b->hi = i + 1;
	*slot = i;
	GRP_UNLOCK(b);
	pthread_mutex_unlock(&g_batch_mu);
	return b;
}
static void batch_leave(Batch *b, int slot) {
	Batch **pp;
	int last;
	if (!b) return;
	pthread_mutex_lock(&g_batch_mu);
	GRP_LOCK(b);
	b->owner[slot] = NULL;
	b->staged[slot] = b->ready[slot] = 0;
	if (b->present) memset(b->present + (size_t)slot * b->key[2], 0, (size_t)b->key[2]);
	while (b->hi > 0 && b->owner[b->hi - 1] == NULL)
		b->hi--;
	last = --b->n_members == 0;
	if (last) {
		for (pp = &g_batches; *pp && *pp != b; pp = &(*pp)->next) {
		}
		if (*pp) *pp = b->next;
	}
	GRP_UNLOCK(b);
	if (last) batch_free(b);
	pthread_mutex_unlock(&g_batch_mu);
}
/* called first thing in every member's process(): the first caller of a tick runs the group's previous tick */
static void batch_tick(Batch *b, uint64_t ticks) {
	int i, units = 0, rc = MSB200_OK;
	if (b->seen_tick == ticks) return;
	GRP_LOCK(b);
	b->seen_tick = ticks;
	for (i = 0; i < b->hi; ++i) {
		if (b->ready[i] > 0 && b->owner[i] && b->collect) b->collect(b->owner[i], b); /* not scheduled since the last flush */
		if (b->staged[i] > units) units = b->staged[i];
	}
	if (units > 0) {
		msb200_ctx_make_current(b->ctx);
		if (b->live != b->hi) {
			switch (b->kind) {
				case BK_RESAMPLE: msb200_resample_set_live((msb200_resample *)b->bank, b->hi); break;
				case BK_EC: msb200_aec_set_live((msb200_aec *)b->bank, b->hi); break;
				case BK_VOLUME: msb200_volume_set_live((msb200_volume *)b->bank, b->hi); break;
				case BK_MIXER: msb200_mixer_set_live((msb200_mixer *)b->bank, b->hi); break;
				case BK_PLC: msb200_plc_set_live((msb200_plc *)b->bank, b->hi); break;
			}
			b->live = b->hi;
		}
		/* slots that staged less than the group's maximum are fed zeros for the missing units */
		for (i = 0; i < b->hi; ++i) {
			if (b->staged[i] < units && b->kind == BK_RESAMPLE) { /* the one stateful bank without per-slot counts */
				const size_t off = ((size_t)i * b->max_units + b->staged[i]) * b->unit_in, n = (size_t)(units - b->staged[i]) * b->unit_in;
				memset(b->in[0] + off, 0, n * sizeof(int16_t));
				if (b->in[1]) memset(b->in[1] + off, 0, n * sizeof(int16_t));
			}
		}
		switch (b->kind) {
			case BK_RESAMPLE: {
				int outlen = 0;
				rc = msb200_resample_process((msb200_resample *)b->bank, b->in[0], b->key[3], b->out, b->unit_out / b->key[2], &outlen);
				b->out_len = outlen * b->key[2];
				break;
			}
			case BK_EC:
				/* per-slot frame counts: streams whose ticks fall differently against the frame grid stage 1 or 2 frames in
				 * different ticks; none of them is ever fed a made-up frame */
				rc = msb200_aec_process_counts((msb200_aec *)b->bank, b->in[0], b->in[1], b->out, units, b->max_units * b->unit_in, b->staged);
				b->out_len = b->unit_out;
				break;
			case BK_VOLUME:
				rc = msb200_volume_process_blocks((msb200_volume *)b->bank, b->in[0], b->unit_in, b->max_units * b->unit_in, units, b->staged);
				b->out_len = b->unit_in;
				break;
			case BK_MIXER:
				rc = msb200_mixer_process((msb200_mixer *)b->bank, b->in[0], b->present, b->out);
				b->out_len = b->unit_out;
				break;
			case BK_PLC: { /* one launch per unit; a unit in which no slot has device work (comfort noise only) is skipped */
				int u;
				for (u = 0; u < units && rc == MSB200_OK; ++u) {
					int any = 0;
					for (i = 0; i < b->hi; ++i) any |= (b->modes[i] = b->present[(size_t)i * b->key[2] + u]);
					if (any)
						rc = msb200_plc_process_strided((msb200_plc *)b->bank, b->in[0] + (size_t)u * b->unit_in, b->unit_in,
						                                b->max_units * b->unit_in, b->modes);
				}
				b->out_len = b->unit_in;
				break;
			}
			case BK_G711DEC: /* arenas are contiguous over the live slots: one flat batch (unstaged units decode garbage nobody reads) */
				rc = msb200_g711_decode(b->ctx, b->key[0], (const uint8_t *)b->in[0], b->out, (size_t)b->hi * b->max_units * b->unit_out);
				b->out_len = b->unit_out;
				break;
			case BK_G711ENC:
				rc = msb200_g711_encode(b->ctx, b->key[0], b->in[0], (uint8_t *)b->out, (size_t)b->hi * b->max_units * b->unit_in);
				b->out_len = b->unit_out;
				break;
		}
		if (rc != MSB200_OK) ms_error("msb200: batch group %p (kind %d) failed: %s", b, b->kind, msb200_last_error());
		b->flushes++;
		b->units_run += (uint64_t)units;
	}
	for (i = 0; i < b->hi; ++i) {
		b->ready[i] = rc == MSB200_OK ? b->staged[i] : 0;
		b->staged[i] = 0;
	}
	if (b->present) memset(b->present, 0, (size_t)b->hi * b->key[2]);
	GRP_UNLOCK(b);
}

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
	msb200_mixer *bank; /* 1 room x 50 pins x nwords; batch mode: the group's bank, this mixer is room `room` */
	int16_t *in;        /* [50][nwords] */
	uint8_t *present;   /* [50] */
	int16_t *out;       /* [50][nwords] (conference) or [nwords] */
	Batch *batch;       /* lockstep batch group (MSB200_BATCH), NULL in synchronous mode */
	int room;
	int pins;           /* pins of the bank: the highest connected pin + 1, rounded up to a multiple of 4 */
} MixerState;
/* the bank is shared with the group's flush in batch mode, with the other synchronous filters otherwise */
#define MIX_LOCK(s) do { if ((s)->batch) GRP_LOCK((s)->batch); else DSP_LOCK(); } while (0)
#define MIX_UNLOCK(s) do { if ((s)->batch) GRP_UNLOCK((s)->batch); else DSP_UNLOCK(); } while (0)

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
	s->room = 0;
	/* the graph is fixed while attached: only the connected pins travel to the GPU (a 16-party room moves 16 rows, not 50) */
	s->pins = 0;
	for (i = 0; i < MIXER_MAX_CHANNELS; ++i)
		if (f->inputs[i] || f->outputs[i]) s->pins = i + 1;
	s->pins = (s->pins + 3) & ~3;
	if (s->pins < 4) s->pins = 4;
	if (s->pins > MIXER_MAX_CHANNELS) s->pins = MIXER_MAX_CHANNELS;
	if (batch_capacity() > 0) {
		const int key[4] = {nwords, s->conf_mode, s->pins, 0};
		s->batch = batch_join(BK_MIXER, f->ticker, key, s->pins * nwords, s->conf_mode ? s->pins * nwords : nwords, 1, s, NULL, &s->room);
	}
	if (s->batch) { /* this mixer is one room of the group's bank; its arenas are slices of the group's pinned arenas */
		s->bank = (msb200_mixer *)s->batch->bank;
		s->in = s->batch->in[0] + (size_t)s->room * s->batch->unit_in;
		s->out = s->batch->out + (size_t)s->room * s->batch->unit_out;
		s->present = s->batch->present + (size_t)s->room * s->pins;
		GRP_LOCK(s->batch);
		for (i = 0; i < s->pins; ++i) {
			msb200_mixer_set_input_gain(s->bank, s->room, i, s->channels[i].gain);
			msb200_mixer_set_active(s->bank, s->room, i, s->channels[i].active);
		}
		GRP_UNLOCK(s->batch);
		return;
	}
	s->in = (int16_t *)ms_malloc0(sizeof(int16_t) * (size_t)s->pins * (size_t)nwords);
	s->out = (int16_t *)ms_malloc0(sizeof(int16_t) * (size_t)s->pins * (size_t)nwords);
	s->present = (uint8_t *)ms_malloc0((size_t)s->pins);
	DSP_LOCK();
	if (dsp_ctx()) {
		DSP_CHECK(msb200_mixer_create(g_ctx, 1, s->pins, nwords, s->conf_mode, &s->bank), "mixer_create");
		for (i = 0; s->bank && i < s->pins; ++i) {
			msb200_mixer_set_input_gain(s->bank, 0, i, s->channels[i].gain);
			msb200_mixer_set_active(s->bank, 0, i, s->channels[i].active);
		}
	}
	DSP_UNLOCK();
}
static void mixer_postprocess(MSFilter *f) {
	MixerState *s = (MixerState *)f->data;
	if (s->batch) {
		batch_leave(s->batch, s->room);
		s->batch = NULL;
		s->bank = NULL;
		s->in = s->out = NULL;
		s->present = NULL;
		return;
	}
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
/* one tick of results (s->out) to the output pins: channel_process_out :113-130 */
static void mixer_emit(MSFilter *f, MixerState *s, int nwords) {
	int i;
	if (s->conf_mode == 0) {
		mblk_t *om = NULL;
		for (i = 0; i < s->pins; ++i) {
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
		for (i = 0; i < s->pins; ++i) {
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
	if (s->batch) { /* the group's previous tick is computed by the first mixer called in this tick; emit this room's share */
		batch_tick(s->batch, f->ticker->ticks);
		if (s->batch->ready[s->room]) {
			s->batch->ready[s->room] = 0;
			mixer_emit(f, s, nwords);
		}
	}
	if (mixer_check_bypass(f, s)) {
		ms_filter_unlock(f);
		return;
	}
	memset(s->present, 0, (size_t)s->pins);
	for (i = 0; i < s->pins; ++i) {
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
	if (s->batch) { /* staged: the group's launch at the start of the next tick mixes every room at once */
		s->batch->staged[s->room] = 1;
		ms_filter_unlock(f);
		return;
	}
	/* the arithmetic: one launch for the whole mixer (sum, gains, minus-own, saturation) */
	DSP_LOCK();
	if (s->bank) DSP_CHECK(msb200_mixer_process(s->bank, s->in, s->present, s->out), "mixer_process");
	DSP_UNLOCK();
	if (s->bank) mixer_emit(f, s, nwords);
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
	if (s->bank && ctl->pin < s->pins) {
		MIX_LOCK(s);
		msb200_mixer_set_input_gain(s->bank, s->room, ctl->pin, ctl->param.gain);
		MIX_UNLOCK(s);
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
	if (s->bank && ctl->pin < s->pins) {
		MIX_LOCK(s);
		msb200_mixer_set_active(s->bank, s->room, ctl->pin, ctl->param.active);
		MIX_UNLOCK(s);
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
	Batch *batch;     /* lockstep batch group (MSB200_BATCH), light path only; joined at the first block */
	int slot;
	mblk_t *held[8];  /* the blocks staged in the current tick: processed in the arena, copied back and forwarded next tick */
	int n_held;
	MSQueue pend;     /* processed blocks waiting for this filter's next process() */
	bool_t batch_off;
} VolState;
static MSFilterDesc b200_volume_desc;
#define VOL_MAX_BLOCK 8192
#define VOL_BATCH_UNITS 8 /* blocks one stream may stage per tick (an upstream MSSpeexEC emits 1-2 frames per 10 ms) */

static void vol_sync_config_to(VolState *v, msb200_volume *bank, int st) { /* DSP lock held */
	if (!bank) return;
	if (v->peer && !v->peer_linked && bank == v->bank && v->peer->desc == &b200_volume_desc && ((VolState *)v->peer->data)->bank) {
		msb200_volume_set_peer(v->bank, 0, ((VolState *)v->peer->data)->bank, 0);
		v->peer_linked = TRUE;
	}
	if (v->gain_dirty) { /* MS_VOLUME_SET_GAIN resets the ramp (gain = target = static, msvolume.c:270-276): apply it once */
		msb200_volume_set_gain(bank, st, v->static_gain);
		v->gain_dirty = FALSE;
	}
	if (!v->dirty) return;
	if (v->noise_gate) {
		msb200_volume_enable_noise_gate(bank, st, 1);
		msb200_volume_set_noise_gate_threshold(bank, st, v->ng_threshold);
		msb200_volume_set_noise_gate_floorgain(bank, st, v->ng_floorgain);
	}
	msb200_volume_remove_dc(bank, st, v->remove_dc);
	msb200_volume_enable_agc(bank, st, v->agc);
	msb200_volume_set_ea_threshold(bank, st, v->ea_thres);
	msb200_volume_set_ea_speed(bank, st, v->ea_speed);
	msb200_volume_set_ea_force(bank, st, v->ea_force);
	msb200_volume_set_ea_sustain(bank, st, v->ea_sustain);
	msb200_volume_set_ea_transmit_threshold(bank, st, v->ea_transmit);
	v->dirty = FALSE;
}
static void vol_sync_config(VolState *v) { /* DSP lock held */
	vol_sync_config_to(v, v->bank, 0);
}
static void vol_leave_batch(VolState *v) {
	int i;
	if (v->batch) batch_leave(v->batch, v->slot);
	v->batch = NULL;
	for (i = 0; i < v->n_held; ++i) freemsg(v->held[i]);
	v->n_held = 0;
}
static void vol_collect(void *owner, Batch *b) { /* processed in place in the arena: copy back into the held blocks */
	VolState *v = (VolState *)owner;
	int u;
	for (u = 0; u < v->n_held; ++u) {
		if (b->ready[v->slot] == v->n_held) {
			memcpy(v->held[u]->b_rptr, b->in[0] + ((size_t)v->slot * b->max_units + u) * b->unit_in, (size_t)b->unit_in * 2);
			ms_queue_put(&v->pend, v->held[u]);
		} else { /* the group's launch failed (logged there): never forward unprocessed audio */
			freemsg(v->held[u]);
		}
	}
	b->ready[v->slot] = 0;
	v->n_held = 0;
}
static void vol_init(MSFilter *f) {
	VolState *v = ms_new0(VolState, 1);
	ms_queue_init(&v->pend);
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
	vol_leave_batch(v);
	ms_queue_flush(&v->pend);
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
static void vol_postprocess(MSFilter *f) {
	vol_leave_batch((VolState *)f->data);
}
static void vol_process(MSFilter *f) {
	VolState *v = (VolState *)f->data;
	mblk_t *m;
	if (v->agc || v->peer != NULL) { /* chunked mode :480-502 */
		if (v->batch) vol_leave_batch(v);
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
	if (v->batch && v->batch->key[0] != v->rate) vol_leave_batch(v);
	if (v->batch) { /* the blocks staged in the previous tick have been processed in the arena by the group's launch */
		batch_tick(v->batch, f->ticker->ticks);
		if (v->n_held && v->batch->staged[v->slot] == 0) vol_collect(v, v->batch);
	}
	while ((m = ms_queue_get(&v->pend)) != NULL)
		ms_queue_put(f->outputs[0], m);
	while ((m = ms_queue_get(f->inputs[0])) != NULL) {
		int n = (int)((m->b_wptr - m->b_rptr) / 2);
		if (!v->batch && !v->batch_off && batch_capacity() > 0 && n > 0 && n <= VOL_MAX_BLOCK) {
			const int key[4] = {v->rate, n, 0, 0};
			v->batch = batch_join(BK_VOLUME, f->ticker, key, n, n, VOL_BATCH_UNITS, v, vol_collect, &v->slot);
			if (v->batch) {
				GRP_LOCK(v->batch);
				msb200_ctx_make_current(v->batch->ctx);
				msb200_volume_reset_stream((msb200_volume *)v->batch->bank, v->slot);
				v->dirty = TRUE;
				v->gain_dirty = v->static_gain != 1.0f;
				GRP_UNLOCK(v->batch);
			} else {
				v->batch_off = TRUE;
			}
		}
		if (v->batch) {
			Batch *b = v->batch;
			if (n == b->key[1] && b->staged[v->slot] < b->max_units && v->n_held == b->staged[v->slot]) {
				if (v->dirty || v->gain_dirty) {
					GRP_LOCK(b);
					msb200_ctx_make_current(b->ctx);
					vol_sync_config_to(v, (msb200_volume *)b->bank, v->slot);
					GRP_UNLOCK(b);
				}
				memcpy(b->in[0] + ((size_t)v->slot * b->max_units + b->staged[v->slot]) * b->unit_in, m->b_rptr, (size_t)n * 2);
				b->staged[v->slot]++;
				v->held[v->n_held++] = m;
				continue;
			}
			ms_warning("MSVolume(b200): irregular block (%d samples, group block %d): leaving the batch group", n, b->key[1]);
			vol_leave_batch(v);
			v->batch_off = TRUE;
			v->dirty = TRUE;
			v->gain_dirty = v->static_gain != 1.0f;
		}
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
	if (v->batch) { /* the slot of the group's bank holds this stream's state */
		GRP_LOCK(v->batch);
		msb200_ctx_make_current(v->batch->ctx);
		rc = msb200_volume_get_state((msb200_volume *)v->batch->bank, v->slot, st) == MSB200_OK ? 0 : -1;
		GRP_UNLOCK(v->batch);
		return rc;
	}
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
                                        .postprocess = vol_postprocess,
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
	Batch *batch;    /* lockstep batch group (MSB200_BATCH); joined at the first block, keyed by rates, channels, block size */
	int slot;
	mblk_t *held;    /* the input block staged in the current tick: its meta data travel to the output block */
	MSQueue pend;    /* resampled blocks waiting for this filter's next process() */
	bool_t batch_off; /* irregular block sizes: this instance stays synchronous */
} RsState;
#define RS_MAX_FRAMES 8192
static void rs_init(MSFilter *f) {
	RsState *s = ms_new0(RsState, 1);
	s->input_rate = 8000;
	s->output_rate = 16000;
	s->in_nchannels = s->out_nchannels = 1;
	ms_queue_init(&s->pend);
	f->data = s;
}
static void rs_leave_batch(RsState *s) {
	if (s->batch) batch_leave(s->batch, s->slot);
	s->batch = NULL;
	if (s->held) freemsg(s->held);
	s->held = NULL;
}
static void rs_uninit(MSFilter *f) {
	RsState *s = (RsState *)f->data;
	rs_leave_batch(s);
	ms_queue_flush(&s->pend);
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
static void rs_collect(void *owner, Batch *b) { /* the slot's resampled block: arena -> mblk (meta data of the input block) */
	RsState *s = (RsState *)owner;
	if (s->held && b->ready[s->slot]) {
		const int outlen = b->out_len / s->in_nchannels;
		mblk_t *om = allocb((size_t)b->out_len * 2, 0);
		memcpy(om->b_wptr, b->out + (size_t)s->slot * b->unit_out, (size_t)b->out_len * 2);
		om->b_wptr += (size_t)b->out_len * 2;
		mblk_meta_copy(s->held, om);
		mblk_set_timestamp_info(om, s->ts);
		s->ts += (uint32_t)outlen;
		if (s->out_nchannels != s->in_nchannels) {
			ms_queue_put(&s->pend, rs_channel_adapt(s->in_nchannels, s->out_nchannels, om));
			freemsg(om);
		} else {
			ms_queue_put(&s->pend, om);
		}
	}
	if (s->held) freemsg(s->held);
	s->held = NULL;
	b->ready[s->slot] = 0;
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
	if (s->batch && (s->batch->key[0] != (int)s->input_rate || s->batch->key[1] != (int)s->output_rate || s->batch->key[2] != s->in_nchannels))
		rs_leave_batch(s); /* rates changed under us (:138-148): a new group is joined at the next block */
	if (s->batch) { /* the group's previous tick is computed by its first member called in this tick; emit our share */
		batch_tick(s->batch, f->ticker->ticks);
		if (s->held && s->batch->staged[s->slot] == 0) rs_collect(s, s->batch);
	}
	while ((im = ms_queue_get(&s->pend)) != NULL)
		ms_queue_put(f->outputs[0], im);
	while ((im = ms_queue_get(f->inputs[0])) != NULL) {
		int inlen = (int)((im->b_wptr - im->b_rptr) / (2 * s->in_nchannels));
		int outcap = (int)(((uint32_t)inlen * s->output_rate) / s->input_rate) + 1;
		int outlen = 0;
		mblk_t *om;
		if (!s->batch && !s->batch_off && batch_capacity() > 0 && inlen > 0 && inlen <= RS_MAX_FRAMES) {
			const int key[4] = {(int)s->input_rate, (int)s->output_rate, s->in_nchannels, inlen};
			s->batch = batch_join(BK_RESAMPLE, f->ticker, key, inlen * s->in_nchannels, outcap * s->in_nchannels, 1, s, rs_collect, &s->slot);
			if (s->batch) {
				GRP_LOCK(s->batch);
				msb200_ctx_make_current(s->batch->ctx);
				msb200_resample_reset_stream((msb200_resample *)s->batch->bank, s->slot);
				GRP_UNLOCK(s->batch);
			} else {
				s->batch_off = TRUE;
			}
		}
		if (s->batch) {
			Batch *b = s->batch;
			if (inlen == b->key[3] && b->staged[s->slot] == 0 && s->held == NULL) {
				memcpy(b->in[0] + (size_t)s->slot * b->unit_in, im->b_rptr, (size_t)b->unit_in * 2);
				b->staged[s->slot] = 1;
				s->held = im;
				continue;
			}
			ms_warning("MSResample(b200): irregular block (%d frames, group block %d): leaving the batch group", inlen, b->key[3]);
			rs_leave_batch(s);
			s->batch_off = TRUE;
		}
		om = allocb((size_t)outcap * 2 * (size_t)s->in_nchannels, 0);
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
static void rs_postprocess(MSFilter *f) {
	rs_leave_batch((RsState *)f->data);
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
                                          .postprocess = rs_postprocess,
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
	Batch *batch; /* lockstep batch group (MSB200_BATCH): frames are staged here and cancelled one tick later */
	int slot;
	MSQueue pend; /* cancelled frames waiting for this filter's next process() */
} EcState;
#define EC_BATCH_MAX_FRAMES 4 /* frames one stream may stage per tick (10 ms at 48 kHz = 1.875 frames of 256) */
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
	ms_queue_init(&s->pend);
	f->data = s;
}
static void ec_collect(void *owner, Batch *b) { /* the slot's cancelled frames: arena -> one mblk per frame */
	EcState *s = (EcState *)owner;
	const int nbytes = s->framesize * 2;
	int u;
	for (u = 0; u < b->ready[s->slot]; ++u) {
		mblk_t *oecho = allocb((size_t)nbytes, 0);
		memcpy(oecho->b_wptr, b->out + ((size_t)s->slot * b->max_units + u) * b->unit_out, (size_t)nbytes);
		oecho->b_wptr += nbytes;
		ms_queue_put(&s->pend, oecho);
	}
	b->ready[s->slot] = 0;
}
static void ec_uninit(MSFilter *f) {
	EcState *s = (EcState *)f->data;
	ms_queue_flush(&s->pend);
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
	if (batch_capacity() > 0) {
		const int key[4] = {s->samplerate, s->tail_length_ms, s->framesize_at_8000, 0};
		s->batch = batch_join(BK_EC, f->ticker, key, s->framesize, s->framesize, EC_BATCH_MAX_FRAMES, s, ec_collect, &s->slot);
		if (s->batch) {
			GRP_LOCK(s->batch);
			msb200_ctx_make_current(s->batch->ctx);
			msb200_aec_reset((msb200_aec *)s->batch->bank, s->slot);
			GRP_UNLOCK(s->batch);
		}
	}
	DSP_LOCK();
	if (!s->batch && dsp_ctx())
		DSP_CHECK(msb200_aec_create(g_ctx, 1, s->samplerate, s->tail_length_ms, s->framesize_at_8000, &s->bank), "aec_create");
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
	if (s->batch) batch_leave(s->batch, s->slot);
	s->batch = NULL;
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
	if (s->batch) { /* frames staged during the previous tick were cancelled by the group's launch: emit ours */
		batch_tick(s->batch, f->ticker->ticks);
		if (s->batch->staged[s->slot] == 0) ec_collect(s, s->batch);
	}
	while ((refm = ms_queue_get(&s->pend)) != NULL)
		ms_queue_put(f->outputs[1], refm);
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
		if (s->batch) { /* stage the (mic, delayed reference) frame pair; it is cancelled at the start of the next tick */
			Batch *b = s->batch;
			const int u = b->staged[s->slot];
			freemsg(oecho);
			if (u < b->max_units) {
				const size_t off = ((size_t)s->slot * b->max_units + u) * b->unit_in;
				memcpy(b->in[0] + off, echo, (size_t)nbytes);
				memcpy(b->in[1] + off, ref, (size_t)nbytes);
				b->staged[s->slot] = u + 1;
			} else {
				ms_warning("MSSpeexEC(b200): more than %d frames in one tick, frame dropped", b->max_units);
			}
			continue;
		}
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

/* ================================================================================================ G.711 codecs
 * MSAlawEnc / MSAlawDec / MSUlawEnc / MSUlawDec (/root/reference/src/audiofilters/alaw.c, ulaw.c): the "decode stub /
 * encode stub" either side of the audio path in a media server. Host logic restated from the
 * reference — encoder: MSBufferizer re-framing to ptime (alaw.c:56-94), fmtp / attr parsing (:96-138), getters
 * (:140-160); decoder: one output block per input block, meta data copied (:199-211). The companding itself runs on the
 * GPU (msb200_g711_*): one call for everything queued in the tick. */
#define G711_BATCH_UNITS 4 /* packets one stream may stage per tick */
typedef struct G711EncState {
	MSBufferizer *bz;
	int ptime, maxptime, law;
	uint32_t ts;
	Batch *batch; /* lockstep batch group (MSB200_BATCH), keyed by law and packet size */
	int slot, n_held;
	mblk_t *held[G711_BATCH_UNITS]; /* output packets staged in the current tick: meta data and timestamp set, payload pending */
	MSQueue pend;
	bool_t batch_off;
} G711EncState;
static void g711_enc_collect(void *owner, Batch *b) { /* encoded payloads: arena -> the held packets */
	G711EncState *s = (G711EncState *)owner;
	int u;
	for (u = 0; u < s->n_held; ++u) {
		if (b->ready[s->slot] == s->n_held) {
			const size_t nb = (size_t)b->unit_out * 2;
			memcpy(s->held[u]->b_wptr, (uint8_t *)b->out + ((size_t)s->slot * b->max_units + u) * nb, nb);
			s->held[u]->b_wptr += nb;
			ms_queue_put(&s->pend, s->held[u]);
		} else {
			freemsg(s->held[u]); /* the group's launch failed (logged there) */
		}
	}
	b->ready[s->slot] = 0;
	s->n_held = 0;
}
static void g711_enc_leave_batch(G711EncState *s) {
	int u;
	if (s->batch) batch_leave(s->batch, s->slot);
	s->batch = NULL;
	for (u = 0; u < s->n_held; ++u) freemsg(s->held[u]);
	s->n_held = 0;
}
static void g711_enc_init_law(MSFilter *f, int law) {
	G711EncState *s = ms_new0(G711EncState, 1);
	s->bz = ms_bufferizer_new();
	s->ptime = 0;
	s->maxptime = MS_DEFAULT_MAX_PTIME < 140 ? MS_DEFAULT_MAX_PTIME : 140;
	s->law = law;
	ms_queue_init(&s->pend);
	f->data = s;
}
static void alaw_enc_init(MSFilter *f) { g711_enc_init_law(f, MSB200_G711_ALAW); }
static void ulaw_enc_init(MSFilter *f) { g711_enc_init_law(f, MSB200_G711_ULAW); }
static void g711_enc_postprocess(MSFilter *f) {
	g711_enc_leave_batch((G711EncState *)f->data);
}
static void g711_enc_uninit(MSFilter *f) {
	G711EncState *s = (G711EncState *)f->data;
	g711_enc_leave_batch(s);
	ms_queue_flush(&s->pend);
	ms_bufferizer_destroy(s->bz);
	ms_free(s);
}
static void g711_enc_process(MSFilter *f) {
	G711EncState *s = (G711EncState *)f->data;
	int frame_per_packet = 2, npk, k;
	size_t size_of_pcm, avail;
	mblk_t *m;
	uint8_t *pcm, *code;
	if (s->ptime >= 10) frame_per_packet = s->ptime / 10;
	if (frame_per_packet <= 0) frame_per_packet = 1;
	if (frame_per_packet > 14) frame_per_packet = 14; /* 140 ms max */
	size_of_pcm = (size_t)160 * frame_per_packet;       /* bytes: 80 samples per 10 ms at 8 kHz */
	while ((m = ms_queue_get(f->inputs[0])) != NULL)
		ms_bufferizer_put(s->bz, m);
	if (s->batch && s->batch->key[1] != (int)size_of_pcm / 2) g711_enc_leave_batch(s); /* ptime changed under us */
	if (s->batch) { /* the packets staged in the previous tick were encoded by the group's launch */
		batch_tick(s->batch, f->ticker->ticks);
		if (s->n_held && s->batch->staged[s->slot] == 0) g711_enc_collect(s, s->batch);
	}
	while ((m = ms_queue_get(&s->pend)) != NULL)
		ms_queue_put(f->outputs[0], m);
	avail = ms_bufferizer_get_avail(s->bz);
	npk = (int)(avail / size_of_pcm);
	if (npk == 0) return;
	if (!s->batch && !s->batch_off && batch_capacity() > 0) {
		const int key[4] = {s->law, (int)size_of_pcm / 2, 0, 0};
		s->batch = batch_join(BK_G711ENC, f->ticker, key, (int)size_of_pcm / 2, (int)size_of_pcm / 4, G711_BATCH_UNITS, s, g711_enc_collect,
		                      &s->slot);
		if (!s->batch) s->batch_off = TRUE;
	}
	if (s->batch) { /* stage whole packets; their payload is filled in at the start of the next tick */
		Batch *b = s->batch;
		while (npk > 0 && b->staged[s->slot] < b->max_units && s->n_held == b->staged[s->slot]) {
			mblk_t *o = allocb(size_of_pcm / 2, 0);
			ms_bufferizer_read(s->bz, (uint8_t *)(b->in[0] + ((size_t)s->slot * b->max_units + b->staged[s->slot]) * b->unit_in), size_of_pcm);
			ms_bufferizer_fill_current_metas(s->bz, o);
			mblk_set_timestamp_info(o, s->ts);
			s->ts += (uint32_t)(size_of_pcm / 2);
			s->held[s->n_held++] = o;
			b->staged[s->slot]++;
			--npk;
		}
		return; /* packets beyond the per-tick capacity stay in the bufferizer for the next tick */
	}
	/* every complete packet of this tick in ONE device call; meta data are taken per packet, as the reference does */
	pcm = (uint8_t *)ms_malloc((size_t)npk * size_of_pcm);
	code = (uint8_t *)ms_malloc((size_t)npk * size_of_pcm / 2);
	{
		mblk_t **outs = (mblk_t **)ms_malloc(sizeof(mblk_t *) * (size_t)npk);
		int rc = MSB200_ENODEV;
		for (k = 0; k < npk; ++k) {
			ms_bufferizer_read(s->bz, pcm + (size_t)k * size_of_pcm, size_of_pcm);
			outs[k] = allocb(size_of_pcm / 2, 0);
			ms_bufferizer_fill_current_metas(s->bz, outs[k]);
		}
		DSP_LOCK();
		if (dsp_ctx()) {
			rc = msb200_g711_encode(g_ctx, s->law, (const int16_t *)pcm, code, (size_t)npk * size_of_pcm / 2);
			if (rc != MSB200_OK) ms_error("msb200: g711_encode failed: %s", msb200_last_error());
		}
		DSP_UNLOCK();
		for (k = 0; k < npk; ++k) {
			if (rc != MSB200_OK) { /* no GPU: never emit a packet that was not encoded */
				freemsg(outs[k]);
				continue;
			}
			memcpy(outs[k]->b_wptr, code + (size_t)k * size_of_pcm / 2, size_of_pcm / 2);
			outs[k]->b_wptr += size_of_pcm / 2;
			mblk_set_timestamp_info(outs[k], s->ts);
			s->ts += (uint32_t)(size_of_pcm / 2);
			ms_queue_put(f->outputs[0], outs[k]);
		}
		ms_free(outs);
	}
	ms_free(pcm);
	ms_free(code);
}
static int g711_enc_add_fmtp(MSFilter *f, void *arg) { /* enc_add_fmtp alaw.c:96-110 */
	const char *fmtp = (const char *)arg;
	G711EncState *s = (G711EncState *)f->data;
	char tmp[30];
	if (fmtp_get_value(fmtp, "maxptime", tmp, sizeof(tmp))) {
		int v = atoi(tmp);
		s->maxptime = v < MS_DEFAULT_MAX_PTIME ? v : MS_DEFAULT_MAX_PTIME;
	}
	if (fmtp_get_value(fmtp, "ptime", tmp, sizeof(tmp))) {
		int v = atoi(tmp);
		ms_message("%s configured with ptime=%s", f->desc->name, tmp);
		s->ptime = v < s->maxptime ? v : s->maxptime;
		if (s->ptime == s->maxptime) ms_message("%s ptime set to maxptime=%i", f->desc->name, s->maxptime);
	}
	return 0;
}
static int g711_enc_add_attr(MSFilter *f, void *arg) { /* enc_add_attr alaw.c:112-138: "ptime:<10..140 step 10>" */
	const char *attr = (const char *)arg;
	G711EncState *s = (G711EncState *)f->data;
	int p;
	/* the reference tests the strings in ascending order with strstr and takes the FIRST hit: "ptime:100" also matches
	 * "ptime:10" and therefore yields 10, like the original if/else chain */
	for (p = 10; p <= 140; p += 10) {
		char key[16];
		snprintf(key, sizeof(key), "ptime:%d", p);
		if (strstr(attr, key) != NULL) {
			s->ptime = p;
			break;
		}
	}
	return 0;
}
static int g711_get_sample_rate(MSFilter *f, void *arg) {
	(void)f;
	*(int *)arg = 8000;
	return 0;
}
static int g711_get_channels(MSFilter *f, void *arg) {
	(void)f;
	*(int *)arg = 1;
	return 0;
}
static int g711_get_ptime(MSFilter *f, void *arg) {
	*(int *)arg = ((G711EncState *)f->data)->ptime;
	return 0;
}
static int g711_dec_have_plc(MSFilter *f, void *arg) {
	(void)f;
	*(int *)arg = 0;
	return 0;
}
static MSFilterMethod g711_enc_methods[] = {{MS_FILTER_ADD_ATTR, g711_enc_add_attr},
                                            {MS_FILTER_ADD_FMTP, g711_enc_add_fmtp},
                                            {MS_FILTER_GET_NCHANNELS, g711_get_channels},
                                            {MS_FILTER_GET_SAMPLE_RATE, g711_get_sample_rate},
                                            {MS_AUDIO_ENCODER_GET_PTIME, g711_get_ptime},
                                            {0, NULL}};
static MSFilterMethod g711_dec_methods[] = {{MS_FILTER_GET_NCHANNELS, g711_get_channels},
                                            {MS_FILTER_GET_SAMPLE_RATE, g711_get_sample_rate},
                                            {MS_DECODER_HAVE_PLC, g711_dec_have_plc},
                                            {0, NULL}};
typedef struct G711DecState {
	int law;
	Batch *batch; /* lockstep batch group (MSB200_BATCH), keyed by law and payload size */
	int slot, n_held;
	mblk_t *held[G711_BATCH_UNITS]; /* output blocks staged in the current tick (meta data copied), samples pending */
	MSQueue pend;
	bool_t batch_off;
} G711DecState;
static void g711_dec_collect(void *owner, Batch *b) {
	G711DecState *s = (G711DecState *)owner;
	int u;
	for (u = 0; u < s->n_held; ++u) {
		if (b->ready[s->slot] == s->n_held) {
			const size_t nb = (size_t)b->unit_out * 2;
			memcpy(s->held[u]->b_wptr, b->out + ((size_t)s->slot * b->max_units + u) * b->unit_out, nb);
			s->held[u]->b_wptr += nb;
			ms_queue_put(&s->pend, s->held[u]);
		} else {
			freemsg(s->held[u]);
		}
	}
	b->ready[s->slot] = 0;
	s->n_held = 0;
}
static void g711_dec_leave_batch(G711DecState *s) {
	int u;
	if (s->batch) batch_leave(s->batch, s->slot);
	s->batch = NULL;
	for (u = 0; u < s->n_held; ++u) freemsg(s->held[u]);
	s->n_held = 0;
}
static void g711_dec_init_law(MSFilter *f, int law) {
	G711DecState *s = ms_new0(G711DecState, 1);
	s->law = law;
	ms_queue_init(&s->pend);
	f->data = s;
}
static void alaw_dec_init(MSFilter *f) { g711_dec_init_law(f, MSB200_G711_ALAW); }
static void ulaw_dec_init(MSFilter *f) { g711_dec_init_law(f, MSB200_G711_ULAW); }
static void g711_dec_postprocess(MSFilter *f) {
	g711_dec_leave_batch((G711DecState *)f->data);
}
static void g711_dec_uninit(MSFilter *f) {
	G711DecState *s = (G711DecState *)f->data;
	g711_dec_leave_batch(s);
	ms_queue_flush(&s->pend);
	ms_free(s);
}
/* batch mode: returns TRUE when the block was staged */
static bool_t g711_dec_stage(MSFilter *f, G711DecState *s, mblk_t *m) {
	const int n = (int)(m->b_wptr - m->b_rptr);
	if (!s->batch && !s->batch_off && batch_capacity() > 0 && n > 0 && (n % 2) == 0 && n <= 4096) {
		const int key[4] = {s->law, n, 0, 0};
		s->batch = batch_join(BK_G711DEC, f->ticker, key, n / 2, n, G711_BATCH_UNITS, s, g711_dec_collect, &s->slot);
		if (!s->batch) s->batch_off = TRUE;
	}
	if (!s->batch) return FALSE;
	{
		Batch *b = s->batch;
		if (n == b->key[1] && b->staged[s->slot] < b->max_units && s->n_held == b->staged[s->slot]) {
			mblk_t *o = allocb((size_t)n * 2, 0);
			mblk_meta_copy(m, o);
			memcpy((uint8_t *)(b->in[0] + ((size_t)s->slot * b->max_units + b->staged[s->slot]) * b->unit_in), m->b_rptr, (size_t)n);
			s->held[s->n_held++] = o;
			b->staged[s->slot]++;
			freemsg(m);
			return TRUE;
		}
		ms_warning("%s(b200): irregular payload (%d bytes, group payload %d): leaving the batch group", f->desc->name, n, b->key[1]);
		g711_dec_leave_batch(s);
		s->batch_off = TRUE;
	}
	return FALSE;
}
static void g711_dec_process_law(MSFilter *f, int law) { /* alaw_dec_process alaw.c:199-211 */
	G711DecState *st = (G711DecState *)f->data;
	mblk_t *m, *list[64];
	size_t total = 0, off = 0;
	int n = 0, k;
	uint8_t *code;
	int16_t *pcm;
	int rc = MSB200_ENODEV;
	if (st->batch) { /* the payloads staged in the previous tick were decoded by the group's launch */
		batch_tick(st->batch, f->ticker->ticks);
		if (st->n_held && st->batch->staged[st->slot] == 0) g711_dec_collect(st, st->batch);
	}
	while ((m = ms_queue_get(&st->pend)) != NULL)
		ms_queue_put(f->outputs[0], m);
	/* everything queued in this tick (usually one RTP payload) goes to the device in ONE call */
	while (n < 64 && (m = ms_queue_get(f->inputs[0])) != NULL) {
		msgpullup(m, (size_t)-1);
		if (g711_dec_stage(f, st, m)) continue; /* lockstep batch mode: decoded by the group, emitted next tick */
		list[n++] = m;
		total += (size_t)(m->b_wptr - m->b_rptr);
	}
	if (n == 0) return;
	code = (uint8_t *)ms_malloc(total ? total : 1);
	pcm = (int16_t *)ms_malloc(total ? total * 2 : 2);
	for (k = 0; k < n; ++k) {
		const size_t len = (size_t)(list[k]->b_wptr - list[k]->b_rptr);
		memcpy(code + off, list[k]->b_rptr, len);
		off += len;
	}
	DSP_LOCK();
	if (dsp_ctx()) {
		rc = msb200_g711_decode(g_ctx, law, code, pcm, total);
		if (rc != MSB200_OK) ms_error("msb200: g711_decode failed: %s", msb200_last_error());
	}
	DSP_UNLOCK();
	off = 0;
	for (k = 0; k < n; ++k) {
		const size_t len = (size_t)(list[k]->b_wptr - list[k]->b_rptr);
		if (rc == MSB200_OK) {
			mblk_t *o = allocb(len * 2, 0);
			mblk_meta_copy(list[k], o);
			memcpy(o->b_wptr, pcm + off, len * 2);
			o->b_wptr += len * 2;
			ms_queue_put(f->outputs[0], o);
		}
		off += len;
		freemsg(list[k]);
	}
	ms_free(code);
	ms_free(pcm);
}
static void alaw_dec_process(MSFilter *f) { g711_dec_process_law(f, MSB200_G711_ALAW); }
static void ulaw_dec_process(MSFilter *f) { g711_dec_process_law(f, MSB200_G711_ULAW); }
static MSFilterDesc b200_alaw_enc_desc = {.id = MS_ALAW_ENC_ID, .name = "MSAlawEnc", .text = "B200: ITU-G.711 alaw encoder (libmsb200dsp)",
                                          .category = MS_FILTER_ENCODER, .enc_fmt = "pcma", .ninputs = 1, .noutputs = 1,
                                          .init = alaw_enc_init, .process = g711_enc_process, .postprocess = g711_enc_postprocess, .uninit = g711_enc_uninit,
                                          .methods = g711_enc_methods};
static MSFilterDesc b200_ulaw_enc_desc = {.id = MS_ULAW_ENC_ID, .name = "MSUlawEnc", .text = "B200: ITU-G.711 ulaw encoder (libmsb200dsp)",
                                          .category = MS_FILTER_ENCODER, .enc_fmt = "pcmu", .ninputs = 1, .noutputs = 1,
                                          .init = ulaw_enc_init, .process = g711_enc_process, .postprocess = g711_enc_postprocess, .uninit = g711_enc_uninit,
                                          .methods = g711_enc_methods};
static MSFilterDesc b200_alaw_dec_desc = {.id = MS_ALAW_DEC_ID, .name = "MSAlawDec", .text = "B200: ITU-G.711 alaw decoder (libmsb200dsp)",
                                          .category = MS_FILTER_DECODER, .enc_fmt = "pcma", .ninputs = 1, .noutputs = 1,
                                          .init = alaw_dec_init, .process = alaw_dec_process, .postprocess = g711_dec_postprocess,
                                          .uninit = g711_dec_uninit, .methods = g711_dec_methods};
static MSFilterDesc b200_ulaw_dec_desc = {.id = MS_ULAW_DEC_ID, .name = "MSUlawDec", .text = "B200: ITU-G.711 ulaw decoder (libmsb200dsp)",
                                          .category = MS_FILTER_DECODER, .enc_fmt = "pcmu", .ninputs = 1, .noutputs = 1,
                                          .init = ulaw_dec_init, .process = ulaw_dec_process, .postprocess = g711_dec_postprocess,
                                          .uninit = g711_dec_uninit, .methods = g711_dec_methods};

/* ================================================================================================ MSAudioFlowControl
 * /root/reference/src/audiofilters/flowcontrol.c:152-279: filter shell (state, methods, drop request in ms -> samples
 * :196-207) on the host, ms_audio_flow_controller_process() :110-150 on the GPU (msb200_flowcontrol_*). Synchronous mode:
 * the controller is only armed for a few hundred ms after a drop event; a disarmed controller forwards blocks untouched
 * without any device call, exactly as the reference's `running` test does (:94-96). */
typedef struct FlowCtlState {
	msb200_flowcontrol *bank; /* 1 stream */
	int samplerate, nchannels, max_block;
	int strategy;
	float silent_threshold;
	bool_t armed; /* host mirror of ms_audio_flow_controller_running() */
} FlowCtlState;
#define FLOWCTL_MAX_BLOCK 8192
static void flowctl_init(MSFilter *f) {
	FlowCtlState *s = ms_new0(FlowCtlState, 1);
	s->strategy = MSB200_FLOWCONTROL_SOFT;
	s->silent_threshold = 0.02f;
	f->data = s;
}
static void flowctl_ensure_bank(FlowCtlState *s) { /* DSP lock held */
	if (s->bank || !dsp_ctx()) return;
	DSP_CHECK(msb200_flowcontrol_create(g_ctx, 1, FLOWCTL_MAX_BLOCK, &s->bank), "flowcontrol_create");
	if (s->bank) msb200_flowcontrol_set_config(s->bank, 0, s->strategy, s->silent_threshold);
}
static void flowctl_preprocess(MSFilter *f) { /* ms_audio_flow_controller_reset */
	FlowCtlState *s = (FlowCtlState *)f->data;
	DSP_LOCK();
	flowctl_ensure_bank(s);
	if (s->bank) msb200_flowcontrol_reset(s->bank, 0);
	DSP_UNLOCK();
	s->armed = FALSE;
}
static void flowctl_process(MSFilter *f) {
	FlowCtlState *s = (FlowCtlState *)f->data;
	mblk_t *m;
	ms_filter_lock(f);
	while ((m = ms_queue_get(f->inputs[0])) != NULL) {
		const int n = (int)((m->b_wptr - m->b_rptr) / 2);
		int32_t left = n;
		if (s->armed && s->bank && n >= 3 && n <= FLOWCTL_MAX_BLOCK) {
			msb200_flowcontrol_state st;
			int rc;
			DSP_LOCK();
			rc = msb200_flowcontrol_process(s->bank, (int16_t *)m->b_rptr, n, &left);
			if (rc == MSB200_OK) rc = msb200_flowcontrol_get_state(s->bank, 0, &st);
			DSP_UNLOCK();
			if (rc != MSB200_OK) {
				ms_error("msb200: flowcontrol_process failed: %s", msb200_last_error());
				left = n;
			} else {
				s->armed = st.total_samples > 0 && st.target_samples > 0;
			}
		}
		if (left <= 0) {
			freemsg(m);
			continue;
		}
		m->b_wptr = m->b_rptr + (size_t)left * 2;
		ms_queue_put(f->outputs[0], m);
	}
	ms_filter_unlock(f);
}
static void flowctl_uninit(MSFilter *f) {
	FlowCtlState *s = (FlowCtlState *)f->data;
	DSP_LOCK();
	msb200_flowcontrol_destroy(s->bank);
	DSP_UNLOCK();
	ms_free(s);
}
static int flowctl_set_config(MSFilter *f, void *arg) {
	FlowCtlState *s = (FlowCtlState *)f->data;
	const MSAudioFlowControlConfig *cfg = (const MSAudioFlowControlConfig *)arg;
	s->strategy = cfg->strategy == MSAudioFlowControlBasic ? MSB200_FLOWCONTROL_BASIC : MSB200_FLOWCONTROL_SOFT;
	s->silent_threshold = cfg->silent_threshold;
	ms_message("MSAudioFlowControl(b200): configured with strategy=[%i] and silent_threshold=[%f].", cfg->strategy, cfg->silent_threshold);
	DSP_LOCK();
	if (s->bank) msb200_flowcontrol_set_config(s->bank, 0, s->strategy, s->silent_threshold);
	DSP_UNLOCK();
	return 0;
}
static int flowctl_drop(MSFilter *f, void *arg) { /* ms_audio_flow_control_drop :196-207 */
	FlowCtlState *s = (FlowCtlState *)f->data;
	const MSAudioFlowControlDropEvent *ev = (const MSAudioFlowControlDropEvent *)arg;
	ms_filter_lock(f);
	if (!s->armed) {
		const uint32_t drop = (ev->drop_ms * (uint32_t)s->samplerate * (uint32_t)s->nchannels) / 1000;
		const uint32_t total = (ev->flow_control_interval_ms * (uint32_t)s->samplerate * (uint32_t)s->nchannels) / 1000;
		ms_message("MSAudioFlowControl(b200): requested to drop %i ms ", (int)ev->drop_ms);
		DSP_LOCK();
		flowctl_ensure_bank(s);
		if (s->bank && msb200_flowcontrol_set_target(s->bank, 0, drop, total) == MSB200_OK) s->armed = total > 0 && drop > 0;
		DSP_UNLOCK();
	}
	ms_filter_unlock(f);
	return 0;
}
static int flowctl_set_sr(MSFilter *f, void *arg) {
	((FlowCtlState *)f->data)->samplerate = *(int *)arg;
	return 0;
}
static int flowctl_get_sr(MSFilter *f, void *arg) {
	*(int *)arg = ((FlowCtlState *)f->data)->samplerate;
	return 0;
}
static int flowctl_set_nch(MSFilter *f, void *arg) {
	((FlowCtlState *)f->data)->nchannels = *(int *)arg;
	return 0;
}
static int flowctl_get_nch(MSFilter *f, void *arg) {
	*(int *)arg = ((FlowCtlState *)f->data)->nchannels;
	return 0;
}
static MSFilterMethod flowctl_methods[] = {{MS_AUDIO_FLOW_CONTROL_SET_CONFIG, flowctl_set_config},
                                           {MS_AUDIO_FLOW_CONTROL_DROP, flowctl_drop},
                                           {MS_FILTER_SET_SAMPLE_RATE, flowctl_set_sr},
                                           {MS_FILTER_GET_SAMPLE_RATE, flowctl_get_sr},
                                           {MS_FILTER_SET_NCHANNELS, flowctl_set_nch},
                                           {MS_FILTER_GET_NCHANNELS, flowctl_get_nch},
                                           {0, NULL}};
static MSFilterDesc b200_flow_control_desc = {.id = MS_AUDIO_FLOW_CONTROL_ID,
                                              .name = "MSAudioFlowControl",
                                              .text = "B200: flow control filter dropping samples when too many are queued (libmsb200dsp)",
                                              .category = MS_FILTER_OTHER,
                                              .ninputs = 1,
                                              .noutputs = 1,
                                              .init = flowctl_init,
                                              .preprocess = flowctl_preprocess,
                                              .process = flowctl_process,
                                              .uninit = flowctl_uninit,
                                              .methods = flowctl_methods};

/* ================================================================================================ MSGenericPLC
 * /root/reference/src/audiofilters/msgenericplc.c:44-157: filter shell and the concealer clock (MSConcealerContext,
 * src/base/mscommon.c:315-362, restated below: it is control logic) on the host; the signal work of every received block
 * (history, 5 ms continuity delay, cross-fade out of a concealed stretch) and of every concealed block
 * (genericplc.c:74-241) on the GPU (msb200_plc_*). Synchronous mode. Rates whose transform sizes need a radix above 5
 * (44.1 kHz) are refused loudly and the stream passes untouched. */
#define PLC_BATCH_UNITS 4 /* units a stream may stage per tick in a batch group: received blocks + one concealed block */
typedef struct PlcState {
	msb200_plc *bank; /* 1 stream (synchronous mode) */
	int bank_rate, bank_block;
	int rate, nchannels;
	int64_t sample_time, plc_start_time; /* MSConcealerContext (max_plc_time = UINT32_MAX, msgenericplc.c:44,51) */
	unsigned long total_plc;
	MSCngData cng_data;
	bool_t cng_set, cng_running, refused;
	/* lockstep batch mode (MSB200_BATCH): the stream's signal state lives in slot `slot` of the group's bank */
	Batch *batch;
	int slot, n_units;
	bool_t batch_off;
	mblk_t *unit_blk[PLC_BATCH_UNITS]; /* received block / prepared silence of each staged unit; NULL: concealed unit */
	uint8_t unit_kind[PLC_BATCH_UNITS]; /* 1 received, 2 concealed, 3 comfort-noise silence (no device work) */
	MSQueue pend;
} PlcState;
static void plc_init(MSFilter *f) {
	PlcState *s = ms_new0(PlcState, 1);
	s->nchannels = 1;
	s->sample_time = -1;
	s->plc_start_time = -1;
	ms_queue_init(&s->pend);
	f->data = s;
}
static void plc_ensure_bank(PlcState *s) { /* as the reference: the context is created once (:57-60) */
	if (s->bank || s->refused) return;
	{
		const int N = ((s->rate * 2 / 40) / 100) * 100, T = s->rate * 5 / 1000;
		s->bank_block = 2 * N - 2 * T;
		DSP_LOCK(); /* the context is created under the lock too: two filters may get here at once */
		if (!dsp_ctx()) {
			DSP_UNLOCK();
			return;
		}
		if (s->rate < 8000 || msb200_plc_create(g_ctx, 1, s->rate, s->bank_block, &s->bank) != MSB200_OK) {
			ms_error("MSGenericPLC(b200): no concealment at %d Hz: %s", s->rate, msb200_last_error());
			s->bank = NULL;
			s->refused = TRUE;
		}
		DSP_UNLOCK();
		s->bank_rate = s->rate;
	}
}
static void plc_preprocess(MSFilter *f) {
	PlcState *s = (PlcState *)f->data;
	if (batch_capacity() <= 0) plc_ensure_bank(s); /* batch mode: the private bank is only made if the stream leaves its group */
}
static void plc_device(PlcState *s, int16_t *io, int n, uint8_t mode) {
	int rc;
	plc_ensure_bank(s);
	if (!s->bank) return;
	if (n > s->bank_block) {
		ms_error("MSGenericPLC(b200): block of %d samples exceeds the %d-sample concealment window", n, s->bank_block);
		return;
	}
	DSP_LOCK();
	rc = msb200_plc_process(s->bank, io, n, &mode);
	DSP_UNLOCK();
	if (rc != MSB200_OK) ms_error("msb200: plc_process failed: %s", msb200_last_error());
}
/* results of the units staged in the previous tick: received blocks get their delayed / cross-faded samples back,
 * concealed blocks are created here, comfort-noise silence passes; everything goes to `pend` in staging order */
static void plc_collect(void *owner, Batch *b) {
	PlcState *s = (PlcState *)owner;
	const bool_t ok = b->ready[s->slot] == s->n_units;
	int u;
	for (u = 0; u < s->n_units; ++u) {
		const int16_t *row = b->in[0] + ((size_t)s->slot * b->max_units + u) * b->unit_in;
		mblk_t *m = s->unit_blk[u];
		if (s->unit_kind[u] == 1) {
			if (ok) {
				memcpy(m->b_rptr, row, (size_t)b->unit_in * 2);
				ms_queue_put(&s->pend, m);
			} else freemsg(m); /* the group's launch failed (logged there): never forward unprocessed audio */
		} else if (s->unit_kind[u] == 2) {
			if (ok) {
				m = allocb((size_t)b->unit_in * 2, 0);
				memcpy(m->b_wptr, row, (size_t)b->unit_in * 2);
				m->b_wptr += (size_t)b->unit_in * 2;
				mblk_set_plc_flag(m, 1);
				ms_queue_put(&s->pend, m);
			}
		} else {
			ms_queue_put(&s->pend, m);
		}
		s->unit_blk[u] = NULL;
	}
	b->ready[s->slot] = 0;
	s->n_units = 0;
}
static void plc_leave_batch(PlcState *s) {
	int u;
	if (s->batch) batch_leave(s->batch, s->slot);
	s->batch = NULL;
	for (u = 0; u < s->n_units; ++u)
		if (s->unit_blk[u]) freemsg(s->unit_blk[u]);
	s->n_units = 0;
}
/* stage one unit of this tick in the group's arena; FALSE: no room left (the stream then leaves the group) */
static bool_t plc_stage(PlcState *s, uint8_t kind, uint8_t mode, mblk_t *m) {
	Batch *b = s->batch;
	const int u = b->staged[s->slot];
	if (u >= b->max_units || u != s->n_units) return FALSE;
	if (kind == 1) memcpy(b->in[0] + ((size_t)s->slot * b->max_units + u) * b->unit_in, m->b_rptr, (size_t)b->unit_in * 2);
	b->present[(size_t)s->slot * b->key[2] + u] = mode;
	b->staged[s->slot] = u + 1;
	s->unit_blk[u] = m;
	s->unit_kind[u] = kind;
	s->n_units = u + 1;
	return TRUE;
}
static void plc_process(MSFilter *f) {
	PlcState *s = (PlcState *)f->data;
	const uint64_t now = f->ticker->time;
	const int tick_samples = s->rate * s->nchannels * f->ticker->interval / 1000;
	mblk_t *m;
	if (s->batch && s->batch->key[0] != s->rate) plc_leave_batch(s);
	if (s->batch) { /* the units staged in the previous tick have been processed in the arena by the group's launch */
		batch_tick(s->batch, f->ticker->ticks);
		if (s->n_units && s->batch->staged[s->slot] == 0) plc_collect(s, s->batch);
	}
	while ((m = ms_queue_get(&s->pend)) != NULL)
		ms_queue_put(f->outputs[0], m);
	while ((m = ms_queue_get(f->inputs[0])) != NULL) {
		size_t msg_size;
		uint8_t mode;
		if (m->b_cont) msgpullup(m, (size_t)-1);
		msg_size = (size_t)(m->b_wptr - m->b_rptr);
		/* ms_concealer_inc_sample_time(concealer, now, duration, TRUE) */
		if (s->sample_time == -1) s->sample_time = (int64_t)now;
		s->sample_time += (unsigned int)((1000 * msg_size) / ((size_t)s->rate * sizeof(int16_t) * (size_t)s->nchannels));
		s->plc_start_time = -1;
		mode = (uint8_t)(MSB200_PLC_PACKET | (s->cng_running ? MSB200_PLC_AFTER_CNG : 0));
		if (s->cng_running) {
			s->cng_running = FALSE;
			s->cng_set = FALSE;
		}
		/* batch mode takes streams whose packets last one ticker interval (the shape of the concealed blocks) */
		if (!s->batch && !s->batch_off && batch_capacity() > 0 && (int)(msg_size / 2) == tick_samples && tick_samples > 0) {
			const int N = ((s->rate * 2 / 40) / 100) * 100, T = s->rate * 5 / 1000;
			const int key[4] = {s->rate, tick_samples, PLC_BATCH_UNITS, 0};
			if (s->rate >= 8000 && tick_samples + 2 * T <= 2 * N)
				s->batch = batch_join(BK_PLC, f->ticker, key, tick_samples, tick_samples, PLC_BATCH_UNITS, s, plc_collect, &s->slot);
			if (s->batch) {
				GRP_LOCK(s->batch);
				msb200_ctx_make_current(s->batch->ctx);
				msb200_plc_reset_stream((msb200_plc *)s->batch->bank, s->slot);
				GRP_UNLOCK(s->batch);
				s->n_units = 0;
			} else {
				s->batch_off = TRUE;
			}
		}
		if (s->batch) {
			if ((int)(msg_size / 2) == s->batch->key[1] && plc_stage(s, 1, mode, m)) continue;
			ms_warning("MSGenericPLC(b200): irregular input (%d samples, group block %d): leaving the batch group",
			           (int)(msg_size / 2), s->batch->key[1]);
			plc_leave_batch(s);
			s->batch_off = TRUE;
		}
		plc_device(s, (int16_t *)m->b_rptr, (int)(msg_size / 2), mode);
		ms_queue_put(f->outputs[0], m);
	}
	/* ms_concealer_context_is_concealement_required(concealer, now) */
	if (s->sample_time != -1 && (uint64_t)s->sample_time <= now) {
		const unsigned int buff_size = (unsigned int)tick_samples * sizeof(int16_t);
		uint8_t kind;
		if (s->plc_start_time == -1) s->plc_start_time = s->sample_time;
		if ((uint32_t)(now - (uint64_t)s->plc_start_time) >= UINT32_MAX) {
			s->sample_time = -1;
			return;
		}
		s->total_plc++;
		s->sample_time += f->ticker->interval; /* ms_concealer_inc_sample_time(..., interval, FALSE) */
		if (s->cng_set) { /* comfort noise without a G.729B decoder is flagged silence (:131-141) */
			s->cng_set = FALSE;
			s->cng_running = TRUE;
			kind = 3;
		} else kind = s->cng_running ? 3 : 2;
		if (s->batch && kind == 2 && tick_samples == s->batch->key[1] && plc_stage(s, 2, MSB200_PLC_CONCEAL, NULL)) return;
		m = allocb(buff_size, 0);
		memset(m->b_wptr, 0, buff_size);
		m->b_wptr += buff_size;
		if (kind == 3) {
			mblk_set_cng_flag(m, 1);
			if (s->batch && plc_stage(s, 3, MSB200_PLC_IDLE, m)) return; /* keeps its place behind the units in flight */
		} else {
			mblk_set_plc_flag(m, 1);
		}
		if (s->batch) {
			ms_warning("MSGenericPLC(b200): no room for this tick's unit: leaving the batch group");
			plc_leave_batch(s);
			s->batch_off = TRUE;
		}
		if (kind == 2) plc_device(s, (int16_t *)m->b_rptr, tick_samples, MSB200_PLC_CONCEAL);
		ms_queue_put(f->outputs[0], m);
	}
}
static void plc_postprocess(MSFilter *f) { /* the group belongs to the ticker the filter is being detached from */
	PlcState *s = (PlcState *)f->data;
	plc_leave_batch(s);
	ms_queue_flush(&s->pend);
	s->batch_off = FALSE;
}
static void plc_uninit(MSFilter *f) {
	PlcState *s = (PlcState *)f->data;
	plc_leave_batch(s);
	ms_queue_flush(&s->pend);
	DSP_LOCK();
	msb200_plc_destroy(s->bank);
	DSP_UNLOCK();
	ms_free(s);
}
static int plc_get_sr(MSFilter *f, void *arg) {
	*(int *)arg = ((PlcState *)f->data)->rate;
	return 0;
}
static int plc_set_sr(MSFilter *f, void *arg) {
	((PlcState *)f->data)->rate = *(int *)arg;
	return 0;
}
static int plc_set_nch(MSFilter *f, void *arg) {
	((PlcState *)f->data)->nchannels = *(int *)arg;
	return 0;
}
static int plc_set_cn(MSFilter *f, void *arg) {
	PlcState *s = (PlcState *)f->data;
	memcpy(&s->cng_data, arg, sizeof(MSCngData));
	s->cng_set = TRUE;
	return 0;
}
static MSFilterMethod plc_methods[] = {{MS_FILTER_SET_SAMPLE_RATE, plc_set_sr},
                                       {MS_FILTER_GET_SAMPLE_RATE, plc_get_sr},
                                       {MS_FILTER_SET_NCHANNELS, plc_set_nch},
                                       {MS_GENERIC_PLC_SET_CN, plc_set_cn},
                                       {0, NULL}};
static MSFilterDesc b200_generic_plc_desc = {.id = MS_GENERIC_PLC_ID,
                                             .name = "MSGenericPLC",
                                             .text = "B200: generic packet-loss concealment (libmsb200dsp)",
                                             .category = MS_FILTER_OTHER,
                                             .ninputs = 1,
                                             .noutputs = 1,
                                             .init = plc_init,
                                             .preprocess = plc_preprocess,
                                             .process = plc_process,
                                             .postprocess = plc_postprocess,
                                             .uninit = plc_uninit,
                                             .methods = plc_methods,
                                             .flags = MS_FILTER_IS_PUMP};

/* ================================================================================================ entry point */
__attribute__((visibility("default"))) void libmsb200filters_init(MSFactory *factory) {
	if (batch_capacity() > 0) {
		/* lockstep batch mode: a filter must be called every tick to emit the result of the block it staged one tick
		 * earlier, whether or not new input arrived (MSAudioMixer and MSChannelAdapter are pumps already) */
		b200_volume_desc.flags |= MS_FILTER_IS_PUMP;
		b200_resample_desc.flags |= MS_FILTER_IS_PUMP;
		b200_speex_ec_desc.flags |= MS_FILTER_IS_PUMP;
		b200_alaw_enc_desc.flags |= MS_FILTER_IS_PUMP;
		b200_alaw_dec_desc.flags |= MS_FILTER_IS_PUMP;
		b200_ulaw_enc_desc.flags |= MS_FILTER_IS_PUMP;
		b200_ulaw_dec_desc.flags |= MS_FILTER_IS_PUMP;
		ms_message("libmsb200filters: lockstep batch mode, %d slots per group (MSB200_BATCH)", batch_capacity());
	}
	ms_factory_register_filter(factory, &b200_audio_mixer_desc);
	ms_factory_register_filter(factory, &b200_volume_desc);
	ms_factory_register_filter(factory, &b200_channel_adapter_desc);
	ms_factory_register_filter(factory, &b200_equalizer_desc);
	ms_factory_register_filter(factory, &b200_resample_desc);
	ms_factory_register_filter(factory, &b200_speex_ec_desc);
	ms_factory_register_filter(factory, &b200_alaw_enc_desc);
	ms_factory_register_filter(factory, &b200_alaw_dec_desc);
	ms_factory_register_filter(factory, &b200_ulaw_enc_desc);
	ms_factory_register_filter(factory, &b200_ulaw_dec_desc);
	ms_factory_register_filter(factory, &b200_flow_control_desc);
	ms_factory_register_filter(factory, &b200_generic_plc_desc);
	msb200p_register_video_filters(factory);
	if (getenv("MSB200_INSTALL_SCALER")) ms_video_set_scaler_impl(msb200p_scaler_desc());
	ms_message("libmsb200filters: B200 DSP filters registered (MSAudioMixer, MSVolume, MSChannelAdapter, MSEqualizer, "
	           "MSResample, MSSpeexEC, MSAlawEnc/Dec, MSUlawEnc/Dec, MSAudioFlowControl, MSGenericPLC, MSPixConv, MSSizeConv%s)",
	           getenv("MSB200_INSTALL_SCALER") ? ", MSScaler" : "");
}
/* batch-group statistics for benchmarks: groups, launches (flushes) and units run so far, summed over all groups */
__attribute__((visibility("default"))) void msb200_filters_batch_stats(int *groups, unsigned long long *flushes, unsigned long long *units) {
	Batch *b;
	int g = 0;
	unsigned long long fl = 0, un = 0;
	pthread_mutex_lock(&g_batch_mu);
	for (b = g_batches; b; b = b->next) {
		g++;
		fl += b->flushes;
		un += b->units_run;
	}
	pthread_mutex_unlock(&g_batch_mu);
	if (groups) *groups = g;
	if (flushes) *flushes = fl;
	if (units) *units = un;
}
/* also exported so that a host can install the scaler explicitly */
__attribute__((visibility("default"))) MSScalerDesc *msb200_ms_scaler_desc(void) {
	return msb200p_scaler_desc();
}
/* ---- shared with the other translation units of the plugin (msb200_plugin.h) */
msb200_ctx *msb200p_sync_ctx(void) {
	msb200_ctx *c;
	DSP_LOCK();
	c = dsp_ctx();
	if (c) msb200_ctx_make_current(c); /* the calling thread may be a ticker that never touched this device (MSB200_DEVICE != 0) */
	DSP_UNLOCK();
	return c;
}
void msb200p_sync_lock(void) {
	DSP_LOCK();
	if (g_ctx) msb200_ctx_make_current(g_ctx);
}
void msb200p_sync_unlock(void) {
	DSP_UNLOCK();
}
int msb200p_batch_capacity(void) {
	return batch_capacity();
}
int msb200p_device_of_ticker(MSTicker *t) {
	int d;
	pthread_mutex_lock(&g_batch_mu);
	d = batch_device_of(t);
	pthread_mutex_unlock(&g_batch_mu);
	return d;
}
