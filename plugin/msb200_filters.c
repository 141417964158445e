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
#include <time.h>

/* the one-in / one-out filters' queue ends */
static inline mblk_t *pin0_next(MSFilter *f) {
	return ms_queue_get(f->inputs[0]);
}
static inline void pin0_send(MSFilter *f, mblk_t *m) {
	ms_queue_put(f->outputs[0], m);
}

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


/* ------------------------------------------------------------------------------------------------ host profile
 * MSB200_PROFILE=1: where a ticker thread's time goes, per filter kind — nanoseconds inside process() (the flush a first
 * caller runs is counted apart: enqueue and wait). Striped counters, read by msb200_filters_host_profile(). */
enum { PF_MIXER, PF_VOLUME, PF_CHAN, PF_EQ, PF_RESAMPLE, PF_EC, PF_G711ENC, PF_G711DEC, PF_FLOWCTL, PF_PLC, PF_FLUSH_ENQUEUE, PF_FLUSH_WAIT, PF_N };
static const char *const g_prof_names[PF_N] = {"MSAudioMixer", "MSVolume", "MSChannelAdapter", "MSEqualizer", "MSResample", "MSSpeexEC",
                                               "G711Enc", "G711Dec", "MSAudioFlowControl", "MSGenericPLC", "flush_enqueue", "flush_wait"};
#define PROF_STRIPES 64
typedef struct ProfStripe {
	uint64_t ns[PF_N], calls[PF_N];
	char pad[64];
} ProfStripe;
static ProfStripe g_prof[PROF_STRIPES];
static int g_prof_on = -1;
static __thread uint64_t t_flush_ns = 0; /* flush time inside the process() call in progress */
static __thread int t_stripe = -1;
static inline int prof_on(void) {
	if (g_prof_on < 0) {
		const char *e = getenv("MSB200_PROFILE");
		g_prof_on = e && atoi(e) > 0;
	}
	return g_prof_on;
}
static inline uint64_t prof_now(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
static void prof_add(int kind, uint64_t ns) {
	static int next_stripe = 0;
	if (t_stripe < 0) t_stripe = __atomic_fetch_add(&next_stripe, 1, __ATOMIC_RELAXED) % PROF_STRIPES;
	__atomic_fetch_add(&g_prof[t_stripe].ns[kind], ns, __ATOMIC_RELAXED);
	__atomic_fetch_add(&g_prof[t_stripe].calls[kind], 1, __ATOMIC_RELAXED);
}
#define PROF_WRAP(kind, fn)                                                                                            \
	static void fn##_timed(MSFilter *f) {                                                                              \
		uint64_t t0, fl0;                                                                                              \
		if (!prof_on()) {                                                                                              \
			fn(f);                                                                                                     \
			return;                                                                                                    \
		}                                                                                                              \
		t0 = prof_now();                                                                                               \
		fl0 = t_flush_ns;                                                                                              \
		fn(f);                                                                                                         \
		prof_add(kind, prof_now() - t0 - (t_flush_ns - fl0));                                                          \
	}

/* ------------------------------------------------------------------------------------------------ batch groups
 * MSB200_BATCH=<slots> turns on the lockstep batch mode: filters of one kind, on one MSTicker, with one configuration
 * key share a bank of <slots> streams (rooms for the mixer). All member process() calls of a ticker come from that
 * ticker's thread (src/base/msticker.c:244-282), so a group's arenas and slot tables are touched by one thread at a
 * time; joining / leaving (preprocess / postprocess, on the attaching thread), control-thread methods that reach the
 * group's bank and the flush serialise on the GROUP's mutex. Every TICKER owns one device context (one CUDA stream)
 * shared by all its groups, so the tickers run concurrently on the GPU and never wait for each other on the host, and a
 * ticker waits for the device ONCE per tick: the first member called in tick T enqueues the work of ALL the ticker's groups
 * back to back (copies in, kernels, copies out: msb200_ctx_set_deferred_sync) and synchronises once.
 *
 *   tick T, first member of the ticker:  flush -> every group's bank runs over everything staged during T-1
 *   tick T, member i:  emit           -> the member's results of T-1 go to its output queue
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
struct Batch;
typedef struct TickerDev { /* a ticker's device context and its groups */
	struct TickerDev *next;
	MSTicker *ticker;
	msb200_ctx *ctx;
	pthread_mutex_t mu;    /* the group list and the joint flush */
	struct Batch *groups;  /* through Batch.tnext */
	int n_groups;
	uint64_t seen_tick;    /* ticker->ticks of the last flush */
	uint64_t flushes;
} TickerDev;
typedef struct Batch {
	struct Batch *next, *tnext;
	TickerDev *td;
	int kind;
	MSTicker *ticker;
	int key[4];
	int cap, n_members, hi, live; /* hi: highest occupied slot + 1; live: what the bank was last told */
	void **owner;      /* [cap] filter state owning the slot, NULL when free */
	msb200_ctx *ctx;   /* the ticker's device context (td->ctx) */
	int run_units, run_rc; /* of the flush in progress */
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
static TickerDev *g_ticker_devs = NULL; /* g_batch_mu */
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
/* the ticker's device context (g_batch_mu held); created with the ticker's first group, destroyed with its last */
static TickerDev *ticker_dev(MSTicker *ticker) {
	TickerDev *td;
	for (td = g_ticker_devs; td; td = td->next)
		if (td->ticker == ticker) return td;
	td = ms_new0(TickerDev, 1);
	td->ticker = ticker;
	td->seen_tick = (uint64_t)-1;
	if (msb200_ctx_create(batch_device_of(ticker), &td->ctx) != MSB200_OK) {
		ms_free(td);
		return NULL;
	}
	pthread_mutex_init(&td->mu, NULL);
	td->next = g_ticker_devs;
	g_ticker_devs = td;
	return td;
}
static void ticker_dev_release(TickerDev *td) { /* g_batch_mu held */
	TickerDev **tp;
	if (td->n_groups > 0) return;
	for (tp = &g_ticker_devs; *tp && *tp != td; tp = &(*tp)->next) {
	}
	if (*tp) *tp = td->next;
	msb200_ctx_destroy(td->ctx);
	pthread_mutex_destroy(&td->mu);
	ms_free(td);
}
static void batch_free(Batch *b) { /* g_batch_mu held; unlinked from g_batches, no members left */
	TickerDev *td = b->td;
	if (td) { /* out of the ticker's flush first (a flush in progress finishes under td->mu) */
		Batch **gp;
		pthread_mutex_lock(&td->mu);
		for (gp = &td->groups; *gp && *gp != b; gp = &(*gp)->tnext) {
		}
		if (*gp) {
			*gp = b->tnext;
			td->n_groups--;
		}
		pthread_mutex_unlock(&td->mu);
	}
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
	}
	if (td) ticker_dev_release(td); /* the last group takes the context with it */
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
		int rc = MSB200_ECUDA;
		TickerDev *td = ticker_dev(ticker);
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
		if (td) {
			b->ctx = td->ctx;
			msb200_ctx_make_current(b->ctx);
			rc = MSB200_OK;
			switch (kind) {
				case BK_RESAMPLE: rc = msb200_resample_create(b->ctx, cap, key[0], key[1], key[2], key[3], (msb200_resample **)&b->bank); break;
				case BK_EC: rc = msb200_aec_create(b->ctx, cap, key[0], key[1], key[2], (msb200_aec **)&b->bank); break;
				case BK_VOLUME: rc = msb200_volume_create(b->ctx, cap, key[0], key[1], (msb200_volume **)&b->bank); break;
				case BK_MIXER: rc = msb200_mixer_create(b->ctx, cap, key[2], key[0], key[1], (msb200_mixer **)&b->bank); break;
				case BK_PLC: rc = msb200_plc_create(b->ctx, cap, key[0], key[1], (msb200_plc **)&b->bank); break;
			}
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
			if (td) ticker_dev_release(td);
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
		pthread_mutex_lock(&td->mu); /* from now on the ticker's flush runs this group too */
		b->td = td;
		b->tnext = td->groups;
		td->groups = b;
		td->n_groups++;
		pthread_mutex_unlock(&td->mu);
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
/* one group's share of the ticker's flush (group lock held, deferred synchronisation on): everything staged since the
 * last flush is enqueued on the ticker's stream — copies in, kernels, copies out */
static void batch_enqueue(Batch *b) {
	int i, units = 0, rc = MSB200_OK;
	for (i = 0; i < b->hi; ++i) {
		if (b->ready[i] > 0 && b->owner[i] && b->collect) b->collect(b->owner[i], b); /* not scheduled since the last flush */
		if (b->staged[i] > units) units = b->staged[i];
	}
	if (units > 0) {
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
	b->run_units = units;
	b->run_rc = rc;
}
/* after the ticker's one synchronisation: the staged units are now results */
static void batch_finish(Batch *b, int sync_rc) {
	int i;
	const int ok = b->run_rc == MSB200_OK && sync_rc == MSB200_OK;
	for (i = 0; i < b->hi; ++i) {
		b->ready[i] = ok ? b->staged[i] : 0;
		b->staged[i] = 0;
	}
	if (b->present) memset(b->present, 0, (size_t)b->hi * b->key[2]);
}
/* called first thing in every member's process(), and by a filter that joins a group from its process() BEFORE it stages
 * its first block (a block staged ahead of the tick's flush would run a tick early and leave the slot empty — for the
 * resampler: zero-fed — at the next one): the first caller of a tick, whatever its group, runs the previous tick of ALL
 * the ticker's groups — their device work goes out back to back on the ticker's stream and the thread waits once */
static void batch_tick(Batch *b, uint64_t ticks) {
	TickerDev *td = b->td;
	Batch *g;
	int rc;
	uint64_t t0 = 0, t1 = 0;
	if (td->seen_tick == ticks) return; /* (only the ticker's own thread writes it) */
	pthread_mutex_lock(&td->mu);
	td->seen_tick = ticks;
	td->flushes++;
	if (prof_on()) t0 = prof_now();
	msb200_ctx_make_current(td->ctx);
	msb200_ctx_set_deferred_sync(td->ctx, 1);
	for (g = td->groups; g; g = g->tnext) {
		GRP_LOCK(g);
		batch_enqueue(g);
	}
	if (prof_on()) t1 = prof_now();
	rc = msb200_ctx_set_deferred_sync(td->ctx, 0); /* synchronises */
	if (prof_on()) {
		const uint64_t t2 = prof_now();
		prof_add(PF_FLUSH_ENQUEUE, t1 - t0);
		prof_add(PF_FLUSH_WAIT, t2 - t1);
		t_flush_ns += t2 - t0;
	}
	if (rc != MSB200_OK) ms_error("msb200: flush of ticker %p failed: %s", td->ticker, msb200_last_error());
	for (g = td->groups; g; g = g->tnext) {
		batch_finish(g, rc);
		GRP_UNLOCK(g);
	}
	pthread_mutex_unlock(&td->mu);
}


/* ================================================================================================ MSAudioMixer
 * What the reference filter does on the host and this one must do too (/root/reference/src/audiofilters/audiomixer.c):
 * per-pin FIFOs read one tick at a time (:78-90), a drift trim every 5 s of ticker time (:92-111), the single-talker
 * shortcut that forwards packets untouched (:219-286) and the hand-out of one block per listener (:313-343). The sums,
 * gains, minus-own and saturation (:33-51, :113-130) run in mixer_kernel. The host side below is organised around the
 * device arenas: a room is a slice of pinned memory [pin][samples] that the pins' FIFOs are drained INTO, and a slice that
 * the listeners' blocks are cut FROM — in batch mode those are slices of the group's arenas, so staging is the read itself. */
#define MIX_PINS 50          /* pins of the reference desc (MIXER_MAX_CHANNELS, audiomixer.c:29) */
#define MIX_SOLO_AFTER_MS 1000 /* a pin silent this long stops counting as a talker (BYPASS_MODE_TIMEOUT, :31) */
#define MIX_TRIM_EVERY_MS 5000
#define MIX_NEVER ((uint64_t)-1)

typedef struct MixPin {
	MSBufferizer fifo;
	float gain;          /* MS_AUDIO_MIXER_SET_INPUT_GAIN */
	bool_t contributes;  /* MS_AUDIO_MIXER_SET_ACTIVE */
	bool_t listens;      /* MS_AUDIO_MIXER_ENABLE_OUTPUT */
	uint64_t heard_at;   /* ticker time of the pin's last packet */
	uint64_t trimmed_at; /* ticker time of the last drift check */
	int low_water;       /* smallest FIFO fill seen since then, -1 = none yet */
} MixPin;

typedef struct MixRoom {
	MixPin pin[MIX_PINS];
	int channels, hz, conference, master_pin;
	int tick_bytes, trim_above; /* bytes per pin per tick; a FIFO that never dips below trim_above is trimmed */
	bool_t solo, one_listener;
	msb200_mixer *bank; /* 1 room x n_dev_pins x words; batch mode: the group's bank, this mixer is room `room` */
	int16_t *in;        /* [n_dev_pins][words] */
	uint8_t *present;   /* [n_dev_pins] */
	int16_t *out;       /* [n_dev_pins][words] (conference) or [words] */
	Batch *batch;       /* lockstep batch group (MSB200_BATCH), NULL in synchronous mode */
	int room;
	int n_dev_pins;     /* pins that travel to the device: highest connected pin + 1, rounded up to a multiple of 4 */
} MixRoom;
/* the bank is shared with the group's flush in batch mode, with the other synchronous filters otherwise */
#define MIX_LOCK(s) do { if ((s)->batch) GRP_LOCK((s)->batch); else DSP_LOCK(); } while (0)
#define MIX_UNLOCK(s) do { if ((s)->batch) GRP_UNLOCK((s)->batch); else DSP_UNLOCK(); } while (0)

static void mixroom_new(MSFilter *f) {
	MixRoom *r = ms_new0(MixRoom, 1);
	MixPin *p;
	r->channels = 1;
	r->hz = 44100;
	r->master_pin = -1;
	for (p = r->pin; p < r->pin + MIX_PINS; ++p) {
		ms_bufferizer_init(&p->fifo);
		p->gain = 1.0f;
		p->contributes = p->listens = TRUE;
	}
	f->data = r;
}
static void mixroom_free(MSFilter *f) {
	MixRoom *r = (MixRoom *)f->data;
	MixPin *p;
	for (p = r->pin; p < r->pin + MIX_PINS; ++p) ms_bufferizer_uninit(&p->fifo);
	ms_free(r);
}
static int mixroom_listeners(const MSFilter *f, const MixRoom *r) {
	int k, n = 0;
	for (k = 0; k < MIX_PINS; ++k) n += f->outputs[k] != NULL && r->pin[k].listens;
	return n;
}
static void mixroom_push_controls(MixRoom *r) { /* bank lock held */
	int k;
	for (k = 0; r->bank && k < r->n_dev_pins; ++k) {
		msb200_mixer_set_input_gain(r->bank, r->room, k, r->pin[k].gain);
		msb200_mixer_set_active(r->bank, r->room, k, r->pin[k].contributes);
	}
}
static void mixroom_attach(MSFilter *f) {
	MixRoom *r = (MixRoom *)f->data;
	int k, words, top = 0;
	r->tick_bytes = (2 * r->channels * r->hz * f->ticker->interval) / 1000;
	r->trim_above = 2 * r->tick_bytes;
	words = r->tick_bytes / 2;
	r->solo = FALSE;
	r->one_listener = mixroom_listeners(f, r) == 1;
	r->room = 0;
	for (k = 0; k < MIX_PINS; ++k) {
		r->pin[k].heard_at = r->pin[k].trimmed_at = MIX_NEVER;
		if (f->inputs[k] || f->outputs[k]) top = k + 1;
	}
	/* the graph is fixed while attached: only the connected pins travel (a 16-party room moves 16 rows, not 50) */
	r->n_dev_pins = (top + 3) & ~3;
	if (r->n_dev_pins < 4) r->n_dev_pins = 4;
	if (r->n_dev_pins > MIX_PINS) r->n_dev_pins = MIX_PINS;
	if (batch_capacity() > 0) {
		const int key[4] = {words, r->conference, r->n_dev_pins, 0};
		r->batch = batch_join(BK_MIXER, f->ticker, key, r->n_dev_pins * words, r->conference ? r->n_dev_pins * words : words, 1, r, NULL,
		                      &r->room);
	}
	if (r->batch) { /* this mixer is one room of the group's bank; its arenas are slices of the group's pinned arenas */
		r->bank = (msb200_mixer *)r->batch->bank;
		r->in = r->batch->in[0] + (size_t)r->room * r->batch->unit_in;
		r->out = r->batch->out + (size_t)r->room * r->batch->unit_out;
		r->present = r->batch->present + (size_t)r->room * r->n_dev_pins;
		GRP_LOCK(r->batch);
		mixroom_push_controls(r);
		GRP_UNLOCK(r->batch);
		return;
	}
	r->in = (int16_t *)ms_malloc0(sizeof(int16_t) * (size_t)r->n_dev_pins * (size_t)words);
	r->out = (int16_t *)ms_malloc0(sizeof(int16_t) * (size_t)r->n_dev_pins * (size_t)words);
	r->present = (uint8_t *)ms_malloc0((size_t)r->n_dev_pins);
	DSP_LOCK();
	if (dsp_ctx()) {
		DSP_CHECK(msb200_mixer_create(g_ctx, 1, r->n_dev_pins, words, r->conference, &r->bank), "mixer_create");
		mixroom_push_controls(r);
	}
	DSP_UNLOCK();
}
static void mixroom_detach(MSFilter *f) {
	MixRoom *r = (MixRoom *)f->data;
	if (r->batch) {
		batch_leave(r->batch, r->room);
		r->batch = NULL;
	} else {
		DSP_LOCK();
		msb200_mixer_destroy(r->bank);
		DSP_UNLOCK();
		ms_free(r->in);
		ms_free(r->out);
		ms_free(r->present);
	}
	r->bank = NULL;
	r->in = r->out = NULL;
	r->present = NULL;
}
static mblk_t *mix_block(const int16_t *samples, int words) {
	mblk_t *m = allocb((size_t)words * 2, 0);
	memcpy(m->b_wptr, samples, (size_t)words * 2);
	m->b_wptr += words * 2;
	return m;
}
/* one tick of results (r->out) to the listeners: everybody shares one block in plain mode (:321-334), every listener
 * gets the row computed for its own pin in conference mode (:336-342) */
static void mixroom_hand_out(MSFilter *f, MixRoom *r, int words) {
	mblk_t *shared = NULL;
	int k;
	for (k = 0; k < r->n_dev_pins; ++k) {
		if (f->outputs[k] == NULL || !r->pin[k].listens) continue;
		if (r->conference) ms_queue_put(f->outputs[k], mix_block(r->out + (size_t)k * words, words));
		else ms_queue_put(f->outputs[k], shared = shared ? dupb(shared) : mix_block(r->out, words));
	}
}
/* Who is talking? A pin counts while packets arrive and for MIX_SOLO_AFTER_MS after its last one (a pin that never sent
 * anything starts its silence clock at the first tick). Returns the number of talkers; *which = the highest of them. */
static int mixroom_talkers(MSFilter *f, MixRoom *r, int *which) {
	const uint64_t now = f->ticker->time;
	int k, n = 0;
	for (k = 0; k < MIX_PINS; ++k) {
		MixPin *p = &r->pin[k];
		bool_t talking;
		if (f->inputs[k] == NULL) continue;
		if (!ms_queue_empty(f->inputs[k])) {
			p->heard_at = now;
			talking = TRUE;
		} else if (p->heard_at == MIX_NEVER) {
			p->heard_at = now;
			talking = FALSE;
		} else {
			talking = now - p->heard_at < MIX_SOLO_AFTER_MS;
		}
		if (talking) {
			*which = k;
			++n;
		}
	}
	return n;
}
/* the single talker's packets go out as they came, to every listener but (in a conference) the talker itself (:219-239).
 * Packet-major: each packet is handed to the last listener as is and duplicated for the others. */
static void mixroom_forward_solo(MSFilter *f, MixRoom *r, int talker) {
	MSQueue *to[MIX_PINS];
	mblk_t *m;
	int k, n = 0;
	for (k = 0; k < MIX_PINS; ++k)
		if (f->outputs[k] && r->pin[k].listens && (k != talker || !r->conference)) to[n++] = f->outputs[k];
	while ((m = ms_queue_get(f->inputs[talker])) != NULL) {
		if (r->one_listener && n == 1) { /* a lone listener receives the packets themselves */
			ms_queue_put(to[0], m);
			continue;
		}
		for (k = 0; k < n; ++k) ms_queue_put(to[k], dupmsg(m));
		freemsg(m);
	}
}
/* drift control of one pin's FIFO: if over MIX_TRIM_EVERY_MS it never held less than trim_above bytes, the excess over
 * half of that is dropped (:92-111). Returns the bytes dropped. */
static int mixpin_trim(MixPin *p, uint64_t now, int trim_above) {
	int fill, dropped = 0;
	if (p->trimmed_at == MIX_NEVER) {
		p->trimmed_at = now;
		p->low_water = -1;
		return 0;
	}
	fill = (int)ms_bufferizer_get_avail(&p->fifo);
	if (p->low_water < 0 || fill < p->low_water) p->low_water = fill;
	if (now - p->trimmed_at < MIX_TRIM_EVERY_MS) return 0;
	if (p->low_water >= trim_above) {
		dropped = p->low_water - trim_above / 2;
		ms_bufferizer_skip_bytes(&p->fifo, dropped);
	}
	p->trimmed_at = now;
	p->low_water = -1;
	return dropped;
}
static void mixroom_tick(MSFilter *f) {
	MixRoom *r = (MixRoom *)f->data;
	const int words = r->tick_bytes / 2;
	int k, talkers, talker = -1;
	ms_filter_lock(f);
	if (r->batch) { /* the group's previous tick was computed by the first mixer called in this tick: hand out this room's share */
		batch_tick(r->batch, f->ticker->ticks);
		if (r->batch->ready[r->room]) {
			r->batch->ready[r->room] = 0;
			mixroom_hand_out(f, r, words);
		}
	}
	talkers = mixroom_talkers(f, r, &talker);
	if (talkers <= 1) { /* nothing to mix: silence all round, or one talker whose packets are passed on untouched */
		if (talkers == 1) {
			if (!r->solo) ms_message("MSAudioMixer(B200) %p: one talker left, forwarding its packets", f);
			r->solo = TRUE;
			mixroom_forward_solo(f, r, talker);
		}
		ms_filter_unlock(f);
		return;
	}
	if (r->solo) ms_message("MSAudioMixer(B200) %p: several talkers again, mixing", f);
	r->solo = FALSE;
	/* drain every pin's queue into its FIFO and read one tick straight into the pin's row of the device arena */
	memset(r->present, 0, (size_t)r->n_dev_pins);
	for (k = 0; k < r->n_dev_pins; ++k) {
		MixPin *p = &r->pin[k];
		int dropped;
		if (f->inputs[k] == NULL) continue;
		ms_bufferizer_put_from_queue(&p->fifo, f->inputs[k]);
		if (ms_bufferizer_read(&p->fifo, (uint8_t *)(r->in + (size_t)k * words), (size_t)r->tick_bytes) != 0) r->present[k] = 1;
		if ((dropped = mixpin_trim(p, f->ticker->time, r->trim_above)) > 0)
			ms_warning("MSAudioMixer(B200): pin %d runs ahead, %d ms dropped", k, (dropped * 1000) / (2 * r->channels * r->hz));
	}
	/* a block goes out every tick while several pins count as talkers, even when none of them had a whole tick of samples
	 * (silence then): the reference is built with ALWAYS_STREAMOUT (audiomixer.c:30, :315-317) */
	if (r->batch) { /* staged: the group's launch at the start of the next tick mixes every room at once */
		r->batch->staged[r->room] = 1;
		ms_filter_unlock(f);
		return;
	}
	DSP_LOCK();
	if (r->bank) DSP_CHECK(msb200_mixer_process(r->bank, r->in, r->present, r->out), "mixer_process");
	DSP_UNLOCK();
	if (r->bank) mixroom_hand_out(f, r, words);
	ms_filter_unlock(f);
}
/* ---- methods (audiomixer.c:348-431) */
static int mixm_set_hz(MSFilter *f, void *arg) {
	((MixRoom *)f->data)->hz = *(int *)arg;
	return 0;
}
static int mixm_get_hz(MSFilter *f, void *arg) {
	*(int *)arg = ((MixRoom *)f->data)->hz;
	return 0;
}
static int mixm_set_channels(MSFilter *f, void *arg) {
	((MixRoom *)f->data)->channels = *(int *)arg;
	return 0;
}
static int mixm_get_channels(MSFilter *f, void *arg) {
	*(int *)arg = ((MixRoom *)f->data)->channels;
	return 0;
}
static MixPin *mixm_pin(MSFilter *f, const MSAudioMixerCtl *ctl, const char *what) {
	if (ctl->pin >= 0 && ctl->pin < MIX_PINS) return &((MixRoom *)f->data)->pin[ctl->pin];
	ms_warning("MSAudioMixer(B200): %s: no pin %i", what, ctl->pin);
	return NULL;
}
static int mixm_set_gain(MSFilter *f, void *arg) {
	MixRoom *r = (MixRoom *)f->data;
	const MSAudioMixerCtl *ctl = (const MSAudioMixerCtl *)arg;
	MixPin *p = mixm_pin(f, ctl, "input gain");
	if (!p) return -1;
	p->gain = ctl->param.gain;
	if (r->bank && ctl->pin < r->n_dev_pins) {
		MIX_LOCK(r);
		msb200_mixer_set_input_gain(r->bank, r->room, ctl->pin, p->gain);
		MIX_UNLOCK(r);
	}
	return 0;
}
static int mixm_set_contributes(MSFilter *f, void *arg) {
	MixRoom *r = (MixRoom *)f->data;
	const MSAudioMixerCtl *ctl = (const MSAudioMixerCtl *)arg;
	MixPin *p = mixm_pin(f, ctl, "active flag");
	if (!p) return -1;
	p->contributes = (bool_t)ctl->param.active;
	if (r->bank && ctl->pin < r->n_dev_pins) {
		MIX_LOCK(r);
		msb200_mixer_set_active(r->bank, r->room, ctl->pin, p->contributes);
		MIX_UNLOCK(r);
	}
	return 0;
}
static int mixm_set_listens(MSFilter *f, void *arg) {
	MixRoom *r = (MixRoom *)f->data;
	const MSAudioMixerCtl *ctl = (const MSAudioMixerCtl *)arg;
	MixPin *p = mixm_pin(f, ctl, "output switch");
	if (!p) return -1;
	ms_filter_lock(f);
	p->listens = (bool_t)ctl->param.enabled;
	r->one_listener = mixroom_listeners(f, r) == 1;
	ms_filter_unlock(f);
	return 0;
}
static int mixm_set_conference(MSFilter *f, void *arg) {
	((MixRoom *)f->data)->conference = *(int *)arg;
	return 0;
}
static int mixm_set_master(MSFilter *f, void *arg) {
	((MixRoom *)f->data)->master_pin = *(int *)arg;
	return 0;
}
static MSFilterMethod mixroom_methods[] = {{MS_FILTER_SET_NCHANNELS, mixm_set_channels},
                                           {MS_FILTER_GET_NCHANNELS, mixm_get_channels},
                                           {MS_FILTER_SET_SAMPLE_RATE, mixm_set_hz},
                                           {MS_FILTER_GET_SAMPLE_RATE, mixm_get_hz},
                                           {MS_AUDIO_MIXER_SET_INPUT_GAIN, mixm_set_gain},
                                           {MS_AUDIO_MIXER_SET_ACTIVE, mixm_set_contributes},
                                           {MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE, mixm_set_conference},
                                           {MS_AUDIO_MIXER_SET_MASTER_CHANNEL, mixm_set_master},
                                           {MS_AUDIO_MIXER_ENABLE_OUTPUT, mixm_set_listens},
                                           {0, NULL}};
PROF_WRAP(PF_MIXER, mixroom_tick)
static MSFilterDesc b200_audio_mixer_desc = {.id = MS_AUDIO_MIXER_ID,
                                             .name = "MSAudioMixer",
                                             .text = "B200: mixes 16 bit sample audio streams (libmsb200dsp)",
                                             .category = MS_FILTER_OTHER,
                                             .ninputs = MIX_PINS,
                                             .noutputs = MIX_PINS,
                                             .init = mixroom_new,
                                             .preprocess = mixroom_attach,
                                             .process = mixroom_tick_timed,
                                             .postprocess = mixroom_detach,
                                             .uninit = mixroom_free,
                                             .methods = mixroom_methods,
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
	float gain_db; /* gain_dirty == 2: the pending gain came in dB */
	bool_t ng_enable_dirty, ng_floor_dirty, ng_floor_set; /* the noise gate's two gain-resetting setters: replayed only on their own account */
	MSBufferizer *buffer;
	float ea_thres, ea_speed, ea_force, ea_transmit;
	int ea_sustain;
	Batch *batch;     /* lockstep batch group (MSB200_BATCH), light path only; joined at the first block */
	int slot;
	mblk_t *held[8];  /* the blocks staged in the current tick: processed in the arena, copied back and forwarded next tick */
	int n_held;
	MSQueue pend;     /* processed blocks waiting for this filter's next process() */
	bool_t batch_off;
	/* echo-limiter links. `watchers` = filters that named THIS one as their peer: while there are any, this filter keeps
	 * its state in its private bank (never in a batch group's slot, which the watcher could not read) and tells them
	 * whenever that bank goes away or is recreated, so that no watcher keeps a pointer into a freed bank. */
	struct VolState *next_vol; /* every B200 MSVolume of the process (g_mu) */
	MSFilter *self;
	int watchers;
} VolState;
static MSFilterDesc b200_volume_desc;
static VolState *g_vols = NULL; /* g_mu */
/* DSP lock held: `gone`'s bank is about to be destroyed or replaced: every filter watching it unlinks on the device side */
static void vol_peer_bank_changed(VolState *gone) {
	VolState *w;
	for (w = g_vols; w; w = w->next_vol)
		if (w->peer && w->peer == gone->self && w->peer_linked) {
			if (w->bank) msb200_volume_set_peer(w->bank, 0, NULL, 0);
			w->peer_linked = FALSE; /* re-linked at the next sync if the peer has a bank again */
		}
}
#define VOL_MAX_BLOCK 8192
#define VOL_BATCH_UNITS 8 /* blocks one stream may stage per tick (an upstream MSSpeexEC emits 1-2 frames per 10 ms) */

static void vol_mark_all_dirty(VolState *v) { /* a fresh bank or slot: everything the user configured is replayed once */
	v->dirty = TRUE;
	v->ng_enable_dirty = v->noise_gate;
	v->ng_floor_dirty = v->ng_floor_set;
}
static void vol_sync_config_to(VolState *v, msb200_volume *bank, int st) { /* DSP lock held */
	if (!bank) return;
	if (v->peer && !v->peer_linked && bank == v->bank && v->peer->desc == &b200_volume_desc && ((VolState *)v->peer->data)->bank) {
		/* the peer's smoothed energy is read from its PRIVATE bank (a watched filter stays out of the batch groups) */
		msb200_volume_set_peer(v->bank, 0, ((VolState *)v->peer->data)->bank, 0);
		v->peer_linked = TRUE;
	}
	if (v->gain_dirty) { /* MS_VOLUME_SET_GAIN resets the ramp (gain = target = static, msvolume.c:270-276): apply it once */
		if (v->gain_dirty == 2) msb200_volume_set_db_gain(bank, st, v->gain_db); /* ... _SET_DB_GAIN leaves the target alone (:262-268) */
		else msb200_volume_set_gain(bank, st, v->static_gain);
		v->gain_dirty = FALSE;
	}
	if (!v->dirty) return;
	/* the gate's two resetting setters (gain = target = floor gain, msvolume.c:352-378) are replayed only when THEY were
	 * called — or once for a fresh bank / slot (vol_mark_all_dirty) — never as a side effect of an unrelated setting */
	if (v->ng_floor_dirty) msb200_volume_set_noise_gate_floorgain(bank, st, v->ng_floorgain);
	if (v->ng_enable_dirty) msb200_volume_enable_noise_gate(bank, st, v->noise_gate ? 1 : 0);
	v->ng_floor_dirty = v->ng_enable_dirty = FALSE;
	msb200_volume_set_noise_gate_threshold(bank, st, v->ng_threshold);
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
	vol_mark_all_dirty(v);
	v->gain_dirty = FALSE; /* the bank starts at gain 1 like volume_init */
	v->self = f;
	DSP_LOCK();
	v->next_vol = g_vols;
	g_vols = v;
	DSP_UNLOCK();
	f->data = v;
}
static void vol_uninit(MSFilter *f) {
	VolState *v = (VolState *)f->data;
	VolState **pp;
	vol_leave_batch(v);
	ms_queue_flush(&v->pend);
	DSP_LOCK();
	vol_peer_bank_changed(v);
	for (pp = &g_vols; *pp && *pp != v; pp = &(*pp)->next_vol) {
	}
	if (*pp) *pp = v->next_vol;
	{
		VolState *w;
		for (w = g_vols; w; w = w->next_vol)
			if (w->peer == f) w->peer = NULL; /* a watcher must not keep the pointer to a destroyed filter */
	}
	if (v->peer && v->peer->desc == &b200_volume_desc) ((VolState *)v->peer->data)->watchers--;
	msb200_volume_destroy(v->bank);
	DSP_UNLOCK();
	ms_bufferizer_destroy(v->buffer);
	ms_free(v);
}
static void vol_preprocess(MSFilter *f) {
	VolState *v = (VolState *)f->data;
	DSP_LOCK();
	if (dsp_ctx() && (!v->bank || v->bank_rate != v->rate)) {
		vol_peer_bank_changed(v);
		msb200_volume_destroy(v->bank);
		v->bank = NULL;
		DSP_CHECK(msb200_volume_create(g_ctx, 1, v->rate, VOL_MAX_BLOCK, &v->bank), "volume_create");
		v->bank_rate = v->rate;
		vol_mark_all_dirty(v);
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
	if (v->agc || v->peer != NULL) {
		/* AGC or an echo-limiter peer: the reference re-frames to 10 ms chunks and runs one at a time (msvolume.c:480-502).
		 * Here every whole chunk the FIFO holds goes to the device in ONE call (the kernel walks the chunks in order; the
		 * peer's energy does not move while this filter runs), then leaves as one block per chunk. */
		const int chunk = (int)(0.01 * (float)v->rate);
		const size_t chunk_bytes = (size_t)chunk * 2;
		size_t k, n_chunks;
		int16_t *run;
		bool_t ok = FALSE;
		if (v->batch) vol_leave_batch(v);
		ms_bufferizer_put_from_queue(v->buffer, f->inputs[0]);
		n_chunks = chunk_bytes ? ms_bufferizer_get_avail(v->buffer) / chunk_bytes : 0;
		if (n_chunks == 0) return;
		run = (int16_t *)ms_malloc(n_chunks * chunk_bytes);
		ms_bufferizer_read(v->buffer, (uint8_t *)run, n_chunks * chunk_bytes);
		DSP_LOCK();
		vol_sync_config(v);
		if (v->bank) {
			DSP_CHECK(msb200_volume_process_blocks(v->bank, run, chunk, (int)(n_chunks * chunk), (int)n_chunks, NULL), "volume_process");
			ok = TRUE;
		}
		DSP_UNLOCK();
		for (k = 0; ok && k < n_chunks; ++k) {
			m = allocb(chunk_bytes, 0);
			memcpy(m->b_wptr, run + k * chunk, chunk_bytes);
			m->b_wptr += chunk_bytes;
			pin0_send(f, m);
		}
		ms_free(run);
		return;
	}
	if (v->batch && v->batch->key[0] != v->rate) vol_leave_batch(v);
	if (v->batch) { /* the blocks staged in the previous tick have been processed in the arena by the group's launch */
		batch_tick(v->batch, f->ticker->ticks);
		if (v->n_held && v->batch->staged[v->slot] == 0) vol_collect(v, v->batch);
	}
	while ((m = ms_queue_get(&v->pend)) != NULL)
		pin0_send(f, m);
	while ((m = pin0_next(f)) != NULL) {
		int n = (int)((m->b_wptr - m->b_rptr) / 2);
		if (v->batch && v->watchers > 0) { /* somebody's echo-limiter peer: the state must live in the private bank */
			vol_leave_batch(v);
			vol_mark_all_dirty(v);
			v->gain_dirty = v->static_gain != 1.0f;
		}
		if (!v->batch && !v->batch_off && v->watchers == 0 && batch_capacity() > 0 && n > 0 && n <= VOL_MAX_BLOCK) {
			const int key[4] = {v->rate, n, 0, 0};
			v->batch = batch_join(BK_VOLUME, f->ticker, key, n, n, VOL_BATCH_UNITS, v, vol_collect, &v->slot);
			if (v->batch) {
				GRP_LOCK(v->batch);
				msb200_ctx_make_current(v->batch->ctx);
				msb200_volume_reset_stream((msb200_volume *)v->batch->bank, v->slot);
				vol_mark_all_dirty(v);
				v->gain_dirty = v->static_gain != 1.0f;
				GRP_UNLOCK(v->batch);
				batch_tick(v->batch, f->ticker->ticks); /* (nothing of ours is staged yet: see batch_tick) */
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
			vol_mark_all_dirty(v);
			v->gain_dirty = v->static_gain != 1.0f;
		}
		if (v->bank && n > 0 && n <= VOL_MAX_BLOCK) {
			DSP_LOCK();
			vol_sync_config(v);
			DSP_CHECK(msb200_volume_process(v->bank, (int16_t *)m->b_rptr, n), "volume_process");
			DSP_UNLOCK();
			pin0_send(f, m);
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
	v->gain_db = *(float *)arg;
	v->static_gain = (float)pow(10, v->gain_db / 10);
	v->gain_dirty = 2;
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
	MSFilter *peer = (MSFilter *)arg;
	DSP_LOCK();
	if (v->peer && v->peer->desc == &b200_volume_desc) ((VolState *)v->peer->data)->watchers--;
	if (v->peer_linked && v->bank) msb200_volume_set_peer(v->bank, 0, NULL, 0);
	v->peer = peer;
	v->peer_linked = FALSE;
	if (peer && peer->desc == &b200_volume_desc) ((VolState *)peer->data)->watchers++;
	DSP_UNLOCK();
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
	v->ng_enable_dirty = TRUE;
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
	v->ng_floor_dirty = v->ng_floor_set = TRUE;
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
PROF_WRAP(PF_VOLUME, vol_process)
static MSFilterDesc b200_volume_desc = {.id = MS_VOLUME_ID,
                                        .name = "MSVolume",
                                        .text = "B200: controls and measures sound volume (libmsb200dsp)",
                                        .category = MS_FILTER_OTHER,
                                        .ninputs = 1,
                                        .noutputs = 1,
                                        .init = vol_init,
                                        .preprocess = vol_preprocess,
                                        .process = vol_process_timed,
                                        .postprocess = vol_postprocess,
                                        .uninit = vol_uninit,
                                        .methods = vol_methods};

/* ================================================================================================ MSChannelAdapter
 * /root/reference/src/audiofilters/chanadapt.c: per block mono -> stereo / stereo -> mono (:99-132), or — with both
 * input pins connected — two mono streams interleaved into one stereo stream, one ticker interval at a time (:68-97).
 * The interleaving / duplication / left-channel pick runs in chanadapt_kernel. */
typedef struct ChanSide { /* one mono input of the two-pin mode */
	MSFlowControlledBufferizer fifo;
	uint8_t *tick; /* one ticker interval of samples */
} ChanSide;
typedef struct ChanAdapter {
	int in_ch, out_ch, hz;
	size_t tick_bytes;
	bool_t sides_ready; /* the two-pin state exists (the reference sets it up when in = 2, out = 1 channels: chanadapt.c:55) */
	ChanSide side[2];
} ChanAdapter;
static void chan_new(MSFilter *f) {
	ChanAdapter *a = ms_new0(ChanAdapter, 1);
	a->in_ch = a->out_ch = 1;
	a->hz = 8000;
	f->data = a;
}
static void chan_free(MSFilter *f) {
	ms_free(f->data);
}
static void chan_attach(MSFilter *f) {
	ChanAdapter *a = (ChanAdapter *)f->data;
	int k;
	DSP_LOCK();
	dsp_ctx();
	DSP_UNLOCK();
	a->sides_ready = a->in_ch == 2 && a->out_ch == 1;
	if (!a->sides_ready) return;
	a->tick_bytes = (size_t)((f->ticker->interval * a->hz) / 1000) * 2;
	for (k = 0; k < 2; ++k) { /* a side never buffers more than two intervals: the excess is dropped at once (:59-64) */
		a->side[k].tick = ms_new(uint8_t, a->tick_bytes);
		ms_flow_controlled_bufferizer_init(&a->side[k].fifo, f, a->hz, 1);
		ms_flow_controlled_bufferizer_set_drop_method(&a->side[k].fifo, MSFlowControlledBufferizerImmediateDrop);
		ms_flow_controlled_bufferizer_set_max_size_ms(&a->side[k].fifo, f->ticker->interval * 2);
	}
}
static void chan_detach(MSFilter *f) {
	ChanAdapter *a = (ChanAdapter *)f->data;
	int k;
	if (!a->sides_ready) return;
	for (k = 0; k < 2; ++k) {
		ms_flow_controlled_bufferizer_uninit(&a->side[k].fifo);
		ms_free(a->side[k].tick);
		a->side[k].tick = NULL;
	}
	a->sides_ready = FALSE;
}
/* device call + hand-over shared by all three modes: `frames` frames from a (and b) into a fresh block of out_bytes */
static void chan_convert(MSFilter *f, int mode, int frames, const int16_t *a, const int16_t *b, size_t out_bytes) {
	mblk_t *om = allocb(out_bytes, 0);
	bool_t ok;
	DSP_LOCK();
	ok = g_ctx != NULL;
	if (ok && frames > 0) DSP_CHECK(msb200_chanadapt_process(g_ctx, mode, 1, frames, a, b, (int16_t *)om->b_wptr), "chanadapt");
	DSP_UNLOCK();
	om->b_wptr += out_bytes;
	if (ok) pin0_send(f, om);
	else freemsg(om); /* no GPU: never forward unprocessed audio */
}
static void chan_tick(MSFilter *f) {
	ChanAdapter *a = (ChanAdapter *)f->data;
	mblk_t *im;
	if (f->inputs[0] && f->inputs[1]) { /* two mono pins -> one stereo stream; a side without a full interval gives silence */
		bool_t full[2];
		int k;
		for (k = 0; k < 2; ++k) {
			ms_flow_controlled_bufferizer_put_from_queue(&a->side[k].fifo, f->inputs[k]);
			full[k] = ms_flow_controlled_bufferizer_get_avail(&a->side[k].fifo) >= a->tick_bytes;
		}
		if (!full[0] && !full[1]) return;
		for (k = 0; k < 2; ++k) ms_flow_controlled_bufferizer_read(&a->side[k].fifo, a->side[k].tick, a->tick_bytes);
		chan_convert(f, MSB200_CHAN_2MONO_TO_STEREO, (int)(a->tick_bytes / 2), full[0] ? (const int16_t *)a->side[0].tick : NULL,
		             full[1] ? (const int16_t *)a->side[1].tick : NULL, a->tick_bytes * 2);
		return;
	}
	while ((im = pin0_next(f)) != NULL) {
		const size_t in_bytes = msgdsize(im);
		if (a->in_ch == a->out_ch) {
			pin0_send(f, im);
			continue;
		}
		if (a->out_ch == 2) chan_convert(f, MSB200_CHAN_MONO_TO_STEREO, (int)(in_bytes / 2), (const int16_t *)im->b_rptr, NULL, in_bytes * 2);
		else chan_convert(f, MSB200_CHAN_STEREO_TO_MONO, (int)(in_bytes / 4), (const int16_t *)im->b_rptr, NULL, in_bytes / 2);
		freemsg(im);
	}
}
static int chanm_set_hz(MSFilter *f, void *arg) {
	((ChanAdapter *)f->data)->hz = *(int *)arg;
	return 0;
}
static int chanm_get_hz(MSFilter *f, void *arg) {
	*(int *)arg = ((ChanAdapter *)f->data)->hz;
	return 0;
}
static int chanm_set_in(MSFilter *f, void *arg) {
	((ChanAdapter *)f->data)->in_ch = *(int *)arg;
	return 0;
}
static int chanm_get_in(MSFilter *f, void *arg) {
	*(int *)arg = ((ChanAdapter *)f->data)->in_ch;
	return 0;
}
static int chanm_set_out(MSFilter *f, void *arg) {
	((ChanAdapter *)f->data)->out_ch = *(int *)arg;
	return 0;
}
static int chanm_get_out(MSFilter *f, void *arg) {
	*(int *)arg = ((ChanAdapter *)f->data)->out_ch;
	return 0;
}
static MSFilterMethod chan_methods[] = {{MS_FILTER_SET_SAMPLE_RATE, chanm_set_hz},
                                        {MS_FILTER_GET_SAMPLE_RATE, chanm_get_hz},
                                        {MS_FILTER_SET_NCHANNELS, chanm_set_in},
                                        {MS_FILTER_GET_NCHANNELS, chanm_get_in},
                                        {MS_CHANNEL_ADAPTER_SET_OUTPUT_NCHANNELS, chanm_set_out},
                                        {MS_CHANNEL_ADAPTER_GET_OUTPUT_NCHANNELS, chanm_get_out},
                                        {0, NULL}};
PROF_WRAP(PF_CHAN, chan_tick)
static MSFilterDesc b200_channel_adapter_desc = {.id = MS_CHANNEL_ADAPTER_ID,
                                                 .name = "MSChannelAdapter",
                                                 .text = "B200: mono/stereo channel adaptation (libmsb200dsp)",
                                                 .category = MS_FILTER_OTHER,
                                                 .ninputs = 2,
                                                 .noutputs = 1,
                                                 .init = chan_new,
                                                 .preprocess = chan_attach,
                                                 .process = chan_tick_timed,
                                                 .postprocess = chan_detach,
                                                 .uninit = chan_free,
                                                 .methods = chan_methods,
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
	while ((m = pin0_next(f)) != NULL) {
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
		pin0_send(f, m);
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
PROF_WRAP(PF_EQ, eq_process)
static MSFilterDesc b200_equalizer_desc = {.id = MS_EQUALIZER_ID,
                                           .name = "MSEqualizer",
                                           .text = "B200: parametric sound equalizer (libmsb200dsp)",
                                           .category = MS_FILTER_OTHER,
                                           .ninputs = 1,
                                           .noutputs = 1,
                                           .init = eq_init,
                                           .preprocess = eq_preprocess,
                                           .process = eq_process_timed,
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
		while ((im = pin0_next(f)) != NULL) {
			if (s->out_nchannels == s->in_nchannels) {
				pin0_send(f, im);
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
		pin0_send(f, im);
	while ((im = pin0_next(f)) != NULL) {
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
				batch_tick(s->batch, f->ticker->ticks);
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
			pin0_send(f, om);
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
PROF_WRAP(PF_RESAMPLE, rs_process)
static MSFilterDesc b200_resample_desc = {.id = MS_RESAMPLE_ID,
                                          .name = "MSResample",
                                          .text = "B200: audio resampler (libmsb200dsp)",
                                          .category = MS_FILTER_OTHER,
                                          .ninputs = 1,
                                          .noutputs = 1,
                                          .init = rs_init,
                                          .preprocess = rs_preprocess,
                                          .process = rs_process_timed,
                                          .postprocess = rs_postprocess,
                                          .uninit = rs_uninit,
                                          .methods = rs_methods};

/* ================================================================================================ MSSpeexEC
 * host logic restated from /root/reference/src/audiofilters/speexec.c:171-216 (configuration), :223-305 (process:
 * reference/echo bufferizers, silence injection on underrun), :308-391 (methods) */
typedef struct EchoCanceller {
	msb200_aec *bank;             /* synchronous mode: a 1-stream bank of its own */
	MSBufferizer far_for_filter;  /* far-end audio as the adaptive filter sees it: behind the playback by `delay_ms` */
	MSFlowControlledBufferizer far_for_playback; /* the same audio on its way to the loudspeaker pin */
	MSBufferizer near;            /* microphone */
	int frame, frame_at_8k, hz, delay_ms, tail_ms, delay_samples;
	char *saved_state;
	bool_t mic_seen, bypass, far_starved;
	Batch *batch; /* lockstep batch group (MSB200_BATCH): frames are staged here and cancelled one tick later */
	int slot;
	MSQueue cleaned; /* cancelled frames waiting for this filter's next process() */
} EchoCanceller;
#define EC_BATCH_MAX_FRAMES 4 /* frames one stream may stage per tick (10 ms at 48 kHz = 1.875 frames of 256) */

static void ec_tune_playback_fifo(EchoCanceller *e) { /* speexec.c:182-186 */
	ms_flow_controlled_bufferizer_set_samplerate(&e->far_for_playback, e->hz);
	ms_flow_controlled_bufferizer_set_max_size_ms(&e->far_for_playback, e->delay_ms);
	ms_flow_controlled_bufferizer_set_granularity_ms(&e->far_for_playback, (e->frame * 1000) / e->hz);
}
static void ec_new(MSFilter *f) {
	EchoCanceller *e = ms_new0(EchoCanceller, 1);
	e->hz = 8000; /* defaults of speex_ec_init, speexec.c:72-91 */
	e->tail_ms = 250;
	e->frame = e->frame_at_8k = 64;
	ms_bufferizer_init(&e->far_for_filter);
	ms_bufferizer_init(&e->near);
	ms_flow_controlled_bufferizer_init(&e->far_for_playback, f, e->hz, 1);
	ms_queue_init(&e->cleaned);
	f->data = e;
}
static void ec_free(MSFilter *f) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	ms_queue_flush(&e->cleaned);
	if (e->saved_state) ms_free(e->saved_state);
	ms_bufferizer_uninit(&e->far_for_filter);
	ms_bufferizer_uninit(&e->near);
	ms_flow_controlled_bufferizer_uninit(&e->far_for_playback);
	ms_free(e);
}
static mblk_t *ec_frame_block(const EchoCanceller *e) { /* an empty block with room for one frame */
	return allocb((size_t)e->frame * 2, 0);
}
/* batch mode: this slot's cancelled frames move from the group's arena into blocks of their own */
static void ec_collect(void *owner, Batch *b) {
	EchoCanceller *e = (EchoCanceller *)owner;
	int u;
	for (u = 0; u < b->ready[e->slot]; ++u) {
		mblk_t *m = ec_frame_block(e);
		memcpy(m->b_wptr, b->out + ((size_t)e->slot * b->max_units + u) * b->unit_out, (size_t)e->frame * 2);
		m->b_wptr += e->frame * 2;
		ms_queue_put(&e->cleaned, m);
	}
	b->ready[e->slot] = 0;
}
static void ec_attach(MSFilter *f) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	mblk_t *lead;
	e->mic_seen = FALSE;
	e->frame = msb200_aec_frame_size_for_rate(e->hz, e->frame_at_8k);
	e->delay_samples = e->delay_ms * e->hz / 1000;
	ms_message("MSSpeexEC(B200): frame %i, filter %i taps, filter path %i samples behind the playback", e->frame,
	           (e->tail_ms * e->hz) / 1000, e->delay_samples);
	if (batch_capacity() > 0) {
		const int key[4] = {e->hz, e->tail_ms, e->frame_at_8k, 0};
		e->batch = batch_join(BK_EC, f->ticker, key, e->frame, e->frame, EC_BATCH_MAX_FRAMES, e, ec_collect, &e->slot);
		if (e->batch) {
			GRP_LOCK(e->batch);
			msb200_ctx_make_current(e->batch->ctx);
			msb200_aec_reset((msb200_aec *)e->batch->bank, e->slot);
			GRP_UNLOCK(e->batch);
		}
	}
	if (!e->batch) {
		DSP_LOCK();
		if (dsp_ctx()) DSP_CHECK(msb200_aec_create(g_ctx, 1, e->hz, e->tail_ms, e->frame_at_8k, &e->bank), "aec_create");
		DSP_UNLOCK();
	}
	/* the filter path starts `delay_samples` of silence behind the playback path (speexec.c:205-213) */
	lead = allocb((size_t)e->delay_samples * 2, 0);
	lead->b_wptr += e->delay_samples * 2;
	ms_bufferizer_put(&e->far_for_filter, lead);
	ec_tune_playback_fifo(e);
}
static void ec_detach(MSFilter *f) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	ms_bufferizer_flush(&e->far_for_filter);
	ms_bufferizer_flush(&e->near);
	ms_flow_controlled_bufferizer_flush(&e->far_for_playback);
	if (e->batch) batch_leave(e->batch, e->slot);
	e->batch = NULL;
	DSP_LOCK();
	msb200_aec_destroy(e->bank);
	DSP_UNLOCK();
	e->bank = NULL;
}
/* far-end packets: once the microphone runs, every packet feeds both paths; before that there is nothing to line them up
 * with and they are dropped (speexec.c:239-250) */
static void ec_take_far_end(MSFilter *f, EchoCanceller *e) {
	mblk_t *m;
	if (f->inputs[0] == NULL) return;
	if (!e->mic_seen) {
		if (!ms_queue_empty(f->inputs[0])) ms_warning("MSSpeexEC(B200): far-end audio before any microphone audio, dropped");
		ms_queue_flush(f->inputs[0]);
		return;
	}
	while ((m = pin0_next(f)) != NULL) {
		ms_bufferizer_put(&e->far_for_filter, dupmsg(m));
		ms_flow_controlled_bufferizer_put(&e->far_for_playback, m);
	}
}
/* One frame for the loudspeaker pin, and the guarantee that the filter path holds this frame too. When the far end
 * starves (less than delay + one frame buffered for the filter) a frame of silence is played AND appended to the filter
 * path, so that both stay aligned (speexec.c:261-285). */
static void ec_play_one_frame(MSFilter *f, EchoCanceller *e) {
	const size_t bytes = (size_t)e->frame * 2;
	mblk_t *spk = ec_frame_block(e);
	const bool_t starving = ms_bufferizer_get_avail(&e->far_for_filter) < (size_t)e->delay_samples * 2 + bytes;
	if (starving) {
		memset(spk->b_wptr, 0, bytes);
		spk->b_wptr += bytes;
		ms_bufferizer_put(&e->far_for_filter, dupmsg(spk));
	} else {
		if (ms_flow_controlled_bufferizer_read(&e->far_for_playback, spk->b_wptr, bytes) == 0)
			ms_fatal("MSSpeexEC(B200): playback path shorter than the filter path");
		spk->b_wptr += bytes;
	}
	if (starving != e->far_starved) {
		if (starving) ms_warning("MSSpeexEC(B200): far end starving, playing silence");
		else ms_message("MSSpeexEC(B200): far end is back");
		e->far_starved = starving;
	}
	pin0_send(f, spk);
}
/* the (microphone, reference) pair of one frame goes to the canceller: at once in synchronous mode, into the group's arena
 * in batch mode (cancelled by the group's launch at the start of the next tick) */
static void ec_cancel_frame(MSFilter *f, EchoCanceller *e, const int16_t *mic, const int16_t *ref) {
	const size_t bytes = (size_t)e->frame * 2;
	if (e->batch) {
		Batch *b = e->batch;
		const int u = b->staged[e->slot];
		if (u >= b->max_units) {
			ms_warning("MSSpeexEC(B200): more than %d frames in one tick, frame dropped", b->max_units);
			return;
		}
		memcpy(b->in[0] + ((size_t)e->slot * b->max_units + u) * b->unit_in, mic, bytes);
		memcpy(b->in[1] + ((size_t)e->slot * b->max_units + u) * b->unit_in, ref, bytes);
		b->staged[e->slot] = u + 1;
		return;
	}
	if (e->bank) {
		mblk_t *clean = ec_frame_block(e);
		DSP_LOCK();
		DSP_CHECK(msb200_aec_process(e->bank, mic, ref, (int16_t *)clean->b_wptr, 1), "aec_process");
		DSP_UNLOCK();
		clean->b_wptr += bytes;
		ms_queue_put(f->outputs[1], clean);
	}
}
static void ec_tick(MSFilter *f) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	const size_t bytes = (size_t)e->frame * 2;
	int16_t *mic, *ref;
	mblk_t *m;
	if (e->batch) { /* frames staged during the previous tick were cancelled by the group's launch: emit ours */
		batch_tick(e->batch, f->ticker->ticks);
		if (e->batch->staged[e->slot] == 0) ec_collect(e, e->batch);
	}
	while ((m = ms_queue_get(&e->cleaned)) != NULL) ms_queue_put(f->outputs[1], m);
	if (e->bypass) { /* both pins straight through (speexec.c:229-237) */
		int pin;
		for (pin = 0; pin < 2; ++pin)
			while ((m = ms_queue_get(f->inputs[pin])) != NULL) ms_queue_put(f->outputs[pin], m);
		return;
	}
	ec_take_far_end(f, e);
	ms_bufferizer_put_from_queue(&e->near, f->inputs[1]);
	mic = (int16_t *)alloca(bytes);
	ref = (int16_t *)alloca(bytes);
	while (ms_bufferizer_read(&e->near, (uint8_t *)mic, bytes) == bytes) {
		e->mic_seen = TRUE;
		ec_play_one_frame(f, e);
		if (ms_bufferizer_read(&e->far_for_filter, (uint8_t *)ref, bytes) == 0)
			ms_fatal("MSSpeexEC(B200): filter path empty after it was topped up");
		ec_cancel_frame(f, e, mic, ref);
	}
}
/* ---- methods (speexec.c:320-391) */
static int ecm_set_hz(MSFilter *f, void *arg) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	e->hz = *(int *)arg;
	ec_tune_playback_fifo(e);
	return 0;
}
static int ecm_get_hz(MSFilter *f, void *arg) {
	*(int *)arg = ((EchoCanceller *)f->data)->hz;
	return 0;
}
static int ecm_set_frame(MSFilter *f, void *arg) {
	((EchoCanceller *)f->data)->frame_at_8k = *(int *)arg;
	return 0;
}
static int ecm_set_delay(MSFilter *f, void *arg) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	e->delay_ms = *(int *)arg;
	ec_tune_playback_fifo(e);
	return 0;
}
static int ecm_get_delay(MSFilter *f, void *arg) {
	*(int *)arg = ((EchoCanceller *)f->data)->delay_ms;
	return 0;
}
static int ecm_set_tail(MSFilter *f, void *arg) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	e->tail_ms = *(int *)arg;
	ec_tune_playback_fifo(e);
	return 0;
}
static int ecm_set_bypass(MSFilter *f, void *arg) {
	((EchoCanceller *)f->data)->bypass = *(bool_t *)arg;
	return 0;
}
static int ecm_get_bypass(MSFilter *f, void *arg) {
	*(bool_t *)arg = ((EchoCanceller *)f->data)->bypass;
	return 0;
}
static int ecm_set_state(MSFilter *f, void *arg) {
	EchoCanceller *e = (EchoCanceller *)f->data;
	if (e->saved_state) ms_free(e->saved_state);
	e->saved_state = ms_strdup((const char *)arg);
	return 0;
}
static int ecm_get_state(MSFilter *f, void *arg) {
	*(char **)arg = ((EchoCanceller *)f->data)->saved_state;
	return 0;
}
static MSFilterMethod ec_methods[] = {{MS_FILTER_SET_SAMPLE_RATE, ecm_set_hz},
                                      {MS_FILTER_GET_SAMPLE_RATE, ecm_get_hz},
                                      {MS_ECHO_CANCELLER_SET_TAIL_LENGTH, ecm_set_tail},
                                      {MS_ECHO_CANCELLER_SET_DELAY, ecm_set_delay},
                                      {MS_ECHO_CANCELLER_SET_FRAMESIZE, ecm_set_frame},
                                      {MS_ECHO_CANCELLER_SET_BYPASS_MODE, ecm_set_bypass},
                                      {MS_ECHO_CANCELLER_GET_BYPASS_MODE, ecm_get_bypass},
                                      {MS_ECHO_CANCELLER_GET_STATE_STRING, ecm_get_state},
                                      {MS_ECHO_CANCELLER_SET_STATE_STRING, ecm_set_state},
                                      {MS_ECHO_CANCELLER_GET_DELAY, ecm_get_delay},
                                      {0, NULL}};
PROF_WRAP(PF_EC, ec_tick)
static MSFilterDesc b200_speex_ec_desc = {.id = MS_SPEEX_EC_ID,
                                          .name = "MSSpeexEC",
                                          .text = "B200: MDF echo canceller + preprocessor (libmsb200dsp)",
                                          .category = MS_FILTER_OTHER,
                                          .ninputs = 2,
                                          .noutputs = 2,
                                          .init = ec_new,
                                          .preprocess = ec_attach,
                                          .process = ec_tick_timed,
                                          .postprocess = ec_detach,
                                          .uninit = ec_free,
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
	if (s->ptime >= 10) frame_per_packet = s->ptime / 10 < 14 ? s->ptime / 10 : 14; /* packets of 10 .. 140 ms */
	size_of_pcm = (size_t)160 * frame_per_packet;       /* bytes: 80 samples per 10 ms at 8 kHz */
	while ((m = pin0_next(f)) != NULL)
		ms_bufferizer_put(s->bz, m);
	if (s->batch && s->batch->key[1] != (int)size_of_pcm / 2) g711_enc_leave_batch(s); /* ptime changed under us */
	if (s->batch) { /* the packets staged in the previous tick were encoded by the group's launch */
		batch_tick(s->batch, f->ticker->ticks);
		if (s->n_held && s->batch->staged[s->slot] == 0) g711_enc_collect(s, s->batch);
	}
	while ((m = ms_queue_get(&s->pend)) != NULL)
		pin0_send(f, m);
	avail = ms_bufferizer_get_avail(s->bz);
	npk = (int)(avail / size_of_pcm);
	if (npk == 0) return;
	if (!s->batch && !s->batch_off && batch_capacity() > 0) {
		const int key[4] = {s->law, (int)size_of_pcm / 2, 0, 0};
		s->batch = batch_join(BK_G711ENC, f->ticker, key, (int)size_of_pcm / 2, (int)size_of_pcm / 4, G711_BATCH_UNITS, s, g711_enc_collect,
		                      &s->slot);
		if (!s->batch) s->batch_off = TRUE;
		else batch_tick(s->batch, f->ticker->ticks);
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
/* one integer parameter of an SDP fmtp line; 0 when the key is absent */
static int g711_fmtp_int(const char *line, const char *key, int *out) {
	char text[32];
	if (!fmtp_get_value(line, key, text, sizeof(text))) return 0;
	*out = atoi(text);
	return 1;
}
static int g711_enc_add_fmtp(MSFilter *f, void *arg) { /* same meaning as enc_add_fmtp, alaw.c:96-110: maxptime caps ptime */
	G711EncState *s = (G711EncState *)f->data;
	int ms = 0;
	if (g711_fmtp_int((const char *)arg, "maxptime", &ms)) s->maxptime = ms < MS_DEFAULT_MAX_PTIME ? ms : MS_DEFAULT_MAX_PTIME;
	if (g711_fmtp_int((const char *)arg, "ptime", &ms)) {
		s->ptime = ms < s->maxptime ? ms : s->maxptime;
		ms_message("%s: ptime %d ms asked, %d ms used (maxptime %d)", f->desc->name, ms, s->ptime, s->maxptime);
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
		else batch_tick(s->batch, f->ticker->ticks);
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
		pin0_send(f, m);
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
			pin0_send(f, o);
		}
		off += len;
		freemsg(list[k]);
	}
	ms_free(code);
	ms_free(pcm);
}
static void alaw_dec_process(MSFilter *f) { g711_dec_process_law(f, MSB200_G711_ALAW); }
static void ulaw_dec_process(MSFilter *f) { g711_dec_process_law(f, MSB200_G711_ULAW); }
PROF_WRAP(PF_G711ENC, g711_enc_process)
static MSFilterDesc b200_alaw_enc_desc = {.id = MS_ALAW_ENC_ID, .name = "MSAlawEnc", .text = "B200: ITU-G.711 alaw encoder (libmsb200dsp)",
                                          .category = MS_FILTER_ENCODER, .enc_fmt = "pcma", .ninputs = 1, .noutputs = 1,
                                          .init = alaw_enc_init, .process = g711_enc_process_timed, .postprocess = g711_enc_postprocess, .uninit = g711_enc_uninit,
                                          .methods = g711_enc_methods};
static MSFilterDesc b200_ulaw_enc_desc = {.id = MS_ULAW_ENC_ID, .name = "MSUlawEnc", .text = "B200: ITU-G.711 ulaw encoder (libmsb200dsp)",
                                          .category = MS_FILTER_ENCODER, .enc_fmt = "pcmu", .ninputs = 1, .noutputs = 1,
                                          .init = ulaw_enc_init, .process = g711_enc_process_timed, .postprocess = g711_enc_postprocess, .uninit = g711_enc_uninit,
                                          .methods = g711_enc_methods};
PROF_WRAP(PF_G711DEC, alaw_dec_process)
static MSFilterDesc b200_alaw_dec_desc = {.id = MS_ALAW_DEC_ID, .name = "MSAlawDec", .text = "B200: ITU-G.711 alaw decoder (libmsb200dsp)",
                                          .category = MS_FILTER_DECODER, .enc_fmt = "pcma", .ninputs = 1, .noutputs = 1,
                                          .init = alaw_dec_init, .process = alaw_dec_process_timed, .postprocess = g711_dec_postprocess,
                                          .uninit = g711_dec_uninit, .methods = g711_dec_methods};
PROF_WRAP(PF_G711DEC, ulaw_dec_process)
static MSFilterDesc b200_ulaw_dec_desc = {.id = MS_ULAW_DEC_ID, .name = "MSUlawDec", .text = "B200: ITU-G.711 ulaw decoder (libmsb200dsp)",
                                          .category = MS_FILTER_DECODER, .enc_fmt = "pcmu", .ninputs = 1, .noutputs = 1,
                                          .init = ulaw_dec_init, .process = ulaw_dec_process_timed, .postprocess = g711_dec_postprocess,
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
	while ((m = pin0_next(f)) != NULL) {
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
		pin0_send(f, m);
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
PROF_WRAP(PF_FLOWCTL, flowctl_process)
static MSFilterDesc b200_flow_control_desc = {.id = MS_AUDIO_FLOW_CONTROL_ID,
                                              .name = "MSAudioFlowControl",
                                              .text = "B200: flow control filter dropping samples when too many are queued (libmsb200dsp)",
                                              .category = MS_FILTER_OTHER,
                                              .ninputs = 1,
                                              .noutputs = 1,
                                              .init = flowctl_init,
                                              .preprocess = flowctl_preprocess,
                                              .process = flowctl_process_timed,
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
		pin0_send(f, m);
	while ((m = pin0_next(f)) != NULL) {
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
				batch_tick(s->batch, f->ticker->ticks);
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
		pin0_send(f, m);
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
		pin0_send(f, m);
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
PROF_WRAP(PF_PLC, plc_process)
static MSFilterDesc b200_generic_plc_desc = {.id = MS_GENERIC_PLC_ID,
                                             .name = "MSGenericPLC",
                                             .text = "B200: generic packet-loss concealment (libmsb200dsp)",
                                             .category = MS_FILTER_OTHER,
                                             .ninputs = 1,
                                             .noutputs = 1,
                                             .init = plc_init,
                                             .preprocess = plc_preprocess,
                                             .process = plc_process_timed,
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
__attribute__((visibility("default"))) void msb200_filters_host_profile_reset(void) { /* e.g. after the warm-up ticks (joins, set-up) */
	memset(g_prof, 0, sizeof(g_prof));
}
/* MSB200_PROFILE=1: one line of JSON, {"<kind>": {"ms": total milliseconds inside process(), "calls": n}, ...} */
__attribute__((visibility("default"))) int msb200_filters_host_profile(char *buf, int size) {
	int k, i, n = 0;
	if (!buf || size < 2) return 0;
	n += snprintf(buf + n, (size_t)(size - n), "{");
	for (k = 0; k < PF_N && n < size; ++k) {
		uint64_t ns = 0, calls = 0;
		for (i = 0; i < PROF_STRIPES; ++i) {
			ns += __atomic_load_n(&g_prof[i].ns[k], __ATOMIC_RELAXED);
			calls += __atomic_load_n(&g_prof[i].calls[k], __ATOMIC_RELAXED);
		}
		if (calls == 0) continue;
		n += snprintf(buf + n, (size_t)(size - n), "%s\"%s\": {\"ms\": %.3f, \"calls\": %llu}", n > 1 ? ", " : "", g_prof_names[k],
		              (double)ns / 1e6, (unsigned long long)calls);
	}
	if (n < size) n += snprintf(buf + n, (size_t)(size - n), "}");
	return n;
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
