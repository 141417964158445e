/* oracle/oracle_bench.c — TEST/BENCH INFRASTRUCTURE (see msb200_oracle.h).
 * Multi-threaded CPU run of the BASELINE cfg2 chain built from the oracle pieces: per stream and per 10 ms tick,
 * 2x resampler (msresample.c:150-161) -> frame re-blocking (speexec.c:252-259) -> echo canceller + preprocessor per
 * frame (speexec.c:297-298) -> volume per block (msvolume.c:505-512). One pthread per requested host core, streams
 * partitioned evenly, free running (no 10 ms sleep) — the CPU-baseline plan of BASELINE.md §3.
 * Used only by bench.py's cpu_baseline / --impl reference legs and by tests. */
#define _GNU_SOURCE
#include "msb200_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
	int first, count, ticks, in_rate, rate, tail_ms;
	float gain;
	const int16_t *ref, *mic;
	int16_t *out;
	long out_samples;
	pthread_barrier_t *bar;
	double t_end;
} worker_t;

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *worker(void *arg) {
	worker_t *w = (worker_t *)arg;
	const int tick_in = w->in_rate / 100, tick = w->rate / 100;
	const int n = w->count;
	orc_resampler **r1 = (orc_resampler **)calloc((size_t)n, sizeof(*r1));
	orc_resampler **r2 = (orc_resampler **)calloc((size_t)n, sizeof(*r2));
	orc_aec **ec = (orc_aec **)calloc((size_t)n, sizeof(*ec));
	orc_volume_state *vol = (orc_volume_state *)calloc((size_t)n, sizeof(*vol));
	int16_t *bref = (int16_t *)calloc((size_t)n * 2048, sizeof(int16_t));
	int16_t *bmic = (int16_t *)calloc((size_t)n * 2048, sizeof(int16_t));
	int *fill = (int *)calloc((size_t)n, sizeof(int));
	long *wr = (long *)calloc((size_t)n, sizeof(long));
	int F = 0;
	for (int i = 0; i < n; ++i) {
		r1[i] = orc_resampler_new(1, w->in_rate, w->rate, 3);
		r2[i] = orc_resampler_new(1, w->in_rate, w->rate, 3);
		ec[i] = orc_aec_new(w->rate, w->tail_ms, 64);
		orc_volume_init(&vol[i], w->rate);
		vol[i].gain = vol[i].target_gain = vol[i].static_gain = w->gain;
		F = orc_aec_frame_size(ec[i]);
	}
	pthread_barrier_wait(w->bar); /* start of the timed region */
	for (int t = 0; t < w->ticks; ++t) {
		for (int i = 0; i < n; ++i) {
			const size_t s = (size_t)(w->first + i);
			const int16_t *ri = w->ref + (s * (size_t)w->ticks + (size_t)t) * (size_t)tick_in;
			const int16_t *mi = w->mic + (s * (size_t)w->ticks + (size_t)t) * (size_t)tick_in;
			int16_t *br = bref + (size_t)i * 2048, *bm = bmic + (size_t)i * 2048;
			int n1 = orc_msresample_block(r1[i], ri, tick_in, br + fill[i]);
			int n2 = orc_msresample_block(r2[i], mi, tick_in, bm + fill[i]);
			(void)n2;
			fill[i] += n1;
			int pos = 0;
			while (fill[i] - pos >= F) {
				int16_t o[1024];
				orc_aec_process_frame(ec[i], bm + pos, br + pos, o);
				orc_volume_process(&vol[i], o, F);
				if (w->out) memcpy(w->out + s * (size_t)w->ticks * (size_t)tick + (size_t)wr[i], o, sizeof(int16_t) * (size_t)F);
				wr[i] += F;
				pos += F;
			}
			memmove(br, br + pos, sizeof(int16_t) * (size_t)(fill[i] - pos));
			memmove(bm, bm + pos, sizeof(int16_t) * (size_t)(fill[i] - pos));
			fill[i] -= pos;
		}
	}
	w->t_end = now_s();
	for (int i = 0; i < n; ++i) {
		w->out_samples += wr[i];
		orc_resampler_free(r1[i]);
		orc_resampler_free(r2[i]);
		orc_aec_free(ec[i]);
	}
	free(r1);
	free(r2);
	free(ec);
	free(vol);
	free(bref);
	free(bmic);
	free(fill);
	free(wr);
	(void)tick;
	return NULL;
}

/* ref/mic: [n_streams][ticks][in_rate/100] s16; out (optional): [n_streams][ticks*rate/100] s16.
 * Returns wall seconds from the common start to the last worker's finish (state construction excluded). */
double orc_chain_bench(int n_streams, int ticks, int threads, int in_rate, int rate, int tail_ms, float gain,
                       const int16_t *ref, const int16_t *mic, int16_t *out, long *out_samples) {
	if (threads < 1) threads = 1;
	if (threads > n_streams) threads = n_streams;
	pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(*th));
	worker_t *ws = (worker_t *)calloc((size_t)threads, sizeof(*ws));
	pthread_barrier_t bar;
	pthread_barrier_init(&bar, NULL, (unsigned)threads + 1);
	int first = 0;
	for (int i = 0; i < threads; ++i) {
		int cnt = n_streams / threads + (i < n_streams % threads ? 1 : 0);
		ws[i] = (worker_t){first, cnt, ticks, in_rate, rate, tail_ms, gain, ref, mic, out, 0, &bar, 0.0};
		first += cnt;
		pthread_create(&th[i], NULL, worker, &ws[i]);
	}
	pthread_barrier_wait(&bar);
	double t0 = now_s(), t1 = t0;
	long total = 0;
	for (int i = 0; i < threads; ++i) {
		pthread_join(th[i], NULL);
		if (ws[i].t_end > t1) t1 = ws[i].t_end;
		total += ws[i].out_samples;
	}
	if (out_samples) *out_samples = total;
	pthread_barrier_destroy(&bar);
	free(th);
	free(ws);
	return t1 - t0;
}
