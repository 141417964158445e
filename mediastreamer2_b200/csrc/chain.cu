// chain.cu — the BASELINE cfg2 audio graph kept resident on the device, one call per 10 ms MSTicker tick:
//
//   ref[in_rate] -> MSResample --\
//                                 +--> MSSpeexEC(rate, tail) -> MSVolume(gain) [-> MSAudioMixer, rooms of P pins]
//   mic[in_rate] -> MSResample --/
//
// (graph shape: /root/reference/src/voip/audiostream.c:1798-1832; conference: src/voip/audioconference.c:209-257).
// The re-framing the reference does with MSBufferizer objects between filters of different block sizes is done with
// block-aligned circular buffers in HBM and integer bookkeeping on the host (all streams advance in lockstep):
//   resampler 10 ms blocks -> EC frames      speexec.c:252-259 (ms_bufferizer_read(&s->echo, ..., framesize*2))
//   EC frames -> mixer 10 ms blocks          audiomixer.c:78-90  (ms_bufferizer_read(..., bytespertick))
// Ring capacities are lcm(tick, frame) samples so that neither a tick-sized nor a frame-sized block ever wraps.
// Per tick: 1 resample launch for both resamplers (2 when they are not integer-ratio up-samplers), 1 AEC launch (0..2 frames
// inside), 1 volume launch that also hands the tick's blocks to the caller (banks of 256+ streams; else a small hand-out
// launch), or volume + 1 mixer launch for conference chains.
//
// Overlap mode (msb200_chain_set_overlap; inside msb200_chain_submit with MSB200_CHAIN_OVERLAP=1): the echo canceller is 90 % of a tick and
// the only kernel that fills the chip; the resamplers (before it) and the volume + hand-out copies (after it) are small,
// latency-bound launches. They move to two side streams: `pre` runs the resamplers of tick T+1 and `post` the volume and
// hand-out of tick T-1 while the context's stream runs the canceller of tick T, so the cancellers of consecutive ticks go
// back to back. Events order what depends on what (resample T -> AEC T -> volume T) and bound how far a side stream may
// run ahead or lag behind, so that no ring slot is rewritten while a pending kernel still needs it (see chain_tick_impl).
#include "msb200_internal.h"

struct msb200_chain {
	msb200_ctx *ctx;
	msb200_chain_params p;
	int S, tick_in, tick, F, cap; // cap = lcm(tick, F) samples per stream in each ring
	msb200_resample *rs_ref, *rs_mic;
	msb200_aec *aec;
	msb200_volume *vol;
	msb200_mixer *mix;
	short *d_ref_ring, *d_mic_ring, *d_out_ring;
	int wpos_in;   // next write position (samples) in the EC input rings
	int fill_in;   // samples buffered and not yet consumed by the EC
	int rframe_in; // next EC input frame index
	int wframe_out; // next EC output frame index in the output ring
	int avail_out;  // samples available to the mixer
	int rpos_out;   // mixer read position (samples)
	uint8_t *d_present1, *d_present0;
	short *d_in_ref, *d_in_mic, *d_stage_out;
	int max_out;
	int launches_last_tick;
	// pipelined host path (msb200_chain_submit / _wait): copy streams, double buffers, events
	cudaStream_t s_in, s_out;
	short *pd_in_ref[2], *pd_in_mic[2], *pd_stage[2];
	cudaEvent_t ev_in[2], ev_done[2], ev_out[2];
	int pipe_ready;            // streams / buffers / events created
	unsigned long long n_submitted, n_collected;
	// overlap mode: side streams and event rings (index = overlap tick number & 3)
	cudaStream_t s_pre, s_post;
	cudaEvent_t ev_pre[4], ev_aec[4], ev_post[4], ev_switch;
	int ov_ready;              // streams / events created
	int ov_enabled;            // msb200_chain_set_overlap
	int ov_active;             // the last tick ran in overlap mode (side streams may hold work the context's stream has not joined)
	unsigned long long n_ov;   // overlap ticks since the side streams were last synchronised with the context's stream
	// optional per-launch timing of the AEC kernel (bench.py roofline)
	bool timing;
	std::vector<cudaEvent_t> *ev; // pairs (start, stop), reused round-robin after being drained
	int ev_used, t_launches, t_frames;
	float t_ms;
};

// hand-out of a tick's EC / volume blocks: ring frames [wframe0, wframe0 + nframes) of every stream -> dst[stream][max_out],
// 16 bytes per thread. A kernel, not cudaMemcpy2DAsync: a device-to-device copy is served by a copy engine, and in the
// pipelined host path the copy engines are busy with the neighbouring ticks' PCIe transfers — the kernel stream then waited
// for them between one tick's volume and the next tick's resamplers.
__global__ void chain_handout_kernel(const short *__restrict__ ring, short *__restrict__ dst, int S, int F, int cap, int K, int max_out,
                                     int wframe0, int nframes) {
	const int vec_per_frame = F >> 3, vec_per_stream = nframes * vec_per_frame;
	const long total = (long)S * vec_per_stream;
	for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
		const int stream = (int)(i / vec_per_stream), rem = (int)(i - (long)stream * vec_per_stream);
		const int f = rem / vec_per_frame, v = rem - f * vec_per_frame;
		int pos = wframe0 + f;
		if (pos >= K) pos -= K;
		const int4 *src = reinterpret_cast<const int4 *>(ring + (size_t)stream * cap + (size_t)pos * F) + v;
		reinterpret_cast<int4 *>(dst + (size_t)stream * max_out + (size_t)f * F)[v] = *src;
	}
}

static int chain_drain_events(msb200_chain *c) {
	for (int i = 0; i + 1 < c->ev_used; i += 2) {
		float ms = 0.f;
		MSB200_CUDA(cudaEventSynchronize((*c->ev)[(size_t)i + 1]));
		MSB200_CUDA(cudaEventElapsedTime(&ms, (*c->ev)[(size_t)i], (*c->ev)[(size_t)i + 1]));
		c->t_ms += ms;
	}
	c->ev_used = 0;
	return MSB200_OK;
}

static int gcd_i(int a, int b) {
	while (b) {
		int t = a % b;
		a = b;
		b = t;
	}
	return a;
}

extern "C" {

int msb200_chain_create(msb200_ctx *ctx, const msb200_chain_params *p, msb200_chain **out) {
	MSB200_CHECK_ARG(ctx && p && out && p->n_streams > 0 && p->in_rate > 0 && p->rate > 0);
	MSB200_CHECK_ARG(p->in_rate % 100 == 0 && p->rate % 100 == 0 && p->in_rate != p->rate);
	MSB200_CHECK_ARG(p->mixer_pins == 0 || (p->mixer_pins >= 2 && p->mixer_pins <= 50 && p->n_streams % p->mixer_pins == 0));
	msb200_chain *c = new msb200_chain();
	memset(c, 0, sizeof(*c));
	c->ctx = ctx;
	c->p = *p;
	c->ev = new std::vector<cudaEvent_t>();
	c->S = p->n_streams;
	c->tick_in = p->in_rate / 100;
	c->tick = p->rate / 100;
	c->F = msb200_aec_frame_size_for_rate(p->rate, p->framesize_at_8000 > 0 ? p->framesize_at_8000 : 64);
	c->cap = c->tick / gcd_i(c->tick, c->F) * c->F;
	int r;
	if ((r = msb200_resample_create(ctx, c->S, p->in_rate, p->rate, 1, c->tick_in, &c->rs_ref))) return r;
	if ((r = msb200_resample_create(ctx, c->S, p->in_rate, p->rate, 1, c->tick_in, &c->rs_mic))) return r;
	if ((r = msb200_aec_create(ctx, c->S, p->rate, p->tail_length_ms > 0 ? p->tail_length_ms : 250,
	                           p->framesize_at_8000 > 0 ? p->framesize_at_8000 : 64, &c->aec))) return r;
	if ((r = msb200_volume_create(ctx, c->S, p->rate, c->F, &c->vol))) return r;
	if ((r = msb200_volume_set_gain(c->vol, -1, p->volume_gain))) return r;
	if (p->mixer_pins > 0) {
		if ((r = msb200_mixer_create(ctx, c->S / p->mixer_pins, p->mixer_pins, c->tick, 1, &c->mix))) return r;
	}
	size_t ring_bytes = (size_t)c->S * c->cap * sizeof(short);
	MSB200_CUDA(cudaMalloc(&c->d_ref_ring, ring_bytes));
	MSB200_CUDA(cudaMalloc(&c->d_mic_ring, ring_bytes));
	MSB200_CUDA(cudaMalloc(&c->d_out_ring, ring_bytes));
	MSB200_CUDA(cudaMemset(c->d_ref_ring, 0, ring_bytes));
	MSB200_CUDA(cudaMemset(c->d_mic_ring, 0, ring_bytes));
	MSB200_CUDA(cudaMemset(c->d_out_ring, 0, ring_bytes));
	MSB200_CUDA(cudaMalloc(&c->d_present1, (size_t)c->S));
	MSB200_CUDA(cudaMalloc(&c->d_present0, (size_t)c->S));
	MSB200_CUDA(cudaMemset(c->d_present1, 1, (size_t)c->S));
	MSB200_CUDA(cudaMemset(c->d_present0, 0, (size_t)c->S));
	c->max_out = p->mixer_pins > 0 ? c->tick : ((c->F - 1 + c->tick) / c->F) * c->F;
	MSB200_CUDA(cudaMalloc(&c->d_in_ref, (size_t)c->S * c->tick_in * sizeof(short)));
	MSB200_CUDA(cudaMalloc(&c->d_in_mic, (size_t)c->S * c->tick_in * sizeof(short)));
	MSB200_CUDA(cudaMalloc(&c->d_stage_out, (size_t)c->S * c->max_out * sizeof(short)));
	*out = c;
	return MSB200_OK;
}

void msb200_chain_destroy(msb200_chain *c) {
	if (!c) return;
	cudaStreamSynchronize(c->ctx->stream);
	msb200_resample_destroy(c->rs_ref);
	msb200_resample_destroy(c->rs_mic);
	msb200_aec_destroy(c->aec);
	msb200_volume_destroy(c->vol);
	msb200_mixer_destroy(c->mix);
	cudaFree(c->d_ref_ring);
	cudaFree(c->d_mic_ring);
	cudaFree(c->d_out_ring);
	cudaFree(c->d_present1);
	cudaFree(c->d_present0);
	cudaFree(c->d_in_ref);
	cudaFree(c->d_in_mic);
	cudaFree(c->d_stage_out);
	if (c->ov_ready) {
		cudaStreamSynchronize(c->s_pre);
		cudaStreamSynchronize(c->s_post);
	}
	if (c->pipe_ready) {
		cudaStreamSynchronize(c->s_in);
		cudaStreamSynchronize(c->s_out);
		for (int k = 0; k < 2; ++k) {
			cudaFree(c->pd_in_ref[k]);
			cudaFree(c->pd_in_mic[k]);
			cudaFree(c->pd_stage[k]);
			cudaEventDestroy(c->ev_in[k]);
			cudaEventDestroy(c->ev_done[k]);
			cudaEventDestroy(c->ev_out[k]);
		}
		cudaStreamDestroy(c->s_in);
		cudaStreamDestroy(c->s_out);
	}
	if (c->ov_ready) {
		for (int k = 0; k < 4; ++k) {
			cudaEventDestroy(c->ev_pre[k]);
			cudaEventDestroy(c->ev_aec[k]);
			cudaEventDestroy(c->ev_post[k]);
		}
		cudaEventDestroy(c->ev_switch);
		cudaStreamDestroy(c->s_pre);
		cudaStreamDestroy(c->s_post);
	}
	for (cudaEvent_t e : *c->ev) cudaEventDestroy(e);
	delete c->ev;
	delete c;
}

int msb200_chain_max_out_samples(msb200_chain *c) {
	return c ? c->max_out : MSB200_EINVAL;
}
int msb200_chain_next_out_samples(msb200_chain *c) {
	if (!c) return MSB200_EINVAL;
	if (c->mix) return c->tick;
	return ((c->fill_in + c->tick) / c->F) * c->F;
}
int msb200_chain_launches_per_tick(msb200_chain *c) {
	return c ? c->launches_last_tick : MSB200_EINVAL;
}
int msb200_chain_enable_kernel_timing(msb200_chain *c, int enabled) {
	MSB200_CHECK_ARG(c);
	c->timing = enabled != 0;
	return MSB200_OK;
}
int msb200_chain_get_kernel_timing(msb200_chain *c, float *aec_ms, int *aec_launches, int *aec_frames) {
	MSB200_CHECK_ARG(c);
	int r = chain_drain_events(c);
	if (r) return r;
	if (aec_ms) *aec_ms = c->t_ms;
	if (aec_launches) *aec_launches = c->t_launches;
	if (aec_frames) *aec_frames = c->t_frames;
	c->t_ms = 0.f;
	c->t_launches = c->t_frames = 0;
	return MSB200_OK;
}
msb200_aec *msb200_chain_aec(msb200_chain *c) {
	return c ? c->aec : nullptr;
}

} // extern "C"

static int chain_overlap_init(msb200_chain *c) {
	if (c->ov_ready) return MSB200_OK;
	MSB200_CUDA(cudaStreamCreateWithFlags(&c->s_pre, cudaStreamNonBlocking));
	MSB200_CUDA(cudaStreamCreateWithFlags(&c->s_post, cudaStreamNonBlocking));
	for (int k = 0; k < 4; ++k) {
		MSB200_CUDA(cudaEventCreateWithFlags(&c->ev_pre[k], cudaEventDisableTiming));
		MSB200_CUDA(cudaEventCreateWithFlags(&c->ev_aec[k], cudaEventDisableTiming));
		MSB200_CUDA(cudaEventCreateWithFlags(&c->ev_post[k], cudaEventDisableTiming));
	}
	MSB200_CUDA(cudaEventCreateWithFlags(&c->ev_switch, cudaEventDisableTiming));
	c->ov_ready = 1;
	return MSB200_OK;
}
// overlap is possible when the rings have room for the run-ahead it allows: the input rings hold the canceller's partial
// frame plus two ticks (resample T+1 may run while AEC T is pending), the output ring the frames of two ticks (AEC T may
// write while volume T-1 is pending); conference chains keep the serial order (the mixer reads the output ring)
static bool chain_can_overlap(const msb200_chain *c) {
	return !c->mix && c->cap >= c->F - 1 + 2 * c->tick && c->cap / c->F >= 2 * (c->max_out / c->F);
}
// the context's stream waits for what the side streams still hold: after this, stream order on the context's stream means
// completion again
static int chain_join(msb200_chain *c) {
	if (!c->ov_active) return MSB200_OK;
	if (c->n_ov > 0) MSB200_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->ev_post[(c->n_ov - 1) & 3], 0)); // post T follows AEC T follows pre T
	c->ov_active = 0;
	c->n_ov = 0;
	return MSB200_OK;
}

// one tick of the graph. overlap: see the header of this file; in_ready (optional): the inputs are complete once this event
// fires (otherwise they must be complete at the call); out_free (optional): d_out may be written once this event fires
static int chain_tick_impl(msb200_chain *c, const void *d_ref_in, const void *d_mic_in, void *d_out, int *out_samples, bool overlap,
                           cudaEvent_t in_ready, cudaEvent_t out_free) {
	const uint64_t l0 = c->ctx->launches;
	cudaStream_t const sA = c->ctx->stream;
	cudaStream_t s = sA, sB = sA, sC = sA;
	int r, got = 0, handed_out = 0;
	if (overlap) {
		if ((r = chain_overlap_init(c))) return r;
		sB = c->s_pre;
		sC = c->s_post;
		if (!c->ov_active) { // entering overlap mode: the side streams start behind everything the context's stream holds
			MSB200_CUDA(cudaEventRecord(c->ev_switch, sA));
			MSB200_CUDA(cudaStreamWaitEvent(sB, c->ev_switch, 0));
			MSB200_CUDA(cudaStreamWaitEvent(sC, c->ev_switch, 0));
			c->ov_active = 1;
			c->n_ov = 0;
		}
		if (in_ready) MSB200_CUDA(cudaStreamWaitEvent(sB, in_ready, 0));
		// the resamplers of tick T rewrite ring samples that the canceller of tick T-2 was the last to read
		if (c->n_ov >= 2) MSB200_CUDA(cudaStreamWaitEvent(sB, c->ev_aec[(c->n_ov - 2) & 3], 0));
	} else {
		if ((r = chain_join(c))) return r;
		if (in_ready) MSB200_CUDA(cudaStreamWaitEvent(sA, in_ready, 0));
		if (out_free) MSB200_CUDA(cudaStreamWaitEvent(sA, out_free, 0));
	}
	// 1. both resamplers write their 10 ms block straight into the EC input rings
	c->ctx->stream = sB;
	// (one launch for both when they are integer-ratio up-samplers in phase, else two: msb200i_resample_launch_pair)
	r = msb200i_resample_launch_pair(c->rs_ref, c->rs_mic, d_ref_in, d_mic_in, c->tick_in, c->tick_in, c->d_ref_ring, c->d_mic_ring, c->cap,
	                                 c->wpos_in, c->cap, &got);
	if (r == MSB200_OK && got != c->tick) {
		msb200_set_error("chain: resampler produced %d samples for a %d-sample tick (non-integer rate ratio?)", got, c->tick);
		r = MSB200_ESTATE;
	}
	c->ctx->stream = sA;
	if (r) return r;
	if (overlap) {
		const int k = (int)(c->n_ov & 3);
		MSB200_CUDA(cudaEventRecord(c->ev_pre[k], sB));
		MSB200_CUDA(cudaStreamWaitEvent(sA, c->ev_pre[k], 0));
		// the canceller of tick T rewrites output-ring frames that the volume / hand-out of tick T-2 was the last to touch
		if (c->n_ov >= 2) MSB200_CUDA(cudaStreamWaitEvent(sA, c->ev_post[(c->n_ov - 2) & 3], 0));
	}
	c->wpos_in = (c->wpos_in + c->tick) % c->cap;
	c->fill_in += c->tick;
	// 2. the EC consumes whole frames (speexec.c:256), volume runs per EC output block (msvolume.c:505-512)
	const int nframes = c->fill_in / c->F;
	c->fill_in -= nframes * c->F;
	const int K = c->cap / c->F;
	const int wframe0 = c->wframe_out;
	if (nframes > 0) {
		if (c->timing) {
			if (c->ev_used + 2 > (int)c->ev->size()) {
				if (c->ev->size() >= 4096) {
					if ((r = chain_drain_events(c))) return r;
				} else {
					cudaEvent_t a, b;
					MSB200_CUDA(cudaEventCreate(&a));
					MSB200_CUDA(cudaEventCreate(&b));
					c->ev->push_back(a);
					c->ev->push_back(b);
				}
			}
			MSB200_CUDA(cudaEventRecord((*c->ev)[(size_t)c->ev_used], s));
		}
		if ((r = msb200i_aec_launch(c->aec, c->d_mic_ring, c->d_ref_ring, c->cap, c->rframe_in, K, c->d_out_ring, c->cap,
		                            c->wframe_out, K, nframes))) return r;
		if (c->timing) {
			MSB200_CUDA(cudaEventRecord((*c->ev)[(size_t)c->ev_used + 1], s));
			c->ev_used += 2;
			c->t_launches++;
			c->t_frames += nframes;
		}
		c->rframe_in = (c->rframe_in + nframes) % K;
	}
	if (overlap) {
		const int k = (int)(c->n_ov & 3);
		MSB200_CUDA(cudaEventRecord(c->ev_aec[k], sA));
		MSB200_CUDA(cudaStreamWaitEvent(sC, c->ev_aec[k], 0));
		if (out_free) MSB200_CUDA(cudaStreamWaitEvent(sC, out_free, 0));
		s = sC;
	}
	if (nframes > 0) {
		c->ctx->stream = sC;
		// (without a mixer the lane kernel also writes the tick's blocks to the caller's buffer: no hand-out launch)
		r = msb200i_volume_launch(c->vol, c->d_out_ring, c->F, c->cap, nframes, c->wframe_out, K, nullptr, c->mix ? nullptr : d_out, c->max_out,
		                          &handed_out);
		c->ctx->stream = sA;
		if (r) return r;
		c->wframe_out = (c->wframe_out + nframes) % K;
	}
	if (!c->mix) {
		// 3a. hand the EC/volume output blocks to the caller: [stream][max_out], first nframes*F valid
		if (handed_out) {
		} else if (nframes > 0 && nframes <= K && (c->F & 7) == 0 && ((uintptr_t)d_out & 15) == 0) {
			const long vecs = (long)c->S * nframes * (c->F >> 3);
			const long want = (vecs + 255) / 256;
			c->ctx->stream = s;
			MSB200_LAUNCH(c->ctx, chain_handout_kernel, (int)(want < (long)c->ctx->sm_count * 8 ? want : (long)c->ctx->sm_count * 8), 256, 0,
			              (const short *)c->d_out_ring, (short *)d_out, c->S, c->F, c->cap, K, c->max_out, wframe0, nframes);
			c->ctx->stream = sA;
		} else
		for (int f = 0; f < nframes; ++f) {
			const int pos = (wframe0 + f) % K;
			MSB200_CUDA(cudaMemcpy2DAsync((short *)d_out + (size_t)f * c->F, (size_t)c->max_out * 2,
			                              c->d_out_ring + (size_t)pos * c->F, (size_t)c->cap * 2, (size_t)c->F * 2,
			                              (size_t)c->S, cudaMemcpyDeviceToDevice, s));
		}
		if (out_samples) *out_samples = nframes * c->F;
	} else {
		// 3b. conference mix: each pin contributes one tick when its bufferizer holds one, zeros otherwise
		c->avail_out += nframes * c->F;
		const bool have = c->avail_out >= c->tick;
		if ((r = msb200i_mixer_launch(c->mix, c->d_out_ring + c->rpos_out, c->cap, have ? c->d_present1 : c->d_present0, d_out))) return r;
		if (have) {
			c->rpos_out = (c->rpos_out + c->tick) % c->cap;
			c->avail_out -= c->tick;
		}
		if (out_samples) *out_samples = c->tick;
	}
	if (overlap) {
		MSB200_CUDA(cudaEventRecord(c->ev_post[c->n_ov & 3], sC));
		c->n_ov++;
	}
	c->launches_last_tick = (int)(c->ctx->launches - l0);
	return MSB200_OK;
}

extern "C" {

int msb200_chain_tick_dev(msb200_chain *c, const void *d_ref_in, const void *d_mic_in, void *d_out, int *out_samples) {
	MSB200_CHECK_ARG(c && d_ref_in && d_mic_in && d_out);
	return chain_tick_impl(c, d_ref_in, d_mic_in, d_out, out_samples, c->ov_enabled && chain_can_overlap(c), nullptr, nullptr);
}
int msb200_chain_set_overlap(msb200_chain *c, int enabled) {
	MSB200_CHECK_ARG(c);
	c->ov_enabled = enabled != 0;
	return enabled ? MSB200_OK : chain_join(c);
}
int msb200_chain_join(msb200_chain *c) {
	MSB200_CHECK_ARG(c);
	return chain_join(c);
}

int msb200_chain_tick(msb200_chain *c, const int16_t *ref_in, const int16_t *mic_in, int16_t *out, int *out_samples) {
	MSB200_CHECK_ARG(c && ref_in && mic_in && out);
	cudaStream_t s = c->ctx->stream;
	const size_t in_bytes = (size_t)c->S * c->tick_in * sizeof(short);
	MSB200_CUDA(cudaMemcpyAsync(c->d_in_ref, ref_in, in_bytes, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaMemcpyAsync(c->d_in_mic, mic_in, in_bytes, cudaMemcpyHostToDevice, s));
	int n = 0;
	int r = chain_tick_impl(c, c->d_in_ref, c->d_in_mic, c->d_stage_out, &n, false, nullptr, nullptr); // serial order: one stream
	if (r) return r;
	if (n > 0)
		MSB200_CUDA(cudaMemcpy2DAsync(out, (size_t)c->max_out * 2, c->d_stage_out, (size_t)c->max_out * 2, (size_t)n * 2,
		                              (size_t)c->S, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	if (out_samples) *out_samples = n;
	return MSB200_OK;
}

// ---- pipelined host path: tick T's input copy and tick T-1's output copy overlap tick T's kernels.
// Three streams: s_in (H2D), the context's stream (kernels), s_out (D2H); double buffers on the device; per slot the
// events in -> done -> out chain the three stages, and `done` / `out` of the slot's previous use gate its reuse.
static int chain_pipe_init(msb200_chain *c) {
	if (c->pipe_ready) return MSB200_OK;
	MSB200_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
	MSB200_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
	for (int k = 0; k < 2; ++k) {
		MSB200_CUDA(cudaMalloc(&c->pd_in_ref[k], (size_t)c->S * c->tick_in * sizeof(short)));
		MSB200_CUDA(cudaMalloc(&c->pd_in_mic[k], (size_t)c->S * c->tick_in * sizeof(short)));
		MSB200_CUDA(cudaMalloc(&c->pd_stage[k], (size_t)c->S * c->max_out * sizeof(short)));
		MSB200_CUDA(cudaEventCreateWithFlags(&c->ev_in[k], cudaEventDisableTiming));
		MSB200_CUDA(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
		MSB200_CUDA(cudaEventCreateWithFlags(&c->ev_out[k], cudaEventDisableTiming));
	}
	c->pipe_ready = 1;
	return MSB200_OK;
}

int msb200_chain_submit(msb200_chain *c, const int16_t *ref_in, const int16_t *mic_in, int16_t *out, int *out_samples) {
	MSB200_CHECK_ARG(c && ref_in && mic_in && out);
	int r = chain_pipe_init(c);
	if (r) return r;
	if (c->n_submitted - c->n_collected >= 2) {
		msb200_set_error("chain: two ticks are already in flight; call msb200_chain_wait() first");
		return MSB200_ESTATE;
	}
	const int k = (int)(c->n_submitted & 1);
	const bool reused = c->n_submitted >= 2; // the slot has a previous occupant whose stages must have drained
	cudaStream_t sc = c->ctx->stream;
	const size_t in_bytes = (size_t)c->S * c->tick_in * sizeof(short);
	// stage 1: inputs -> device (after the kernels of tick T-2 stopped reading this slot's input buffers)
	if (reused) MSB200_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_done[k], 0));
	static const int dbg = getenv("MSB200_CHAIN_DBG") ? atoi(getenv("MSB200_CHAIN_DBG")) : 0; // 1: no H2D, 2: no D2H, 3: neither (interference experiments)
	if (!(dbg & 1)) {
	MSB200_CUDA(cudaMemcpyAsync(c->pd_in_ref[k], ref_in, in_bytes, cudaMemcpyHostToDevice, c->s_in));
	MSB200_CUDA(cudaMemcpyAsync(c->pd_in_mic[k], mic_in, in_bytes, cudaMemcpyHostToDevice, c->s_in));
	}
	MSB200_CUDA(cudaEventRecord(c->ev_in[k], c->s_in));
	// stage 2: the tick's kernels (after the inputs landed and tick T-2's output left this slot's staging buffer). In
	// overlap mode the resamplers wait for the inputs on their side stream and the hand-out copies for the staging buffer
	// on theirs; `done` is recorded behind the hand-out, which follows the canceller, which follows the resamplers
	// (off unless MSB200_CHAIN_OVERLAP=1: with host buffers the tick is not bounded by the small kernels — measured on
	// B200, 4096 streams: 0.908 ms per tick without, 0.938 ms with the side streams' extra event traffic)
	static const bool ov_env = getenv("MSB200_CHAIN_OVERLAP") && atoi(getenv("MSB200_CHAIN_OVERLAP")) != 0;
	const bool ov = ov_env && chain_can_overlap(c);
	int n = 0;
	if ((r = chain_tick_impl(c, c->pd_in_ref[k], c->pd_in_mic[k], c->pd_stage[k], &n, ov, c->ev_in[k], reused ? c->ev_out[k] : nullptr))) return r;
	MSB200_CUDA(cudaEventRecord(c->ev_done[k], ov ? c->s_post : sc));
	// stage 3: output -> host
	MSB200_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_done[k], 0));
	// both sides are [stream][max_out]: when the block fills its rows the hand-over is ONE linear copy instead of a pitched
	// one of n_streams short rows (A/B: MSB200_CHAIN_D2H_2D=1 keeps the pitched copy)
	static const bool d2h_2d = getenv("MSB200_CHAIN_D2H_2D") && atoi(getenv("MSB200_CHAIN_D2H_2D")) != 0;
	if (dbg & 2) {
	} else if (n > 0 && n == c->max_out && !d2h_2d)
		MSB200_CUDA(cudaMemcpyAsync(out, c->pd_stage[k], (size_t)c->S * c->max_out * 2, cudaMemcpyDeviceToHost, c->s_out));
	else if (n > 0)
		MSB200_CUDA(cudaMemcpy2DAsync(out, (size_t)c->max_out * 2, c->pd_stage[k], (size_t)c->max_out * 2, (size_t)n * 2,
		                              (size_t)c->S, cudaMemcpyDeviceToHost, c->s_out));
	MSB200_CUDA(cudaEventRecord(c->ev_out[k], c->s_out));
	c->n_submitted++;
	if (out_samples) *out_samples = n;
	return MSB200_OK;
}

int msb200_chain_wait(msb200_chain *c) {
	MSB200_CHECK_ARG(c);
	if (c->n_collected >= c->n_submitted) return MSB200_OK; // nothing in flight
	const int k = (int)(c->n_collected & 1);
	MSB200_CUDA(cudaEventSynchronize(c->ev_out[k]));
	c->n_collected++;
	return MSB200_OK;
}

} // extern "C"
