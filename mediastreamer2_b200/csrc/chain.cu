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
// Per tick: 2 resample launches, 1 AEC launch (0..2 frames inside), 1 volume launch, optionally 1 mixer launch.
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
	// optional per-launch timing of the AEC kernel (bench.py roofline)
	bool timing;
	std::vector<cudaEvent_t> *ev; // pairs (start, stop), reused round-robin after being drained
	int ev_used, t_launches, t_frames;
	float t_ms;
};

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

int msb200_chain_tick_dev(msb200_chain *c, const void *d_ref_in, const void *d_mic_in, void *d_out, int *out_samples) {
	MSB200_CHECK_ARG(c && d_ref_in && d_mic_in && d_out);
	const uint64_t l0 = c->ctx->launches;
	cudaStream_t s = c->ctx->stream;
	int r, got = 0;
	// 1. both resamplers write their 10 ms block straight into the EC input rings
	if ((r = msb200i_resample_launch(c->rs_ref, d_ref_in, c->tick_in, c->tick_in, c->d_ref_ring, c->cap, c->wpos_in, c->cap, &got))) return r;
	if (got != c->tick) {
		msb200_set_error("chain: resampler produced %d samples for a %d-sample tick (non-integer rate ratio?)", got, c->tick);
		return MSB200_ESTATE;
	}
	if ((r = msb200i_resample_launch(c->rs_mic, d_mic_in, c->tick_in, c->tick_in, c->d_mic_ring, c->cap, c->wpos_in, c->cap, &got))) return r;
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
		if ((r = msb200i_volume_launch(c->vol, c->d_out_ring, c->F, c->cap, nframes, c->wframe_out, K))) return r;
		c->wframe_out = (c->wframe_out + nframes) % K;
	}
	if (!c->mix) {
		// 3a. hand the EC/volume output blocks to the caller: [stream][max_out], first nframes*F valid
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
	c->launches_last_tick = (int)(c->ctx->launches - l0);
	return MSB200_OK;
}

int msb200_chain_tick(msb200_chain *c, const int16_t *ref_in, const int16_t *mic_in, int16_t *out, int *out_samples) {
	MSB200_CHECK_ARG(c && ref_in && mic_in && out);
	cudaStream_t s = c->ctx->stream;
	const size_t in_bytes = (size_t)c->S * c->tick_in * sizeof(short);
	MSB200_CUDA(cudaMemcpyAsync(c->d_in_ref, ref_in, in_bytes, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaMemcpyAsync(c->d_in_mic, mic_in, in_bytes, cudaMemcpyHostToDevice, s));
	int n = 0;
	int r = msb200_chain_tick_dev(c, c->d_in_ref, c->d_in_mic, c->d_stage_out, &n);
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
	MSB200_CUDA(cudaMemcpyAsync(c->pd_in_ref[k], ref_in, in_bytes, cudaMemcpyHostToDevice, c->s_in));
	MSB200_CUDA(cudaMemcpyAsync(c->pd_in_mic[k], mic_in, in_bytes, cudaMemcpyHostToDevice, c->s_in));
	MSB200_CUDA(cudaEventRecord(c->ev_in[k], c->s_in));
	// stage 2: the tick's kernels (after the inputs landed and tick T-2's output left this slot's staging buffer)
	MSB200_CUDA(cudaStreamWaitEvent(sc, c->ev_in[k], 0));
	if (reused) MSB200_CUDA(cudaStreamWaitEvent(sc, c->ev_out[k], 0));
	int n = 0;
	if ((r = msb200_chain_tick_dev(c, c->pd_in_ref[k], c->pd_in_mic[k], c->pd_stage[k], &n))) return r;
	MSB200_CUDA(cudaEventRecord(c->ev_done[k], sc));
	// stage 3: output -> host
	MSB200_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_done[k], 0));
	if (n > 0)
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
