// audio.cu — MSAudioMixer, MSVolume, MSChannelAdapter, MSEqualizer, MSResample as batched sm_100a kernels.
//
// Every kernel here is HBM/latency bound integer or float32 work: the design rules are coalesced
// 16-byte accesses, shared-memory staging where a block is re-read, and enough CTAs to cover 148 SMs.
// Bit-exactness rules: integer paths reproduce the reference's C semantics (truncating casts and
// divisions, asymmetric +-32767 clamp); float32 state machines and dot products use __fmul_rn/__fadd_rn in the
// reference's evaluation order so that nvcc never contracts them into FMAs.
#include "msb200_internal.h"
#include "mixer_internal.cuh"

#include <cmath>

// ======================================================================================================= mixer
// reference: /root/reference/src/audiofilters/audiomixer.c  (accumulate :33-38, saturate :40-44, apply_gain :46-51,
// channel_process_in :78-90, channel_process_out :113-130, make_output :210-217)

// mode 0: full (sum + outputs); mode 1: partial sums only -> d_sum; mode 2: outputs from given d_sum
// (the fused cross-GPU exchange is its own kernel: mixer_xchg.cu)
template <int VEC>
__global__ void __launch_bounds__(128) mixer_kernel(const short *__restrict__ in, const uint8_t *__restrict__ present,
                                                    const float *__restrict__ gain, const uint8_t *__restrict__ active,
                                                    short *__restrict__ out, int *__restrict__ sum_io, int n_rooms,
                                                    int n_pins, int nwords, int conf_mode, int mode, long in_pin_vecs) {
	typedef typename s16vec<VEC>::type V;
	const int nvec = nwords / VEC;
	const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long)n_rooms * nvec) return;
	const int room = (int)(gid / nvec), col = (int)(gid % nvec);
	const size_t chan0 = (size_t)room * n_pins;
	// in_pin_vecs: distance between consecutive pins' blocks in V units (nvec when the input is dense)
	const V *inv = reinterpret_cast<const V *>(in) + chan0 * in_pin_vecs + col;
	int sum[VEC];
#pragma unroll
	for (int k = 0; k < VEC; ++k) sum[k] = 0;
	if (mode == 0 || mode == 1) {
#pragma unroll 4
		for (int p = 0; p < n_pins; ++p) {
			if (!present[chan0 + p] || !active[chan0 + p]) continue;
			const float g = gain[chan0 + p];
			int s[VEC];
			unpack_s16<VEC>(inv[(size_t)p * in_pin_vecs], s);
#pragma unroll
			for (int k = 0; k < VEC; ++k) sum[k] += mix_contrib(s[k], g);
		}
	} else {
#pragma unroll
		for (int k = 0; k < VEC; ++k) sum[k] = sum_io[(size_t)room * nwords + col * VEC + k];
	}
	if (mode == 1) {
#pragma unroll
		for (int k = 0; k < VEC; ++k) sum_io[(size_t)room * nwords + col * VEC + k] = sum[k];
		return;
	}
	if (!conf_mode) { // make_output :210-217
		int o[VEC];
#pragma unroll
		for (int k = 0; k < VEC; ++k) o[k] = mix_sat(sum[k]);
		reinterpret_cast<V *>(out)[(size_t)room * nvec + col] = pack_s16<VEC>(o);
		return;
	}
	V *outv = reinterpret_cast<V *>(out) + chan0 * nvec + col;
#pragma unroll 4
	for (int p = 0; p < n_pins; ++p) { // channel_process_out :113-130
		int o[VEC];
		if (active[chan0 + p] && present[chan0 + p]) {
			const float g = gain[chan0 + p];
			int s[VEC];
			unpack_s16<VEC>(inv[(size_t)p * in_pin_vecs], s);
#pragma unroll
			for (int k = 0; k < VEC; ++k) o[k] = mix_sat(sum[k] - mix_contrib(s[k], g));
		} else {
#pragma unroll
			for (int k = 0; k < VEC; ++k) o[k] = mix_sat(sum[k]);
		}
		outv[(size_t)p * nvec] = pack_s16<VEC>(o);
	}
}

// Conference rooms of up to 16 pins (BASELINE cfg3's shape): every pin's column is loaded ONCE and stays in registers
// between the sum and the per-listener outputs — as packed post-gain contributions, which fit 16 bits (apply_gain saturates,
// audiomixer.c:46-51) — so the DRAM / L2 traffic is exactly the algorithmic P*2n in + P*2n out. Same arithmetic as
// mixer_kernel mode 0.
#define MIX_REG_PINS 16
template <int VEC>
__global__ void __launch_bounds__(128) mixer_conf_regs_kernel(const short *__restrict__ in, const uint8_t *__restrict__ present,
                                                              const float *__restrict__ gain, const uint8_t *__restrict__ active,
                                                              short *__restrict__ out, int n_rooms, int n_pins, int nwords,
                                                              long in_pin_vecs) {
	typedef typename s16vec<VEC>::type V;
	const int nvec = nwords / VEC;
	const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= (long)n_rooms * nvec) return;
	const int room = (int)(gid / nvec), col = (int)(gid % nvec);
	const size_t chan0 = (size_t)room * n_pins;
	const V *inv = reinterpret_cast<const V *>(in) + chan0 * in_pin_vecs + col;
	V c[MIX_REG_PINS];
	unsigned live = 0;
#pragma unroll
	for (int p = 0; p < MIX_REG_PINS; ++p)
		if (p < n_pins && present[chan0 + p] && active[chan0 + p]) {
			live |= 1u << p;
			c[p] = inv[(size_t)p * in_pin_vecs]; // all loads issued before the first use
		}
	int sum[VEC];
#pragma unroll
	for (int k = 0; k < VEC; ++k) sum[k] = 0;
#pragma unroll
	for (int p = 0; p < MIX_REG_PINS; ++p)
		if (live >> p & 1u) {
			const float g = gain[chan0 + p];
			int s[VEC];
			unpack_s16<VEC>(c[p], s);
#pragma unroll
			for (int k = 0; k < VEC; ++k) {
				s[k] = mix_contrib(s[k], g);
				sum[k] += s[k];
			}
			c[p] = pack_s16<VEC>(s);
		}
	V *outv = reinterpret_cast<V *>(out) + chan0 * nvec + col;
#pragma unroll
	for (int p = 0; p < MIX_REG_PINS; ++p)
		if (p < n_pins) { // channel_process_out :113-130
			int o[VEC];
			if (live >> p & 1u) {
				int s[VEC];
				unpack_s16<VEC>(c[p], s);
#pragma unroll
				for (int k = 0; k < VEC; ++k) o[k] = mix_sat(sum[k] - s[k]);
			} else {
#pragma unroll
				for (int k = 0; k < VEC; ++k) o[k] = mix_sat(sum[k]);
			}
			outv[(size_t)p * nvec] = pack_s16<VEC>(o);
		}
}

int msb200i_mixer_upload(msb200_mixer *m) {
	if (!m->dirty) return MSB200_OK;
	size_t n = (size_t)m->n_rooms * m->n_pins;
	MSB200_CUDA(cudaMemcpyAsync(m->d_gain, m->h_gain.data(), n * sizeof(float), cudaMemcpyHostToDevice, m->ctx->stream));
	MSB200_CUDA(cudaMemcpyAsync(m->d_active, m->h_active.data(), n, cudaMemcpyHostToDevice, m->ctx->stream));
	// the host vectors may be modified right after we return: make the copy complete first
	MSB200_CUDA(cudaStreamSynchronize(m->ctx->stream));
	m->dirty = false;
	return MSB200_OK;
}

static int mixer_launch(msb200_mixer *m, const void *d_in, const void *d_present, void *d_out, void *d_sum, int mode,
                        long in_pin_stride = 0) {
	int r = msb200i_mixer_upload(m);
	if (r) return r;
	const int nw = m->nwords;
	const bool al16 = ((uintptr_t)d_in % 16 == 0) && (d_out == nullptr || (uintptr_t)d_out % 16 == 0);
	if (in_pin_stride <= 0) in_pin_stride = nw;
	int vec = (nw % 8 == 0 && al16) ? 8 : (nw % 4 == 0 && al16) ? 4 : (nw % 2 == 0 && al16) ? 2 : 1;
	while (vec > 1 && in_pin_stride % vec) vec /= 2;
	// prefer more, narrower threads when the grid would not cover the chip (148 SMs x >=4 CTAs of 128)
	while (vec > 2 && (long)m->live * (nw / vec) < 148L * 4 * 128) vec /= 2;
	if (mode == 0 && m->conf_mode && m->n_pins <= MIX_REG_PINS && vec >= 4) { // register-resident rooms
		if (vec == 8 && (long)m->live * (nw / 8) < 148L * 8 * 128) vec = 4; // cover the chip before widening the vectors
		const long nth = (long)m->live * (nw / vec);
		const int g = (int)((nth + 127) / 128);
		if (vec == 8)
			MSB200_LAUNCH(m->ctx, mixer_conf_regs_kernel<8>, g, 128, 0, (const short *)d_in, (const uint8_t *)d_present, m->d_gain,
			              m->d_active, (short *)d_out, m->live, m->n_pins, nw, in_pin_stride / vec);
		else
			MSB200_LAUNCH(m->ctx, mixer_conf_regs_kernel<4>, g, 128, 0, (const short *)d_in, (const uint8_t *)d_present, m->d_gain,
			              m->d_active, (short *)d_out, m->live, m->n_pins, nw, in_pin_stride / vec);
		return MSB200_OK;
	}
	const long nthreads = (long)m->live * (nw / vec);
	const int block = 128, grid = (int)((nthreads + block - 1) / block);
#define MIX_ARGS                                                                                                       \
	(const short *)d_in, (const uint8_t *)d_present, m->d_gain, m->d_active, (short *)d_out, (int *)d_sum, m->live, \
	    m->n_pins, nw, m->conf_mode, mode, in_pin_stride / vec
	switch (vec) {
		case 8: MSB200_LAUNCH(m->ctx, mixer_kernel<8>, grid, block, 0, MIX_ARGS); break;
		case 4: MSB200_LAUNCH(m->ctx, mixer_kernel<4>, grid, block, 0, MIX_ARGS); break;
		case 2: MSB200_LAUNCH(m->ctx, mixer_kernel<2>, grid, block, 0, MIX_ARGS); break;
		default: MSB200_LAUNCH(m->ctx, mixer_kernel<1>, grid, block, 0, MIX_ARGS); break;
	}
#undef MIX_ARGS
	return MSB200_OK;
}

extern "C" {

int msb200_mixer_create(msb200_ctx *ctx, int n_rooms, int n_pins, int nwords, int conf_mode, msb200_mixer **out) {
	MSB200_CHECK_ARG(ctx && out && n_rooms > 0 && n_pins > 0 && n_pins <= 50 /* MIXER_MAX_CHANNELS :29 */ && nwords > 0);
	msb200_mixer *m = new msb200_mixer();
	m->ctx = ctx;
	m->n_rooms = m->live = n_rooms;
	m->n_pins = n_pins;
	m->nwords = nwords;
	m->conf_mode = conf_mode;
	size_t n = (size_t)n_rooms * n_pins;
	m->h_gain.assign(n, 1.0f);  // channel_init :64-70
	m->h_active.assign(n, 1);
	m->dirty = true;
	MSB200_CUDA(cudaMalloc(&m->d_gain, n * sizeof(float)));
	MSB200_CUDA(cudaMalloc(&m->d_active, n));
	*out = m;
	return MSB200_OK;
}
void msb200_mixer_destroy(msb200_mixer *m) {
	if (!m) return;
	cudaStreamSynchronize(m->ctx->stream);
	cudaFree(m->d_gain);
	cudaFree(m->d_active);
	m->in.release();
	m->present.release();
	m->out.release();
	delete m;
}
int msb200_mixer_set_input_gain(msb200_mixer *m, int room, int pin, float gain) {
	MSB200_CHECK_ARG(m && room >= 0 && room < m->n_rooms && pin >= 0 && pin < m->n_pins);
	m->h_gain[(size_t)room * m->n_pins + pin] = gain;
	m->dirty = true;
	return MSB200_OK;
}
int msb200_mixer_set_active(msb200_mixer *m, int room, int pin, int active) {
	MSB200_CHECK_ARG(m && room >= 0 && room < m->n_rooms && pin >= 0 && pin < m->n_pins);
	m->h_active[(size_t)room * m->n_pins + pin] = active ? 1 : 0;
	m->dirty = true;
	return MSB200_OK;
}
int msb200_mixer_process_dev(msb200_mixer *m, const void *d_in, const void *d_present, void *d_out) {
	MSB200_CHECK_ARG(m && d_in && d_present && d_out);
	return mixer_launch(m, d_in, d_present, d_out, nullptr, 0);
}
int msb200_mixer_partial_dev(msb200_mixer *m, const void *d_in, const void *d_present, void *d_sum) {
	MSB200_CHECK_ARG(m && d_in && d_present && d_sum);
	return mixer_launch(m, d_in, d_present, nullptr, d_sum, 1);
}
int msb200_mixer_finish_dev(msb200_mixer *m, const void *d_in, const void *d_present, const void *d_sum, void *d_out) {
	MSB200_CHECK_ARG(m && d_in && d_present && d_sum && d_out);
	return mixer_launch(m, d_in, d_present, d_out, (void *)d_sum, 2);
}
int msb200_mixer_set_live(msb200_mixer *m, int n_live) {
	MSB200_CHECK_ARG(m && n_live >= 0 && n_live <= m->n_rooms);
	m->live = n_live;
	return MSB200_OK;
}
int msb200_mixer_process(msb200_mixer *m, const int16_t *in, const uint8_t *present, int16_t *out) {
	MSB200_CHECK_ARG(m && in && present && out);
	if (m->live == 0) return MSB200_OK;
	size_t nch = (size_t)m->live * m->n_pins;
	size_t in_bytes = nch * m->nwords * 2;
	size_t out_bytes = m->conf_mode ? in_bytes : (size_t)m->live * m->nwords * 2;
	int r;
	if ((r = m->in.reserve(in_bytes)) || (r = m->present.reserve(nch)) || (r = m->out.reserve(out_bytes))) return r;
	cudaStream_t s = m->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(m->in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaMemcpyAsync(m->present.p, present, nch, cudaMemcpyHostToDevice, s));
	if ((r = mixer_launch(m, m->in.p, m->present.p, m->out.p, nullptr, 0))) return r;
	MSB200_CUDA(cudaMemcpyAsync(out, m->out.p, out_bytes, cudaMemcpyDeviceToHost, s));
	MSB200_HOST_DONE(m->ctx);
	return MSB200_OK;
}

} // extern "C"

int msb200i_mixer_launch(msb200_mixer *m, const void *d_in, long in_pin_stride, const void *d_present, void *d_out) {
	MSB200_CHECK_ARG(m && d_in && d_present && d_out);
	return mixer_launch(m, d_in, d_present, d_out, nullptr, 0, in_pin_stride);
}

// ======================================================================================================= volume
// reference: /root/reference/src/audiofilters/msvolume.c light path :503-513
//   update_energy :388-407, volume_noise_gate_process :240-260, apply_gain :409-445, saturate :382-384

struct msb200_volume {
	msb200_ctx *ctx;
	int n, rate, max_block;
	int live; // streams [0, live) are processed (msb200_volume_set_live); == n by default
	msb200_volume *peer_bank; // bank whose states the echo limiter reads (may be this one)
	msb200_volume_state *d_state;
	msb200_devbuf io;
	int kernel_choice = 0; // msb200_volume_set_kernel: 0 = by bank size, 1 = warp per stream, 2 = lane per stream
};

__device__ __forceinline__ int vol_sat(int v) { // :382-384
	return v > 32767 ? 32767 : (v < -32767 ? -32767 : v);
}

// One block's worth of msvolume.c for one stream, given the block's energy sum (accumulated in the reference's order), its
// peak and its sample sum: update_energy :388-407, the echo limiter :201-238, AGC :172-184, the noise gate :240-260 and the
// gain ramp of apply_gain :409-445. Shared by the two kernels below so that they cannot differ.
__device__ __forceinline__ void vol_update(msb200_volume_state &v, float acc, int pk, int dcsum, int nsamples,
                                           const msb200_volume_state *__restrict__ peer_states, int &intgain, int &apply,
                                           int &remove_dc, int &dc_prev) {
	const float max_e = 32768 * 0.7f;
	// en = (float)((sqrt(acc / n) + 1) / max_e)  — float division, then double sqrt/add/div (C promotions)
	float q = __fdiv_rn(acc, (float)nsamples);
	float en = (float)((sqrt((double)q) + 1.0) / (double)max_e);
	v.energy = __fadd_rn(__fmul_rn(en, 0.2f), __fmul_rn(v.energy, (1.0f - 0.2f)));
	v.level_pk = __fdiv_rn((float)pk, max_e);
	v.instant_energy = en;
	float tgain = v.static_gain;
	if (v.peer >= 0 && peer_states) { // volume_echo_avoider_process :201-238
		const float peer_e = peer_states[v.peer].energy;
		if (peer_e > v.lt_speaker_en) v.lt_speaker_en = peer_e;
		else v.lt_speaker_en = __fadd_rn(__fmul_rn(0.005f, peer_e), __fmul_rn(0.995f, v.lt_speaker_en));
		const float mic_spk_ratio = __fdiv_rn(v.energy, __fadd_rn(v.lt_speaker_en, v.ea_thres));
		if (peer_e > v.ea_thres) {
			if (mic_spk_ratio > v.ea_transmit_thres) {
				v.target_gain = v.static_gain;
				v.fast_upramp = 1;
			} else {
				v.target_gain = __fdiv_rn(v.static_gain, __fadd_rn(1.f, __fmul_rn(peer_e, v.force))); // compute_gain :186-189
				v.sustain_dur = v.sustain_time;
			}
		} else if (v.sustain_dur > 0) {
			v.sustain_dur -= (nsamples * 1000) / v.sample_rate;
		} else {
			v.target_gain = v.static_gain;
			v.fast_upramp = 1;
		}
		tgain = v.target_gain;
	}
	if (v.agc_enabled) tgain = __fdiv_rn(tgain, __fdiv_rn(__fadd_rn(0.5f, v.level_pk), 1.f)); // volume_agc_process :172-184
	if (v.noise_gate_enabled) {
		float t = v.ng_floorgain;
		if (v.instant_energy > v.ng_threshold) {
			v.ng_noise_dur = 400;
			t = 1.0f;
		} else if (v.ng_noise_dur > 0) {
			v.ng_noise_dur -= (nsamples * 1000) / v.sample_rate;
			t = 1.0f;
		}
		v.ng_gain = __fadd_rn(__fmul_rn(v.ng_gain, 0.75f), __fmul_rn(t, 0.25f));
	}
	if (v.gain < tgain) {
		if (v.gain < v.ng_floorgain) v.gain = v.ng_floorgain;
		v.gain = __fmul_rn(v.gain, v.fast_upramp ? (1 + 0.4f * 3) : __fadd_rn(1.f, v.vol_upramp));
		if (v.gain > tgain) v.gain = tgain;
	} else if (v.gain > tgain) {
		v.gain = __fmul_rn(v.gain, 1 - 0.4f);
		if (v.gain < tgain) v.gain = tgain;
		v.fast_upramp = 0;
	}
	float gain = __fmul_rn(v.gain, v.ng_gain);
	intgain = __float2int_rz(__fmul_rn(gain, 4096.f));
	remove_dc = v.remove_dc;
	dc_prev = v.dc_offset;
	apply = remove_dc ? 1 : (gain != 1.0f);
	if (remove_dc) v.dc_offset = (v.dc_offset * 7 + dcsum * 2 / (nsamples * 2)) / 8;
}

// One warp per stream. The block is staged in shared memory by coalesced loads; lane 0 reproduces the reference's
// strictly sequential float32 accumulation (acc += s*s) so `energy` is bit-exact; peak and DC sums are integer and
// reduced across the warp; then all lanes apply the Q12 gain and store.
__global__ void __launch_bounds__(256) volume_kernel(short *__restrict__ io, msb200_volume_state *__restrict__ st,
                                                     int n_streams, int nsamples, int stride, int nblocks, int block0,
                                                     int ring_blocks, const msb200_volume_state *__restrict__ peer_states,
                                                     const int *__restrict__ counts) {
	extern __shared__ short vsm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int stream = blockIdx.x * (blockDim.x >> 5) + warp;
	if (stream >= n_streams) return;
	short *buf = vsm + (size_t)warp * ((nsamples + 1) & ~1);
	if (counts) nblocks = min(nblocks, counts[stream]); // ragged batches: this stream staged fewer blocks than the bank's maximum
	// nblocks consecutive blocks (the reference processes one mblk at a time, msvolume.c:505-512), optionally laid out
	// in a block-aligned circular buffer
	for (int blk = 0; blk < nblocks; ++blk) {
	const int bpos = ring_blocks > 0 ? (block0 + blk) % ring_blocks : blk;
	short *g = io + (size_t)stream * stride + (size_t)bpos * nsamples;
	__syncwarp();
	int pk = 0, dcsum = 0;
	for (int i = lane; i < nsamples; i += 32) {
		int s = g[i];
		buf[i] = (short)s;
		int a = s < 0 ? -s : s;
		pk = a > pk ? a : pk;
		dcsum += s;
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		int opk = __shfl_xor_sync(0xffffffffu, pk, o);
		pk = opk > pk ? opk : pk;
		dcsum += __shfl_xor_sync(0xffffffffu, dcsum, o);
	}
	__syncwarp();
	int intgain = 4096, apply = 0, dc_prev = 0, remove_dc = 0;
	if (lane == 0) {
		msb200_volume_state v = st[stream];
		float acc = 0.f;
		// only the float additions are sequential (the reference's accumulation order): loads, squares and conversions
		// of eight samples are issued together so that their latency is paid once per group, not once per sample
		int i = 0;
		for (; i + 8 <= nsamples; i += 8) {
			float q[8];
#pragma unroll
			for (int k = 0; k < 8; ++k) {
				const int s = buf[i + k];
				q[k] = (float)(s * s);
			}
#pragma unroll
			for (int k = 0; k < 8; ++k) acc = __fadd_rn(acc, q[k]);
		}
		for (; i < nsamples; ++i) {
			const int s = buf[i];
			acc = __fadd_rn(acc, (float)(s * s));
		}
		vol_update(v, acc, pk, dcsum, nsamples, peer_states, intgain, apply, remove_dc, dc_prev);
		st[stream] = v;
	}
	intgain = __shfl_sync(0xffffffffu, intgain, 0);
	apply = __shfl_sync(0xffffffffu, apply, 0);
	remove_dc = __shfl_sync(0xffffffffu, remove_dc, 0);
	dc_prev = __shfl_sync(0xffffffffu, dc_prev, 0);
	if (!apply) continue; // gain == 1: the reference leaves the block untouched (:441)
	for (int i = lane; i < nsamples; i += 32) {
		int s = buf[i];
		if (remove_dc) s -= dc_prev;
		g[i] = (short)vol_sat((s * intgain) / 4096); // C truncating division
	}
	}
}

// Large banks: one LANE per stream for the sequential part. A warp owns 32 consecutive streams: it stages their blocks in
// shared memory with coalesced 32-bit loads (rows `pitch` samples apart, pitch / 2 odd: the column walk below is free of
// bank conflicts), every lane then runs the reference's sequential float accumulation, peak and DC sums and the state
// update for ITS stream (state in registers across the blocks of a launch), and the warp applies the 32 gains and stores
// coalesced. Same operations per stream as volume_kernel, 1/32 of its issue slots in the sequential part — that kernel
// spends a whole warp's slots on lane 0 there and is issue-bound from a few thousand streams on.
// VW warps per CTA share the two parallel parts — staging the 32 rows and applying the 32 gains (rows l = warp, warp + VW,
// ...) — around the sequential part, which stays with the lanes of warp 0. With one warp per CTA (the first form of this
// kernel) a 4096-stream bank was 128 lone warps walking ~3500 instructions per block on their own: 21 us per tick of two
// blocks, nearly all of it staging and gain arithmetic that four warps now do side by side.
#define VOL_LANES_WARPS 4
__global__ void __launch_bounds__(32 * VOL_LANES_WARPS) volume_lanes_kernel(short *__restrict__ io, msb200_volume_state *__restrict__ st, int n_streams,
                                                          int nsamples, int stride, int nblocks, int block0, int ring_blocks,
                                                          const msb200_volume_state *__restrict__ peer_states,
                                                          const int *__restrict__ counts, int pitch, short *__restrict__ copy_out,
                                                          int copy_stride) {
	// copy_out (optional): every processed block ALSO goes to copy_out[stream][blk * nsamples ...], rows copy_stride samples
	// apart — the chain's hand-out of a tick's blocks, which was a launch of its own
	extern __shared__ short vsm[];
	__shared__ int sh_gain[32], sh_dc[32]; // per stream of the CTA: the block's integer gain, the DC to remove first
	__shared__ unsigned sh_has;            // streams whose block gets a gain pass at all (bit per stream)
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, s0 = blockIdx.x * 32, stream = s0 + lane;
	const bool valid = stream < n_streams;
	const int my_blocks = valid ? (counts ? min(nblocks, counts[stream]) : nblocks) : 0; // (the same in every warp)
	int max_blocks = my_blocks;
#pragma unroll
	for (int o = 16; o; o >>= 1) max_blocks = max(max_blocks, __shfl_xor_sync(0xffffffffu, max_blocks, o));
	if (max_blocks == 0) return; // uniform over the CTA
	msb200_volume_state v;
	if (warp == 0 && my_blocks > 0) v = st[stream];
	const int words = nsamples >> 1, pw = pitch >> 1;
	unsigned *wsm = reinterpret_cast<unsigned *>(vsm);
	const unsigned wbase = (unsigned)__cvta_generic_to_shared(wsm);
	for (int blk = 0; blk < max_blocks; ++blk) {
		const int bpos = ring_blocks > 0 ? (block0 + blk) % ring_blocks : blk;
		const unsigned has = __ballot_sync(0xffffffffu, blk < my_blocks); // streams that have this block
		__syncthreads(); // the previous block's gain pass has read its rows
		// (asynchronous copies: every word of the rows is in flight at once, nothing waits on a register)
		for (int l = warp; l < 32; l += VOL_LANES_WARPS) {
			if (!((has >> l) & 1u)) continue;
			const unsigned *g = reinterpret_cast<const unsigned *>(io + (size_t)(s0 + l) * stride + (size_t)bpos * nsamples);
			for (int i = lane; i < words; i += 32)
				asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(wbase + 4u * (unsigned)(l * pw + i)), "l"(g + i) : "memory");
		}
		asm volatile("cp.async.wait_all;\n" ::: "memory");
		__syncthreads();
		if (warp == 0) {
			int intgain = 4096, apply = 0, dc_prev = 0, remove_dc = 0;
			if (blk < my_blocks) {
				const short *buf = vsm + (size_t)lane * pitch;
				float acc = 0.f;
				int pk = 0, dcsum = 0, i = 0;
				for (; i + 8 <= nsamples; i += 8) { // (loads and conversions of eight samples in flight together; the adds in order)
					float q[8];
#pragma unroll
					for (int k = 0; k < 8; ++k) {
						const int x = buf[i + k];
						pk = max(pk, x < 0 ? -x : x);
						dcsum += x;
						q[k] = (float)(x * x);
					}
#pragma unroll
					for (int k = 0; k < 8; ++k) acc = __fadd_rn(acc, q[k]);
				}
				for (; i < nsamples; ++i) {
					const int x = buf[i];
					pk = max(pk, x < 0 ? -x : x);
					dcsum += x;
					acc = __fadd_rn(acc, (float)(x * x));
				}
				vol_update(v, acc, pk, dcsum, nsamples, peer_states, intgain, apply, remove_dc, dc_prev);
			}
			sh_gain[lane] = intgain;
			sh_dc[lane] = remove_dc ? dc_prev : 0;
			const unsigned todo = __ballot_sync(0xffffffffu, apply != 0); // gain == 1 and no DC removal: the block stays as it is (:441)
			if (lane == 0) sh_has = todo;
		}
		__syncthreads();
		const unsigned todo = sh_has;
		for (int l = warp; l < 32; l += VOL_LANES_WARPS) {
			const bool ap = (todo >> l) & 1u;
			if (!((has >> l) & 1u) || (!ap && !copy_out)) continue;
			const int ig = sh_gain[l], dcp = sh_dc[l];
			unsigned *g = reinterpret_cast<unsigned *>(io + (size_t)(s0 + l) * stride + (size_t)bpos * nsamples);
			unsigned *c = copy_out ? reinterpret_cast<unsigned *>(copy_out + (size_t)(s0 + l) * copy_stride + (size_t)blk * nsamples) : nullptr;
			for (int i = lane; i < words; i += 32) {
				unsigned w = wsm[l * pw + i];
				if (ap) {
					const int a = vol_sat((((int)(short)(w & 0xffffu) - dcp) * ig) / 4096); // C truncating division
					const int b = vol_sat((((int)(short)(w >> 16) - dcp) * ig) / 4096);
					w = ((unsigned)a & 0xffffu) | ((unsigned)b << 16);
					g[i] = w;
				}
				if (c) c[i] = w;
			}
		}
	}
	if (warp == 0 && my_blocks > 0) st[stream] = v;
}

// stream == -1 applies the setter to every stream of the bank (one round trip)
static int volume_update_state(msb200_volume *v, int stream, void (*fn)(msb200_volume_state *, float, int), float f, int i) {
	MSB200_CHECK_ARG(v && stream >= -1 && stream < v->n);
	const int first = stream < 0 ? 0 : stream, count = stream < 0 ? v->n : 1;
	std::vector<msb200_volume_state> st((size_t)count);
	cudaStream_t s = v->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(st.data(), v->d_state + first, sizeof(msb200_volume_state) * (size_t)count, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	for (auto &x : st) fn(&x, f, i);
	MSB200_CUDA(cudaMemcpyAsync(v->d_state + first, st.data(), sizeof(msb200_volume_state) * (size_t)count, cudaMemcpyHostToDevice, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}

extern "C" {

static void volume_state_init(msb200_volume_state *s, int sample_rate) { // volume_init :86-120
	memset(s, 0, sizeof(*s));
	s->static_gain = s->gain = s->target_gain = 1.f;
	s->ng_threshold = 0.1f;
	s->ng_floorgain = 0.005f;
	s->ng_gain = 1.f;
	s->sample_rate = sample_rate;
	s->ea_thres = 0.1f;          // noise_thres :42
	s->ea_transmit_thres = 4.f;  // transmit_thres :43
	s->force = 4.f;              // en_weight :41
	s->vol_upramp = 0.4f;
	s->sustain_time = 200;
	s->peer = -1;
}
int msb200_volume_reset_stream(msb200_volume *v, int stream) { // a fresh MSVolume in this slot (a stream joining a running bank)
	MSB200_CHECK_ARG(v && stream >= 0 && stream < v->n);
	msb200_volume_state s;
	volume_state_init(&s, v->rate);
	MSB200_CUDA(cudaStreamSynchronize(v->ctx->stream));
	MSB200_CUDA(cudaMemcpy(v->d_state + stream, &s, sizeof(s), cudaMemcpyHostToDevice));
	return MSB200_OK;
}
int msb200_volume_create(msb200_ctx *ctx, int n_streams, int sample_rate, int max_block, msb200_volume **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && sample_rate > 0 && max_block > 0 && max_block <= 8192);
	msb200_volume *v = new msb200_volume();
	v->ctx = ctx;
	v->n = v->live = n_streams;
	v->rate = sample_rate;
	v->max_block = max_block;
	std::vector<msb200_volume_state> init((size_t)n_streams);
	for (auto &s : init) volume_state_init(&s, sample_rate);
	v->peer_bank = nullptr;
	MSB200_CUDA(cudaMalloc(&v->d_state, sizeof(msb200_volume_state) * (size_t)n_streams));
	MSB200_CUDA(cudaMemcpy(v->d_state, init.data(), sizeof(msb200_volume_state) * (size_t)n_streams, cudaMemcpyHostToDevice));
	*out = v;
	return MSB200_OK;
}
void msb200_volume_destroy(msb200_volume *v) {
	if (!v) return;
	cudaStreamSynchronize(v->ctx->stream);
	cudaFree(v->d_state);
	v->io.release();
	delete v;
}
int msb200_volume_set_gain(msb200_volume *v, int stream, float gain) { // volume_set_gain :270-276
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) { s->gain = s->target_gain = s->static_gain = f; }, gain, 0);
}
int msb200_volume_set_db_gain(msb200_volume *v, int stream, float db) { // volume_set_db_gain :262-268 (pow(10, db/10), sic)
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) { s->gain = s->static_gain = (float)pow(10, f / 10); }, db, 0);
}
int msb200_volume_enable_noise_gate(msb200_volume *v, int stream, int enabled) { // :352-359
	return volume_update_state(v, stream, [](msb200_volume_state *s, float, int e) {
		s->noise_gate_enabled = e ? 1 : 0;
		if (s->noise_gate_enabled) s->gain = s->target_gain = s->ng_floorgain;
	}, 0.f, enabled);
}
int msb200_volume_set_noise_gate_threshold(msb200_volume *v, int stream, float thr) {
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) { s->ng_threshold = f; }, thr, 0);
}
int msb200_volume_set_noise_gate_floorgain(msb200_volume *v, int stream, float g) { // :367-378
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) {
		s->ng_floorgain = f < 0.005f ? 0.005f : f;
		if (s->noise_gate_enabled) s->gain = s->target_gain = s->ng_floorgain;
	}, g, 0);
}
int msb200_volume_remove_dc(msb200_volume *v, int stream, int enabled) {
	return volume_update_state(v, stream, [](msb200_volume_state *s, float, int e) { s->remove_dc = e; }, 0.f, enabled);
}
int msb200_volume_enable_agc(msb200_volume *v, int stream, int enabled) {
	return volume_update_state(v, stream, [](msb200_volume_state *s, float, int e) { s->agc_enabled = e ? 1 : 0; }, 0.f, enabled);
}
int msb200_volume_set_peer(msb200_volume *v, int stream, msb200_volume *peer_bank, int peer_stream) {
	MSB200_CHECK_ARG(v && (peer_bank == nullptr || (peer_stream >= 0 && peer_stream < peer_bank->n)));
	MSB200_CHECK_ARG(v->peer_bank == nullptr || peer_bank == nullptr || v->peer_bank == peer_bank); // one peer bank per bank
	if (peer_bank) v->peer_bank = peer_bank;
	else if (v->n == 1) v->peer_bank = nullptr; // the bank's only stream unlinked: nothing may dereference the old peer bank
	return volume_update_state(v, stream, [](msb200_volume_state *s, float, int p) { s->peer = p; }, 0.f, peer_bank ? peer_stream : -1);
}
int msb200_volume_set_ea_threshold(msb200_volume *v, int stream, float thr) {
	MSB200_CHECK_ARG(thr >= 0 && thr <= 1); // "threshold must be in range [0..1]" :308-311
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) { s->ea_thres = f; }, thr, 0);
}
int msb200_volume_set_ea_speed(msb200_volume *v, int stream, float speed) {
	MSB200_CHECK_ARG(speed >= 0 && speed <= .5f); // :327-330
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) { s->vol_upramp = f; }, speed, 0);
}
int msb200_volume_set_ea_force(msb200_volume *v, int stream, float force) {
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) { s->force = f; }, force, 0);
}
int msb200_volume_set_ea_sustain(msb200_volume *v, int stream, int ms) {
	return volume_update_state(v, stream, [](msb200_volume_state *s, float, int i) { s->sustain_time = i; }, 0.f, ms);
}
int msb200_volume_set_ea_transmit_threshold(msb200_volume *v, int stream, float thr) {
	return volume_update_state(v, stream, [](msb200_volume_state *s, float f, int) { s->ea_transmit_thres = f; }, thr, 0);
}
int msb200_volume_get_state(msb200_volume *v, int stream, msb200_volume_state *st) {
	MSB200_CHECK_ARG(v && st && stream >= 0 && stream < v->n);
	MSB200_CUDA(cudaMemcpyAsync(st, v->d_state + stream, sizeof(*st), cudaMemcpyDeviceToHost, v->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(v->ctx->stream));
	return MSB200_OK;
}
int msb200_volume_process_dev(msb200_volume *v, void *d_io, int nsamples, int stride) {
	MSB200_CHECK_ARG(stride >= nsamples);
	return msb200i_volume_launch(v, d_io, nsamples, stride, 1, 0, 0, nullptr, nullptr, 0, nullptr);
}
int msb200_volume_process_blocks(msb200_volume *v, int16_t *io, int nsamples, int stride_samples, int nblocks, const int32_t *counts) {
	MSB200_CHECK_ARG(v && io && nblocks > 0 && stride_samples >= nsamples * nblocks);
	if (v->live == 0) return MSB200_OK;
	// only the staged part of every row crosses PCIe: rows of nblocks*nsamples out of stride_samples
	const size_t pitch = (size_t)stride_samples * 2, row = (size_t)nblocks * nsamples * 2;
	const size_t bytes = ((size_t)v->live * pitch + 15) & ~(size_t)15, cbytes = counts ? sizeof(int32_t) * (size_t)v->live : 0;
	int r = v->io.reserve(bytes + cbytes);
	if (r) return r;
	cudaStream_t s = v->ctx->stream;
	int *d_counts = counts ? (int *)((char *)v->io.p + bytes) : nullptr;
	MSB200_CUDA(cudaMemcpy2DAsync(v->io.p, pitch, io, pitch, row, (size_t)v->live, cudaMemcpyHostToDevice, s));
	if (counts) MSB200_CUDA(cudaMemcpyAsync(d_counts, counts, cbytes, cudaMemcpyHostToDevice, s));
	if ((r = msb200i_volume_launch(v, v->io.p, nsamples, stride_samples, nblocks, 0, 0, d_counts, nullptr, 0, nullptr))) return r;
	MSB200_CUDA(cudaMemcpy2DAsync(io, pitch, v->io.p, pitch, row, (size_t)v->live, cudaMemcpyDeviceToHost, s));
	MSB200_HOST_DONE(v->ctx);
	return MSB200_OK;
}
int msb200_volume_set_kernel(msb200_volume *v, int choice) {
	MSB200_CHECK_ARG(v && choice >= 0 && choice <= 2);
	v->kernel_choice = choice;
	return MSB200_OK;
}
int msb200_volume_set_live(msb200_volume *v, int n_live) {
	MSB200_CHECK_ARG(v && n_live >= 0 && n_live <= v->n);
	v->live = n_live;
	return MSB200_OK;
}
int msb200_volume_process(msb200_volume *v, int16_t *io, int nsamples) {
	MSB200_CHECK_ARG(v && io);
	if (v->live == 0) return MSB200_OK;
	size_t bytes = (size_t)v->live * nsamples * 2;
	int r = v->io.reserve(bytes);
	if (r) return r;
	cudaStream_t s = v->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(v->io.p, io, bytes, cudaMemcpyHostToDevice, s));
	if ((r = msb200_volume_process_dev(v, v->io.p, nsamples, nsamples))) return r;
	MSB200_CUDA(cudaMemcpyAsync(io, v->io.p, bytes, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}

} // extern "C"

int msb200i_volume_launch(msb200_volume *v, void *d_io, int nsamples, int stride, int nblocks, int block0, int ring_blocks,
                          const int *d_counts, void *d_copy_out, int copy_stride, int *copied) {
	MSB200_CHECK_ARG(v && d_io && nsamples > 0 && nsamples <= v->max_block && nblocks > 0);
	if (copied) *copied = 0;
	// one warp per stream, the block staged in shared memory: 8 warps per CTA while that fits the default 48 KB, fewer for
	// long blocks (48 kHz stereo at 40 ms is 3840 samples; max_block goes up to 8192), opting in beyond 48 KB
	int warps = 8;
	const size_t per_warp = (size_t)((nsamples + 1) & ~1) * sizeof(short);
	while (warps > 1 && warps * per_warp > 48 * 1024) warps >>= 1;
	const size_t smem = warps * per_warp;
	if (v->live == 0) return MSB200_OK;
	{ // lane per stream: even blocks on 4-byte boundaries whose 32 rows fit shared memory; by default for banks of 256+ streams
		const int pitch = (nsamples & 3) == 2 ? nsamples : ((nsamples + 3) & ~3) + 2; // even, pitch / 2 odd
		const size_t lsmem = (size_t)32 * pitch * sizeof(short);
		const bool can = (nsamples & 1) == 0 && (stride & 1) == 0 && ((uintptr_t)d_io & 3) == 0 && lsmem <= 96 * 1024;
		if (can && (v->kernel_choice == 2 || (v->kernel_choice == 0 && v->live >= 256))) {
			MSB200_SMEM_OPTIN(volume_lanes_kernel, v->ctx, lsmem);
			const bool cp = d_copy_out && copied && ((uintptr_t)d_copy_out & 3) == 0 && (copy_stride & 1) == 0 && copy_stride >= nblocks * nsamples;
			MSB200_LAUNCH(v->ctx, volume_lanes_kernel, msb200_div_up(v->live, 32), 32 * VOL_LANES_WARPS, lsmem, (short *)d_io, v->d_state, v->live, nsamples,
			              stride, nblocks, block0, ring_blocks,
			              (const msb200_volume_state *)(v->peer_bank ? v->peer_bank->d_state : nullptr), d_counts, pitch,
			              cp ? (short *)d_copy_out : (short *)nullptr, copy_stride);
			if (cp) *copied = 1;
			return MSB200_OK;
		}
	}
	MSB200_SMEM_OPTIN(volume_kernel, v->ctx, smem);
	MSB200_LAUNCH(v->ctx, volume_kernel, msb200_div_up(v->live, warps), warps * 32, smem, (short *)d_io, v->d_state, v->live,
	              nsamples, stride, nblocks, block0, ring_blocks,
	              (const msb200_volume_state *)(v->peer_bank ? v->peer_bank->d_state : nullptr), d_counts);
	return MSB200_OK;
}

// ======================================================================================================= channel adapter
// reference: /root/reference/src/audiofilters/chanadapt.c:68-131

__global__ void chanadapt_kernel(const short *__restrict__ in, const short *__restrict__ in2, short *__restrict__ out,
                                 long total, int mode) {
	long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
	const long step = (long)gridDim.x * blockDim.x;
	for (; i < total; i += step) {
		if (mode == MSB200_CHAN_MONO_TO_STEREO) { // :112-120
			int s = (unsigned short)in[i];
			reinterpret_cast<int *>(out)[i] = s | (s << 16);
		} else if (mode == MSB200_CHAN_STEREO_TO_MONO) { // :121-129 (left channel)
			out[i] = (short)(reinterpret_cast<const int *>(in)[i] & 0xffff);
		} else { // :78-96 (absent pin -> zeros)
			int l = in ? (unsigned short)in[i] : 0, r = in2 ? (unsigned short)in2[i] : 0;
			reinterpret_cast<int *>(out)[i] = l | (r << 16);
		}
	}
}

extern "C" {
int msb200_chanadapt_process_dev(msb200_ctx *ctx, int mode, int n_streams, int frames, const void *d_in,
                                 const void *d_in2, void *d_out) {
	MSB200_CHECK_ARG(ctx && d_out && n_streams > 0 && frames > 0 && mode >= 0 && mode <= 2);
	MSB200_CHECK_ARG(mode == MSB200_CHAN_2MONO_TO_STEREO || d_in != nullptr);
	long total = (long)n_streams * frames;
	int block = 256;
	long want = (total + block - 1) / block;
	int grid = (int)(want < (long)ctx->sm_count * 16 ? want : (long)ctx->sm_count * 16);
	MSB200_LAUNCH(ctx, chanadapt_kernel, grid, block, 0, (const short *)d_in, (const short *)d_in2, (short *)d_out, total, mode);
	return MSB200_OK;
}
int msb200_chanadapt_process(msb200_ctx *ctx, int mode, int n_streams, int frames, const int16_t *in, const int16_t *in2,
                             int16_t *out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && frames > 0 && mode >= 0 && mode <= 2);
	size_t mono = (size_t)n_streams * frames * 2;
	size_t in_bytes = mode == MSB200_CHAN_STEREO_TO_MONO ? mono * 2 : mono;
	size_t out_bytes = mode == MSB200_CHAN_STEREO_TO_MONO ? mono : mono * 2;
	void *d_in = nullptr, *d_in2 = nullptr, *d_out = nullptr;
	cudaStream_t s = ctx->stream;
	MSB200_CUDA(cudaMalloc(&d_out, out_bytes));
	if (in) {
		MSB200_CUDA(cudaMalloc(&d_in, in_bytes));
		MSB200_CUDA(cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, s));
	}
	if (in2 && mode == MSB200_CHAN_2MONO_TO_STEREO) {
		MSB200_CUDA(cudaMalloc(&d_in2, in_bytes));
		MSB200_CUDA(cudaMemcpyAsync(d_in2, in2, in_bytes, cudaMemcpyHostToDevice, s));
	}
	int r = msb200_chanadapt_process_dev(ctx, mode, n_streams, frames, d_in, d_in2, d_out);
	if (!r) {
		cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s);
		cudaStreamSynchronize(s);
	}
	cudaFree(d_in);
	cudaFree(d_in2);
	cudaFree(d_out);
	return r;
}
} // extern "C"

// ======================================================================================================= equalizer
// reference: /root/reference/src/audiofilters/equalizer.c:263-288, /root/reference/src/utils/dsptools.c:253-268

struct msb200_equalizer {
	msb200_ctx *ctx;
	int n, rate, nfft, max_block;
	float *d_taps;      // [stream][nfft]
	float *d_hist;      // [stream][nfft-1] previous inputs, oldest first
	uint8_t *d_active;  // [stream]
	std::vector<std::vector<float>> fft_cpx; // host gain tables, packed real spectrum (equalizer.c:49-55)
	msb200_devbuf io;
};

// x86 semantics of the C cast (int16_t)float in word16_to_int16 (equalizer.c:251-255): truncate to int32
// (out-of-range -> 0x80000000), keep the low 16 bits.
__device__ __forceinline__ short c_cast_f32_s16(float f) {
	int i;
	if (!(f > -2147483904.0f && f < 2147483648.0f)) i = (int)0x80000000u;
	else i = __float2int_rz(f);
	return (short)(i & 0xffff);
}

// One CTA per stream, one thread per output sample. x[] = history (ord-1) ++ block in shared memory as float.
// y[i] = sum_{j=ord-1..0} num[j]*x[i-j], accumulated in exactly that order with separate multiply and add.
__global__ void __launch_bounds__(512) eq_fir_kernel(short *__restrict__ io, const float *__restrict__ taps,
                                                     float *__restrict__ hist, const uint8_t *__restrict__ active,
                                                     int nsamples, int stride, int ord) {
	extern __shared__ float esm[];
	float *num = esm;          // [ord]
	float *x = esm + ord;      // [ord-1 + nsamples]
	const int stream = blockIdx.x;
	if (!active[stream]) return;
	short *g = io + (size_t)stream * stride;
	float *h = hist + (size_t)stream * (ord - 1);
	for (int i = threadIdx.x; i < ord; i += blockDim.x) num[i] = taps[(size_t)stream * ord + i];
	for (int i = threadIdx.x; i < ord - 1; i += blockDim.x) x[i] = h[i];
	for (int i = threadIdx.x; i < nsamples; i += blockDim.x) x[ord - 1 + i] = (float)g[i];
	__syncthreads();
	for (int i = threadIdx.x; i < nsamples; i += blockDim.x) {
		const float *xi = x + i; // xi[ord-1-j] == x[i-j] in block coordinates
		float acc = __fmul_rn(xi[0], num[ord - 1]);
#pragma unroll 8
		for (int j = ord - 2; j >= 0; --j) acc = __fadd_rn(acc, __fmul_rn(num[j], xi[ord - 1 - j]));
		g[i] = c_cast_f32_s16(acc);
	}
	// new history = last ord-1 inputs
	for (int i = threadIdx.x; i < ord - 1; i += blockDim.x) h[i] = x[nsamples + i];
}

static void eq_flatten(std::vector<float> &t, int nfft) { // equalizer_state_flatten :49-55
	t.assign((size_t)nfft, 0.f);
	float val = 1.0f / (float)nfft;
	t[0] = val;
	for (int i = 1; i < nfft; i += 2) t[(size_t)i] = val;
}
static int eq_hz_to_index(int rate, int nfft, int hz) { // :98-111
	if (hz < 0) return -1;
	if (hz > rate / 2) hz = rate / 2;
	int ret = ((hz * nfft) + (rate / 2)) / rate;
	if (ret == nfft / 2) ret = (nfft / 2) - 1;
	return ret;
}
static int eq_index2hz(int rate, int nfft, int index) { // :113-115
	return (index * rate + nfft / 2) / nfft;
}
static float eq_gainpoint(int f, int freq_0, float sqrt_gain, int freq_bw) { // :131-138
	float k1 = ((float)(f * f) - (float)(freq_0 * freq_0));
	k1 *= k1;
	float k2 = (float)(f * freq_bw);
	k2 *= k2;
	return (k1 + k2 * sqrt_gain) / (k1 + k2 / sqrt_gain);
}
static void eq_point_set(std::vector<float> &t, int nfft, int i, float gain) { // :140-148
	int index = 1 + ((i - 1) * 2);
	if (index >= 0 && index < nfft) t[(size_t)index] = (t[(size_t)index] * (float)(int)(gain * 32768)) / 32768;
}
// ms_ifft (dsptools.c:373-376) for the equalizer's sizes (128 / 256 / 512 points): kiss_fftri2 (kiss_fftr.c:261-296) over the
// float kiss_fft (kiss_fft.c: digit-reversal gather, then radix-2 / radix-4 passes from the innermost level out), every
// float operation in the reference's order, twiddles from the same double-precision expressions, so that the taps - and with
// the FIR's summation order (eq_fir_kernel) every output sample - equal the reference's bit for bit. Host code.
namespace {
struct EqCpx {
	float r, i;
};
inline EqCpx eq_cmul(EqCpx a, EqCpx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
inline EqCpx eq_cadd(EqCpx a, EqCpx b) { return {a.r + b.r, a.i + b.i}; }
inline EqCpx eq_csub(EqCpx a, EqCpx b) { return {a.r - b.r, a.i - b.i}; }
bool eq_kiss_irfft(const float *spec, float *out, int nfft) {
	const double pi = 3.14159265358979323846264338327;
	const int n = nfft / 2;
	std::vector<EqCpx> tw((size_t)n), super((size_t)n), tmp((size_t)n), F((size_t)n);
	for (int i = 0; i < n; ++i) {
		const double phase = -((-2 * pi / n) * i); // inverse transform
		tw[(size_t)i] = {(float)cos(phase), (float)sin(phase)};
		const double ph2 = pi * (((double)i) / n + .5);
		super[(size_t)i] = {(float)cos(ph2), (float)sin(ph2)};
	}
	int p[16], m[16], st[16], nf = 0, rem = n, stride = 1, radix = 4; // kf_factor
	do {
		while (rem % radix) {
			radix = radix == 4 ? 2 : (radix == 2 ? 3 : radix + 2);
			if (radix > 32000 || radix * radix > rem) radix = rem;
		}
		rem /= radix;
		if ((radix != 4 && radix != 2) || nf >= 16) return false;
		p[nf] = radix;
		m[nf] = rem;
		st[nf] = stride;
		stride *= radix;
		++nf;
	} while (rem > 1);
	tmp[0] = {spec[0] + spec[2 * n - 1], spec[0] - spec[2 * n - 1]};
	for (int k = 1; k <= n / 2; ++k) {
		const EqCpx fk = {spec[2 * k - 1], spec[2 * k]}, fnkc = {spec[2 * (n - k) - 1], -spec[2 * (n - k)]};
		const EqCpx fek = eq_cadd(fk, fnkc), fok = eq_cmul(eq_csub(fk, fnkc), super[(size_t)k]);
		tmp[(size_t)k] = eq_cadd(fek, fok);
		EqCpx r = eq_csub(fek, fok);
		r.i *= -1;
		tmp[(size_t)(n - k)] = r;
	}
	for (int o = 0; o < n; ++o) { // kf_shuffle
		int src = 0;
		for (int d = 0; d < nf; ++d) src += ((o / m[d]) % p[d]) * st[d];
		F[(size_t)o] = tmp[(size_t)src];
	}
	for (int d = nf - 1; d >= 0; --d) {
		const int mm = m[d], s = st[d];
		for (int b = 0; b < s * mm; ++b) {
			const int j = b % mm;
			EqCpx *f = F.data() + (size_t)(b / mm) * p[d] * mm + j;
			if (p[d] == 2) { // kf_bfly2
				const EqCpx t = eq_cmul(f[mm], tw[(size_t)(j * s)]);
				f[mm] = eq_csub(f[0], t);
				f[0] = eq_cadd(f[0], t);
			} else { // kf_bfly4, inverse
				const EqCpx s0 = eq_cmul(f[mm], tw[(size_t)(j * s)]), s1 = eq_cmul(f[2 * mm], tw[(size_t)(2 * j * s)]),
				            s2 = eq_cmul(f[3 * mm], tw[(size_t)(3 * j * s)]);
				const EqCpx s5 = eq_csub(f[0], s1);
				f[0] = eq_cadd(f[0], s1);
				const EqCpx s3 = eq_cadd(s0, s2), s4 = eq_csub(s0, s2);
				f[2 * mm] = eq_csub(f[0], s3);
				f[0] = eq_cadd(f[0], s3);
				f[mm] = {s5.r - s4.i, s5.i + s4.r};
				f[3 * mm] = {s5.r + s4.i, s5.i - s4.r};
			}
		}
	}
	for (int o = 0; o < n; ++o) {
		out[2 * o] = F[(size_t)o].r;
		out[2 * o + 1] = F[(size_t)o].i;
	}
	return true;
}
} // namespace
// gain table -> impulse response: ms_ifft + time_shift (equalizer.c:184-193) + Hamming (:203-213)
static void eq_design(const std::vector<float> &spec, int n, std::vector<float> &fir) {
	fir.assign((size_t)n, 0.f);
	const int half = n / 2;
	std::vector<float> t((size_t)n, 0.f);
	eq_kiss_irfft(spec.data(), t.data(), n);
	for (int i = 0; i < n; ++i) fir[(size_t)((i + half) % n)] = t[(size_t)i]; // time shift: swap halves
	for (int i = 0; i < n; ++i) {
		float x = (float)((float)i * 2 * M_PI / (float)n);
		float w = (float)(0.54 - (0.46 * cos((double)x))); // C's cos(): double (a bare cos(float) would be cosf in C++)
		fir[(size_t)i] = w * fir[(size_t)i];
	}
}
static int eq_upload_taps(msb200_equalizer *e, int stream) {
	std::vector<float> fir;
	eq_design(e->fft_cpx[(size_t)stream], e->nfft, fir);
	MSB200_CUDA(cudaMemcpyAsync(e->d_taps + (size_t)stream * e->nfft, fir.data(), sizeof(float) * (size_t)e->nfft,
	                            cudaMemcpyHostToDevice, e->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(e->ctx->stream));
	return MSB200_OK;
}

extern "C" {

int msb200_equalizer_create(msb200_ctx *ctx, int n_streams, int sample_rate, int max_block, msb200_equalizer **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && sample_rate > 0 && max_block > 0 && max_block <= 8192);
	msb200_equalizer *e = new msb200_equalizer();
	e->ctx = ctx;
	e->n = n_streams;
	e->rate = sample_rate;
	e->nfft = sample_rate < 16000 ? 128 : (sample_rate < 32000 ? 256 : 512); // equalizer_rate_update :57-79
	e->max_block = max_block;
	e->fft_cpx.resize((size_t)n_streams);
	MSB200_CUDA(cudaMalloc(&e->d_taps, sizeof(float) * (size_t)n_streams * e->nfft));
	MSB200_CUDA(cudaMalloc(&e->d_hist, sizeof(float) * (size_t)n_streams * (e->nfft - 1)));
	MSB200_CUDA(cudaMalloc(&e->d_active, (size_t)n_streams));
	MSB200_CUDA(cudaMemset(e->d_hist, 0, sizeof(float) * (size_t)n_streams * (e->nfft - 1)));
	MSB200_CUDA(cudaMemset(e->d_active, 1, (size_t)n_streams));
	std::vector<float> flat, fir, all((size_t)n_streams * e->nfft);
	eq_flatten(flat, e->nfft);
	eq_design(flat, e->nfft, fir);
	for (int s = 0; s < n_streams; ++s) {
		e->fft_cpx[(size_t)s] = flat;
		memcpy(&all[(size_t)s * e->nfft], fir.data(), sizeof(float) * (size_t)e->nfft);
	}
	MSB200_CUDA(cudaMemcpy(e->d_taps, all.data(), sizeof(float) * all.size(), cudaMemcpyHostToDevice));
	*out = e;
	return MSB200_OK;
}
void msb200_equalizer_destroy(msb200_equalizer *e) {
	if (!e) return;
	cudaStreamSynchronize(e->ctx->stream);
	cudaFree(e->d_taps);
	cudaFree(e->d_hist);
	cudaFree(e->d_active);
	e->io.release();
	delete e;
}
int msb200_equalizer_nfft(msb200_equalizer *e) {
	return e ? e->nfft : MSB200_EINVAL;
}
int msb200_equalizer_set_gain(msb200_equalizer *e, int stream, float frequency, float gain, float width) {
	MSB200_CHECK_ARG(e && stream >= 0 && stream < e->n);
	// equalizer_state_set :150-177
	std::vector<float> &t = e->fft_cpx[(size_t)stream];
	const int rate = e->rate, nfft = e->nfft;
	int freq_0 = (int)frequency, freq_bw = (int)width, i, f;
	int delta_f = eq_index2hz(rate, nfft, 1);
	float sqrt_gain = (float)sqrt(gain);
	int mid = eq_hz_to_index(rate, nfft, freq_0);
	MSB200_CHECK_ARG(mid >= 0);
	freq_bw -= delta_f / 2;
	if (freq_bw < delta_f / 2) freq_bw = delta_f / 2;
	i = mid;
	eq_point_set(t, nfft, i, gain);
	do {
		i++;
		f = eq_index2hz(rate, nfft, i);
		gain = eq_gainpoint(f - delta_f, freq_0, sqrt_gain, freq_bw);
		eq_point_set(t, nfft, i, gain);
	} while (i < nfft / 2 && (gain > 1.1 || gain < 0.9));
	i = mid;
	do {
		i--;
		f = eq_index2hz(rate, nfft, i);
		gain = eq_gainpoint(f + delta_f, freq_0, sqrt_gain, freq_bw);
		eq_point_set(t, nfft, i, gain);
	} while (i >= 0 && (gain > 1.1 || gain < 0.9));
	return eq_upload_taps(e, stream);
}
int msb200_equalizer_get_gain(msb200_equalizer *e, int stream, float frequency, float *gain) { // :121-125
	MSB200_CHECK_ARG(e && gain && stream >= 0 && stream < e->n);
	int idx = eq_hz_to_index(e->rate, e->nfft, (int)frequency);
	*gain = idx >= 0 ? e->fft_cpx[(size_t)stream][(size_t)idx * 2] * (float)e->nfft : 0.f;
	return MSB200_OK;
}
int msb200_equalizer_set_active(msb200_equalizer *e, int stream, int active) {
	MSB200_CHECK_ARG(e && stream >= 0 && stream < e->n);
	uint8_t a = active ? 1 : 0;
	MSB200_CUDA(cudaMemcpyAsync(e->d_active + stream, &a, 1, cudaMemcpyHostToDevice, e->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(e->ctx->stream));
	return MSB200_OK;
}
int msb200_equalizer_design(int nfft, const float *gain_table, float *taps) { // pure host code: no device needed
	MSB200_CHECK_ARG(gain_table && taps && (nfft == 128 || nfft == 256 || nfft == 512));
	std::vector<float> spec(gain_table, gain_table + nfft), fir;
	eq_design(spec, nfft, fir);
	memcpy(taps, fir.data(), sizeof(float) * (size_t)nfft);
	return MSB200_OK;
}

int msb200_equalizer_set_taps(msb200_equalizer *e, int stream, const float *taps) {
	MSB200_CHECK_ARG(e && taps && stream >= 0 && stream < e->n);
	MSB200_CUDA(cudaMemcpyAsync(e->d_taps + (size_t)stream * e->nfft, taps, sizeof(float) * (size_t)e->nfft,
	                            cudaMemcpyHostToDevice, e->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(e->ctx->stream));
	return MSB200_OK;
}
int msb200_equalizer_get_taps(msb200_equalizer *e, int stream, float *taps) {
	MSB200_CHECK_ARG(e && taps && stream >= 0 && stream < e->n);
	MSB200_CUDA(cudaMemcpyAsync(taps, e->d_taps + (size_t)stream * e->nfft, sizeof(float) * (size_t)e->nfft,
	                            cudaMemcpyDeviceToHost, e->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(e->ctx->stream));
	return MSB200_OK;
}
int msb200_equalizer_process_dev(msb200_equalizer *e, void *d_io, int nsamples, int stride) {
	MSB200_CHECK_ARG(e && d_io && nsamples > 0 && nsamples <= e->max_block && stride >= nsamples);
	size_t smem = sizeof(float) * (size_t)(e->nfft + e->nfft - 1 + nsamples);
	int block = nsamples >= 512 ? 512 : ((nsamples + 31) & ~31);
	MSB200_SMEM_OPTIN(eq_fir_kernel, e->ctx, smem);
	MSB200_LAUNCH(e->ctx, eq_fir_kernel, e->n, block, smem, (short *)d_io, e->d_taps, e->d_hist, e->d_active, nsamples,
	              stride, e->nfft);
	return MSB200_OK;
}
int msb200_equalizer_process(msb200_equalizer *e, int16_t *io, int nsamples) {
	MSB200_CHECK_ARG(e && io);
	size_t bytes = (size_t)e->n * nsamples * 2;
	int r = e->io.reserve(bytes);
	if (r) return r;
	cudaStream_t s = e->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(e->io.p, io, bytes, cudaMemcpyHostToDevice, s));
	if ((r = msb200_equalizer_process_dev(e, e->io.p, nsamples, nsamples))) return r;
	MSB200_CUDA(cudaMemcpyAsync(io, e->io.p, bytes, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}

} // extern "C"

// ======================================================================================================= flow control
// MSAudioFlowControl: ms_audio_flow_controller_process() /root/reference/src/audiofilters/flowcontrol.c
// :110-150 — while a drop target is armed, every block either passes, is dropped whole (basic strategy; almost silent
// frame; too many samples to delete) or loses `todrop` samples, each taken from the middle of the flattest three-sample
// run (discard_well_choosed_samples :58-92; ties go to the LAST position). Integer work, bit-exact; the frame power is
// the reference's sequential float32 sum (one thread, additions in order). One CTA per stream: the block lives in
// shared memory, every deletion is a block-wide arg-min followed by a shifted copy between two buffers.
struct msb200_flowcontrol {
	msb200_ctx *ctx;
	int n, max_block;
	msb200_flowcontrol_state *d_state;
	int *d_out_n;
	msb200_devbuf io;
};

__global__ void __launch_bounds__(128) flowctl_kernel(short *__restrict__ io, int stride, int nsamples,
                                                      msb200_flowcontrol_state *__restrict__ st, int *__restrict__ out_n, int n_streams) {
	extern __shared__ short fsm[];
	short *cur = fsm, *nxt = fsm + ((nsamples + 1) & ~1);
	__shared__ int sh_mode, sh_todrop, sh_val[4], sh_pos[4];
	const int stream = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
	if (stream >= n_streams) return;
	short *g = io + (size_t)stream * stride;
	for (int i = t; i < nsamples; i += blockDim.x) cur[i] = g[i];
	__syncthreads();
	if (t == 0) {
		msb200_flowcontrol_state c = st[stream];
		int mode = 0; // 0 pass, 1 drop the block, 2 delete sh_todrop samples
		unsigned todrop = 0;
		const unsigned n = (unsigned)nsamples;
		if (c.total_samples > 0 && c.target_samples > 0) {
			c.current_pos += n;
			if (c.strategy == 0) {
				if (c.current_dropped + n <= c.target_samples) {
					c.current_dropped += n;
					mode = 1;
				}
			} else {
				const unsigned th = (unsigned)(((unsigned long long)c.target_samples * (unsigned long long)c.current_pos) /
				                               (unsigned long long)c.total_samples);
				todrop = th > c.current_dropped ? th - c.current_dropped : 0;
				if (todrop > 0) {
					bool silent = false;
					if (n <= c.target_samples) { // compute_frame_power :100-108
						float acc = 0.f;
						for (unsigned i = 0; i < n; ++i) {
							const int v = cur[i];
							acc = __fadd_rn(acc, (float)(v * v));
						}
						silent = __fdiv_rn(sqrtf(__fdiv_rn(acc, (float)n)), 32768 * 0.7f) < c.silent_threshold;
					}
					if (silent || !(todrop * 8 < n)) {
						todrop = n;
						mode = 1;
					} else {
						mode = 2;
					}
					c.current_dropped += todrop;
				}
			}
			if (c.current_pos >= c.total_samples) c.target_samples = 0;
			st[stream] = c;
		}
		sh_mode = mode;
		sh_todrop = (int)todrop;
	}
	__syncthreads();
	const int mode = sh_mode;
	int n = nsamples;
	if (mode == 1) {
		if (t == 0) out_n[stream] = 0;
		return;
	}
	if (mode == 2) {
		for (int d = 0; d < sh_todrop; ++d) {
			// arg-min of |s[i]-s[i+1]| + |s[i+1]-s[i+2]| over i < n-2, the LAST position among equal minima
			int best = 32768 * 4, pos = 0;
			for (int i = t; i + 2 < n; i += blockDim.x) {
				const int a = cur[i], b = cur[i + 1], c2 = cur[i + 2];
				const int v = abs(a - b) + abs(b - c2);
				if (v <= best) {
					best = v;
					pos = i;
				}
			}
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, pos, o);
				if (ob < best || (ob == best && op > pos)) {
					best = ob;
					pos = op;
				}
			}
			if (lane == 0) {
				sh_val[warp] = best;
				sh_pos[warp] = pos;
			}
			__syncthreads();
			best = sh_val[0];
			pos = sh_pos[0];
#pragma unroll
			for (int w = 1; w < 4; ++w)
				if (sh_val[w] < best || (sh_val[w] == best && sh_pos[w] > pos)) {
					best = sh_val[w];
					pos = sh_pos[w];
				}
			// the reference starts from min_diff = 32768 with `<=`: a run whose measure exceeds it is never selected and
			// position 0 is removed instead (cannot happen for 16-bit samples: the measure is <= 2 * 65535... it can)
			if (best > 32768) pos = 0;
			for (int i = t; i < n - 1; i += blockDim.x) nxt[i] = cur[i <= pos ? i : i + 1];
			--n;
			__syncthreads();
			short *sw = cur;
			cur = nxt;
			nxt = sw;
		}
		for (int i = t; i < n; i += blockDim.x) g[i] = cur[i];
	}
	if (t == 0) out_n[stream] = n;
}

extern "C" {

static void flowctl_state_init(msb200_flowcontrol_state *s) { // ms_audio_flow_controller_init :37-41
	memset(s, 0, sizeof(*s));
	s->strategy = MSB200_FLOWCONTROL_SOFT;
	s->silent_threshold = 0.02f;
}
int msb200_flowcontrol_create(msb200_ctx *ctx, int n_streams, int max_block, msb200_flowcontrol **out) {
	MSB200_CHECK_ARG(ctx && out && n_streams > 0 && max_block >= 3 && max_block <= 8192);
	msb200_flowcontrol *f = new msb200_flowcontrol();
	f->ctx = ctx;
	f->n = n_streams;
	f->max_block = max_block;
	std::vector<msb200_flowcontrol_state> init((size_t)n_streams);
	for (auto &s : init) flowctl_state_init(&s);
	MSB200_CUDA(cudaMalloc(&f->d_state, sizeof(msb200_flowcontrol_state) * (size_t)n_streams));
	MSB200_CUDA(cudaMalloc(&f->d_out_n, sizeof(int) * (size_t)n_streams));
	MSB200_CUDA(cudaMemcpy(f->d_state, init.data(), sizeof(msb200_flowcontrol_state) * (size_t)n_streams, cudaMemcpyHostToDevice));
	*out = f;
	return MSB200_OK;
}
void msb200_flowcontrol_destroy(msb200_flowcontrol *f) {
	if (!f) return;
	cudaStreamSynchronize(f->ctx->stream);
	cudaFree(f->d_state);
	cudaFree(f->d_out_n);
	f->io.release();
	delete f;
}
int msb200_flowcontrol_get_state(msb200_flowcontrol *f, int stream, msb200_flowcontrol_state *st) {
	MSB200_CHECK_ARG(f && st && stream >= 0 && stream < f->n);
	MSB200_CUDA(cudaMemcpyAsync(st, f->d_state + stream, sizeof(*st), cudaMemcpyDeviceToHost, f->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(f->ctx->stream));
	return MSB200_OK;
}
static int flowctl_put_state(msb200_flowcontrol *f, int stream, const msb200_flowcontrol_state *st) {
	MSB200_CUDA(cudaMemcpyAsync(f->d_state + stream, st, sizeof(*st), cudaMemcpyHostToDevice, f->ctx->stream));
	MSB200_CUDA(cudaStreamSynchronize(f->ctx->stream));
	return MSB200_OK;
}
int msb200_flowcontrol_set_config(msb200_flowcontrol *f, int stream, int strategy, float silent_threshold) {
	MSB200_CHECK_ARG(f && stream >= 0 && stream < f->n && (strategy == MSB200_FLOWCONTROL_BASIC || strategy == MSB200_FLOWCONTROL_SOFT));
	msb200_flowcontrol_state st;
	int r = msb200_flowcontrol_get_state(f, stream, &st);
	if (r) return r;
	st.strategy = strategy;
	st.silent_threshold = silent_threshold;
	return flowctl_put_state(f, stream, &st);
}
int msb200_flowcontrol_set_target(msb200_flowcontrol *f, int stream, uint32_t samples_to_drop, uint32_t total_samples) {
	MSB200_CHECK_ARG(f && stream >= 0 && stream < f->n);
	msb200_flowcontrol_state st;
	int r = msb200_flowcontrol_get_state(f, stream, &st);
	if (r) return r;
	st.target_samples = samples_to_drop; // ms_audio_flow_controller_set_target :51-56
	st.total_samples = total_samples;
	st.current_pos = 0;
	st.current_dropped = 0;
	return flowctl_put_state(f, stream, &st);
}
int msb200_flowcontrol_reset(msb200_flowcontrol *f, int stream) { // ms_audio_flow_controller_reset :30-35 (the configuration stays)
	MSB200_CHECK_ARG(f && stream >= 0 && stream < f->n);
	msb200_flowcontrol_state st;
	int r = msb200_flowcontrol_get_state(f, stream, &st);
	if (r) return r;
	st.target_samples = st.total_samples = st.current_pos = st.current_dropped = 0;
	return flowctl_put_state(f, stream, &st);
}
int msb200_flowcontrol_process_dev(msb200_flowcontrol *f, void *d_io, int nsamples, int stride, void *d_out_nsamples) {
	MSB200_CHECK_ARG(f && d_io && d_out_nsamples && nsamples >= 3 && nsamples <= f->max_block && stride >= nsamples);
	const size_t smem = 2 * sizeof(short) * (size_t)((nsamples + 1) & ~1);
	MSB200_LAUNCH(f->ctx, flowctl_kernel, f->n, 128, smem, (short *)d_io, stride, nsamples, f->d_state, (int *)d_out_nsamples, f->n);
	return MSB200_OK;
}
int msb200_flowcontrol_process(msb200_flowcontrol *f, int16_t *io, int nsamples, int32_t *out_nsamples) {
	MSB200_CHECK_ARG(f && io && out_nsamples);
	const size_t bytes = (size_t)f->n * nsamples * 2;
	int r = f->io.reserve(bytes);
	if (r) return r;
	cudaStream_t s = f->ctx->stream;
	MSB200_CUDA(cudaMemcpyAsync(f->io.p, io, bytes, cudaMemcpyHostToDevice, s));
	if ((r = msb200_flowcontrol_process_dev(f, f->io.p, nsamples, nsamples, f->d_out_n))) return r;
	MSB200_CUDA(cudaMemcpyAsync(io, f->io.p, bytes, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaMemcpyAsync(out_nsamples, f->d_out_n, sizeof(int) * (size_t)f->n, cudaMemcpyDeviceToHost, s));
	MSB200_CUDA(cudaStreamSynchronize(s));
	return MSB200_OK;
}

} // extern "C"
