"""GPU parity: the CUDA banks called through the C ABI vs the CPU oracle on the same seeded inputs. Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
from _oracle import OrcVolumeState, ptr
from mediastreamer2_b200 import _lib
from mediastreamer2_b200 import filters as F

pytestmark = pytest.mark.gpu


def noise(seed, shape, amp):
    return np.random.default_rng(seed).integers(-amp, amp + 1, size=shape).astype(np.int16)


# ------------------------------------------------------------------------------------------------ mixer
@pytest.mark.parametrize("rooms,pins,nwords,conf", [(3, 16, 480, True), (5, 4, 160, True), (2, 50, 480, False),
                                                    (7, 3, 441, True), (4, 16, 482, True), (1024, 16, 480, True)])
def test_mixer_bit_exact(ctx, rooms, pins, nwords, conf):
    L = O.oracle()
    pcm = noise(1, (rooms, pins, nwords), 6000)
    pcm[:, :, ::97] = 30000  # force saturation (BASELINE cfg3 pattern)
    pcm[0, 0, :8] = -32768
    rng = np.random.default_rng(2)
    gain = np.ones((rooms, pins), np.float32)
    active = np.ones((rooms, pins), np.uint8)
    present = (rng.random((rooms, pins)) > 0.1).astype(np.uint8)
    m = F.AudioMixer(ctx, rooms, pins, nwords, conf)
    for r in range(min(rooms, 8)):
        gain[r, min(3, pins - 1)] = 0.5
        m.set_input_gain(r, min(3, pins - 1), 0.5)
        gain[r, 0] = 1.7
        m.set_input_gain(r, 0, 1.7)
        if pins > 2:
            active[r, 2] = 0
            m.set_active(r, 2, False)
    got = m.process(pcm, present)
    exp = np.zeros_like(got)
    L.orc_mixer_process(rooms, pins, nwords, int(conf), ptr(gain), ptr(active), ptr(pcm), ptr(present), ptr(exp))
    m.close()
    assert np.array_equal(got, exp)


def test_mixer_two_phase_equals_single_pass(ctx):
    """partial sums (the thing NCCL all-reduces) + finish == single-pass kernel == oracle."""
    L = O.oracle()
    lib = ctx.lib
    rooms, pins, nwords = 6, 16, 480
    pcm = noise(9, (rooms, pins, nwords), 9000)
    present = np.ones((rooms, pins), np.uint8)
    m = F.AudioMixer(ctx, rooms, pins, nwords, True)
    d_in, d_pr = ctx.dev_alloc(pcm.nbytes), ctx.dev_alloc(present.nbytes)
    d_sum, d_out = ctx.dev_alloc(rooms * nwords * 4), ctx.dev_alloc(pcm.nbytes)
    ctx.h2d(d_in, pcm)
    ctx.h2d(d_pr, present)
    _lib.check(lib.msb200_mixer_partial_dev(m.h, d_in, d_pr, d_sum))
    sums = np.zeros((rooms, nwords), np.int32)
    ctx.d2h(sums, d_sum)
    _lib.check(lib.msb200_mixer_finish_dev(m.h, d_in, d_pr, d_sum, d_out))
    out = np.zeros_like(pcm)
    ctx.d2h(out, d_out)
    gain, active = np.ones((rooms, pins), np.float32), np.ones((rooms, pins), np.uint8)
    esum = np.zeros((rooms, nwords), np.int32)
    L.orc_mixer_partial(rooms, pins, nwords, ptr(gain), ptr(active), ptr(pcm), ptr(present), ptr(esum))
    exp = np.zeros_like(pcm)
    L.orc_mixer_process(rooms, pins, nwords, 1, ptr(gain), ptr(active), ptr(pcm), ptr(present), ptr(exp))
    for p in (d_in, d_pr, d_sum, d_out):
        ctx.dev_free(p)
    m.close()
    assert np.array_equal(sums, esum)
    assert np.array_equal(out, exp)


# ------------------------------------------------------------------------------------------------ volume
@pytest.mark.parametrize("kernel", [1, 2])  # one warp per stream / one lane per stream (what banks of 256+ streams run)
@pytest.mark.parametrize("nsamples", [480, 256, 160, 441, 322])
def test_volume_bit_exact_including_float_state(ctx, nsamples, kernel):
    L = O.oracle()
    n, T, rate = 37, 30, 48000
    v = F.Volume(ctx, n, rate)
    v.set_kernel(kernel)  # (odd block lengths stay on the warp kernel either way)
    cfgs = []
    for s in range(n):
        cfg = dict(gain=[0.8, 1.0, 2.5, 0.3][s % 4], ng=(s % 3 == 1), dc=(s % 5 == 2))
        cfgs.append(cfg)
        v.set_gain(s, cfg["gain"])
        if cfg["ng"]:
            v.enable_noise_gate(s, True)
            v.set_noise_gate_threshold(s, 0.05)
            v.set_noise_gate_floorgain(s, 0.02)
        if cfg["dc"]:
            v.remove_dc(s, True)
    states = []
    for s in range(n):
        st = OrcVolumeState()
        L.orc_volume_init(C.byref(st), rate)
        st.gain = st.target_gain = st.static_gain = cfgs[s]["gain"]
        if cfgs[s]["ng"]:
            st.noise_gate_enabled, st.ng_threshold, st.ng_floorgain = 1, 0.05, 0.02
            st.gain = st.target_gain = 0.02
        st.remove_dc = int(cfgs[s]["dc"])
        states.append(st)
    t = np.arange(T * nsamples)
    for k in range(T):
        env = 1.0 if (k // 5) % 2 == 0 else 0.01
        x = (noise(100 + k, (n, nsamples), 12000).astype(np.float64) * env + 700).astype(np.int16)
        if k == 3:
            x[:, :] = 32767
        if k == 4:
            x[:, :] = -32768
        got = v.process(x)
        exp = x.copy()
        for s in range(n):
            L.orc_volume_process(C.byref(states[s]), ptr(exp[s]), nsamples)
        assert np.array_equal(got, exp), k
    for s in range(n):
        st = v.state(s)
        for f in ("energy", "level_pk", "instant_energy", "gain", "ng_gain"):
            assert np.float32(getattr(st, f)).tobytes() == np.float32(getattr(states[s], f)).tobytes(), (s, f)
        assert st.dc_offset == states[s].dc_offset and st.ng_noise_dur == states[s].ng_noise_dur
    v.close()


@pytest.mark.parametrize("nsamples", [3840, 8192])
def test_volume_long_blocks(ctx, nsamples):
    """blocks beyond the default 48 KB of shared memory (48 kHz stereo at 40 ms = 3840 samples; the bank's limit 8192): the
    launch opts in / narrows its CTAs instead of failing with 'invalid argument'"""
    L = O.oracle()
    n, rate = 5, 48000
    v = F.Volume(ctx, n, rate, max_block=8192)
    states = []
    for s in range(n):
        v.set_gain(s, 0.5 + 0.25 * s)
        st = OrcVolumeState()
        L.orc_volume_init(C.byref(st), rate)
        st.gain = st.target_gain = st.static_gain = 0.5 + 0.25 * s
        states.append(st)
    for k in range(3):
        x = noise(300 + k, (n, nsamples), 15000)
        got = v.process(x)
        exp = x.copy()
        for s in range(n):
            L.orc_volume_process(C.byref(states[s]), ptr(exp[s]), nsamples)
        assert np.array_equal(got, exp), k
    v.close()


# ------------------------------------------------------------------------------------------------ channel adapter
def test_chanadapt_bit_exact(ctx):
    L = O.oracle()
    ad = F.ChannelAdapter(ctx)
    n, frames = 9, 480
    a, b = noise(1, (n, frames), 30000), noise(2, (n, frames), 30000)
    st = noise(3, (n, frames * 2), 30000)
    for mode, ins in ((0, (a, None)), (1, (st, None)), (2, (a, b)), (2, (a, None)), (2, (None, b))):
        got = ad.process(mode, *ins)
        exp = np.zeros_like(got)
        L.orc_chanadapt(mode, n, frames, ptr(ins[0]), ptr(ins[1]), ptr(exp))
        assert np.array_equal(got, exp), mode


# ------------------------------------------------------------------------------------------------ equalizer
@pytest.mark.parametrize("rate", [8000, 16000, 48000])
def test_equalizer_fir_bit_exact_given_equal_taps(ctx, rate):
    L = O.oracle()
    n, T = 5, 6
    ns = rate // 100
    e = F.Equalizer(ctx, n, rate)
    rng = np.random.default_rng(4)
    taps = (rng.standard_normal((n, e.nfft)) / 12).astype(np.float32)
    mems = np.zeros((n, e.nfft), np.float32)
    for s in range(n):
        e.set_taps(s, taps[s])
    e.set_active(3, False)
    for k in range(T):
        x = noise(50 + k, (n, ns), 32767)
        got = e.process(x)
        exp = x.copy()
        for s in range(n):
            if s != 3:
                L.orc_fir_s16(ptr(taps[s]), e.nfft, ptr(mems[s]), ptr(exp[s]), ns)
        assert np.array_equal(got, exp), k
    e.close()


def test_equalizer_design_matches_oracle_design(ctx):
    """MS_EQUALIZER_SET_GAIN path: host tap design vs the oracle's design. Both go through the bit-exact restatement of the
    reference's float kiss_fft (tests/test_equalizer_design_host.py pins the taps on the CPU; the oracle equals the
    reference filter): equal taps, equal samples."""
    L = O.oracle()
    rate = 16000
    e = F.Equalizer(ctx, 2, rate)
    o = L.orc_equalizer_new(rate)
    for (f, g, w) in [(1000, 2.0, 200), (3000, 0.4, 400)]:
        e.set_gain(1, f, g, w)
        L.orc_equalizer_set_gain(o, f, g, w)
    taps = np.ctypeslib.as_array(L.orc_equalizer_taps(o), shape=(e.nfft,)).copy()
    assert np.array_equal(e.get_taps(1).view(np.uint32), taps.view(np.uint32))
    assert abs(e.get_gain(1, 1000.0) - L.orc_equalizer_get_gain(o, 1000.0)) == 0
    # MS_EQUALIZER_GET_GAIN reads fft_cpx[idx*2] (equalizer.c:121-125), an imaginary-part slot: 0 for an untouched table
    o0 = L.orc_equalizer_new(rate)
    assert e.get_gain(0, 1000.0) == L.orc_equalizer_get_gain(o0, 1000.0)
    L.orc_equalizer_free(o0)
    x = noise(8, (2, 160), 9000)
    got = e.process(x)
    exp = x[1].copy()
    L.orc_equalizer_process(o, ptr(exp), 160)
    assert np.array_equal(got[1], exp)
    L.orc_equalizer_free(o)
    e.close()


# ------------------------------------------------------------------------------------------------ resampler
@pytest.mark.parametrize("in_rate,out_rate,nch", [(8000, 48000, 1), (16000, 48000, 1), (48000, 16000, 1), (44100, 48000, 1),
                                                  (48000, 8000, 1), (16000, 48000, 2), (48000, 44100, 1), (11025, 48000, 1)])
def test_resample_bit_exact_vs_oracle(ctx, in_rate, out_rate, nch):
    """BASELINE cfg1 (8 kHz -> 48 kHz) and the other ratios; GPU == scalar-order oracle bit for bit, including the
    block lengths (inlen*out/in+1 cap, msresample.c:151-152)."""
    L = O.oracle()
    n, T = 6, 25
    frames = in_rate // 100
    r = F.Resample(ctx, n, in_rate, out_rate, nch, max_in_frames=frames)
    orcs = [L.orc_resampler_new(nch, in_rate, out_rate, 3) for _ in range(n)]
    t = np.arange(T * frames)
    for k in range(T):
        seg = t[k * frames:(k + 1) * frames]
        x = np.stack([(8000 * np.sin(2 * np.pi * (300 + 137 * s) * seg / in_rate)).astype(np.int16) for s in range(n)])
        x = x + noise(k, (n, frames), 200)
        if k == 7:
            x[:] = 32767  # saturation in WORD2INT
        if nch == 2:
            x = np.stack([x, -x], axis=-1).reshape(n, frames * 2)
        x = np.ascontiguousarray(x.astype(np.int16))
        got = r.process(x)
        for s in range(n):
            out = np.zeros((frames * out_rate // in_rate + 8) * nch, np.int16)
            m = L.orc_msresample_block(orcs[s], ptr(x[s]), frames, ptr(out))
            assert got.shape[1] == m * nch, (k, got.shape, m)
            assert np.array_equal(got[s], out[:m * nch]), (k, s)
    for o in orcs:
        L.orc_resampler_free(o)
    r.close()


def test_resample_cfg1_hello_like_stream_matches_oracle(ctx):
    """cfg1: one stream, 80-sample ticks at 8 kHz -> 480 samples per tick at 48 kHz."""
    L = O.oracle()
    r = F.Resample(ctx, 1, 8000, 48000, 1, 80)
    o = L.orc_resampler_new(1, 8000, 48000, 3)
    rng = np.random.default_rng(0)
    speechish = np.cumsum(rng.standard_normal(80 * 50)) * 300
    x = np.clip(speechish - speechish.mean(), -20000, 20000).astype(np.int16)
    for k in range(50):
        blk = x[k * 80:(k + 1) * 80]
        got = r.process(blk[None, :])
        out = np.zeros(488, np.int16)
        m = L.orc_msresample_block(o, ptr(blk), 80, ptr(out))
        assert m == 480 and got.shape == (1, 480)
        assert np.array_equal(got[0], out[:480])
    L.orc_resampler_free(o)
    r.close()


def test_volume_lane_kernel_equals_warp_kernel_on_ragged_blocks(ctx):
    """msb200_volume_process_blocks with per-stream counts (what the plugin's groups stage: 0 .. 3 blocks of 256 per tick) on a
    bank large enough to run the lane-per-stream kernel by default: bytes and every state field equal the warp-per-stream
    kernel's (which the tests above pin against the oracle)"""
    n, nblocks, ns, rate, T = 300, 3, 256, 48000, 6
    rng = np.random.default_rng(8)
    outs, states = [], []
    for kernel in (1, 0):  # 0: by size -> lanes (300 >= 256)
        v = F.Volume(ctx, n, rate)
        v.set_kernel(kernel)
        for s in range(n):
            v.set_gain(s, [0.8, 1.0, 1.7][s % 3])
            if s % 4 == 1:
                v.remove_dc(s, True)
            if s % 5 == 2:
                v.enable_noise_gate(s, True)
        r = np.random.default_rng(9)
        got = []
        for k in range(T):
            io = (r.standard_normal((n, nblocks * ns)) * (9000 if k % 2 == 0 else 60)).astype(np.int16) + 300
            counts = r.integers(0, nblocks + 1, size=n).astype(np.int32)
            _lib.check(ctx.lib.msb200_volume_process_blocks(v.h, ptr(io), ns, nblocks * ns, nblocks, ptr(counts)))
            got.append(io.copy())
        outs.append(np.stack(got))
        states.append([bytes(v.state(s)) for s in range(n)])
        v.close()
    assert np.array_equal(outs[0], outs[1])
    assert states[0] == states[1]


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("agc,peer", [(True, False), (False, True), (True, True)])
def test_volume_chunked_mode_bit_exact(ctx, agc, peer, kernel):
    """AGC / echo-limiter peer (msvolume.c:480-502) on 10 ms chunks: speaker bank first, then the microphone bank whose
    streams read their peer's energy; GPU == oracle (pinned vs the reference in test_oracle_vs_reference.py)."""
    L = O.oracle()
    n_streams, T, rate = 9, 40, 16000
    n = rate // 100
    spk_bank, mic_bank = F.Volume(ctx, n_streams, rate), F.Volume(ctx, n_streams, rate)
    spk_bank.set_kernel(kernel)
    mic_bank.set_kernel(kernel)
    st_spk, st_mic = [], []
    for s in range(n_streams):
        a, b = OrcVolumeState(), OrcVolumeState()
        L.orc_volume_init(C.byref(a), rate)
        L.orc_volume_init(C.byref(b), rate)
        b.gain = b.target_gain = b.static_gain = 1.5
        mic_bank.set_gain(s, 1.5)
        if agc:
            b.agc_enabled = 1
            mic_bank.enable_agc(s, True)
        if peer:
            ps = (s + 1) % n_streams  # peer = another stream of the speaker bank
            b.peer = ps
            b.ea_thres, b.vol_upramp, b.force, b.sustain_time = 0.05, 0.3, 6.0, 100
            mic_bank.set_peer(s, spk_bank, ps)
            mic_bank.set_ea(s, threshold=0.05, speed=0.3, force=6.0, sustain=100)
        st_spk.append(a)
        st_mic.append(b)
    rng = np.random.default_rng(3)
    for k in range(T):
        env = 9000 if (k // 6) % 2 == 0 else 40
        spk = (rng.standard_normal((n_streams, n)) * env).astype(np.int16)
        mic = (rng.standard_normal((n_streams, n)) * 3000).astype(np.int16)
        got_spk = spk_bank.process(spk)
        got_mic = mic_bank.process(mic)
        exp_spk, exp_mic = spk.copy(), mic.copy()
        for s in range(n_streams):
            L.orc_volume_process(C.byref(st_spk[s]), ptr(exp_spk[s]), n)
        for s in range(n_streams):
            pe = C.c_float(st_spk[st_mic[s].peer].energy) if peer else None
            L.orc_volume_process_chunk(C.byref(st_mic[s]), C.byref(pe) if peer else None, ptr(exp_mic[s]), n)
        assert np.array_equal(got_spk, exp_spk), k
        assert np.array_equal(got_mic, exp_mic), k
    for s in range(n_streams):
        st = mic_bank.state(s)
        for f in ("energy", "gain", "target_gain", "lt_speaker_en"):
            assert np.float32(getattr(st, f)).tobytes() == np.float32(getattr(st_mic[s], f)).tobytes(), (s, f)
        assert st.sustain_dur == st_mic[s].sustain_dur
    spk_bank.close()
    mic_bank.close()
