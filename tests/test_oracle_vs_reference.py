"""Pins the CPU restatements (oracle/oracle_audio.c, oracle_video.c) against the UNMODIFIED reference filters running in
an unmodified MSTicker graph (oracle/_ref/libms2ref.so). Bit-exact for everything asserted here."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
from _oracle import EqualizerGain, MixerCtl, OrcVolumeState, RefGraph, ptr


def lcg_noise(seed, n, amp):
    rng = np.random.default_rng(seed)
    return rng.integers(-amp, amp + 1, size=n).astype(np.int16)


def run_ref_mixer(pcm, rate, conf, gains=None, inactive=(), nticks_extra=2):
    """pcm [pins][ticks*nwords]"""
    P, total = pcm.shape
    nwords = rate // 100
    T = total // nwords
    g = RefGraph()
    mix = g.new("MSAudioMixer")
    g.call_int(mix, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_int(mix, "MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE", int(conf))
    for p, gain in (gains or {}).items():
        ctl = MixerCtl(pin=p)
        ctl.param.gain = gain
        assert g.call(mix, "MS_AUDIO_MIXER_SET_INPUT_GAIN", ctl) == 0
    for p in inactive:
        ctl = MixerCtl(pin=p)
        ctl.param.active = 0
        assert g.call(mix, "MS_AUDIO_MIXER_SET_ACTIVE", ctl) == 0
    srcs, sinks = [], []
    for p in range(P):
        s = g.source(pcm[p], nwords * 2)
        k = g.sink()
        g.link(s, 0, mix, p)
        g.link(mix, p, k, 0)
        srcs.append(s)
        sinks.append(k)
    g.run(srcs[0], T + nticks_extra)
    outs = [g.read(k)[0] for k in sinks]
    g.close()
    return outs, T * nwords


@pytest.mark.parametrize("conf", [True, False])
def test_mixer_oracle_bit_exact_vs_reference(conf):
    L = O.oracle()
    rate, P, T = 48000, 16, 12
    nwords = rate // 100
    pcm = np.stack([lcg_noise(100 + p, T * nwords, 6000) for p in range(P)])
    pcm[:, 3 * nwords:4 * nwords] = np.where(pcm[:, 3 * nwords:4 * nwords] > 0, 30000, -30000)  # saturation path
    pcm[2, 5 * nwords:6 * nwords] = -32768
    gains = {3: 0.5, 5: 1.7}
    inactive = (7,)
    outs, n = run_ref_mixer(pcm, rate, conf, gains, inactive)
    gain = np.ones(P, np.float32)
    for p, v in gains.items():
        gain[p] = v
    active = np.ones(P, np.uint8)
    active[list(inactive)] = 0
    present = np.ones(P, np.uint8)
    for t in range(T):
        blk = np.ascontiguousarray(pcm[:, t * nwords:(t + 1) * nwords])
        out = np.zeros((P, nwords) if conf else (1, nwords), np.int16)
        L.orc_mixer_process(1, P, nwords, int(conf), ptr(gain), ptr(active), ptr(blk), ptr(present), ptr(out))
        for p in range(P):
            expect = outs[p][t * nwords:(t + 1) * nwords]
            got = out[p] if conf else out[0]
            assert np.array_equal(got, expect), (t, p)


def test_mixer_absent_pin_contributes_zeros():
    """A pin whose bufferizer underruns contributes zeros and receives the full mix (audiomixer.c:88, 113-130)."""
    L = O.oracle()
    rate, P, T = 16000, 4, 6
    nwords = rate // 100
    pcm = np.stack([lcg_noise(7 + p, T * nwords, 9000) for p in range(P)])
    g = RefGraph()
    mix = g.new("MSAudioMixer")
    g.call_int(mix, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_int(mix, "MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE", 1)
    srcs, sinks = [], []
    for p in range(P):
        s = g.source()
        for t in range(T):
            if p == 2 and t in (2, 3):
                continue  # pin 2 starves on ticks 2 and 3
            g.push(s, t, pcm[p, t * nwords:(t + 1) * nwords])
        k = g.sink()
        g.link(s, 0, mix, p)
        g.link(mix, p, k, 0)
        srcs.append(s)
        sinks.append(k)
    g.run(srcs[0], T)
    outs = [g.read(k)[0] for k in sinks]
    g.close()
    gain = np.ones(P, np.float32)
    active = np.ones(P, np.uint8)
    for t in range(T):
        present = np.ones(P, np.uint8)
        if t in (2, 3):
            present[2] = 0
        blk = np.ascontiguousarray(pcm[:, t * nwords:(t + 1) * nwords])
        out = np.zeros((P, nwords), np.int16)
        L.orc_mixer_process(1, P, nwords, 1, ptr(gain), ptr(active), ptr(blk), ptr(present), ptr(out))
        for p in range(P):
            assert np.array_equal(out[p], outs[p][t * nwords:(t + 1) * nwords]), (t, p)


@pytest.mark.parametrize("cfg", [dict(gain=0.8), dict(gain=1.0), dict(gain=2.5), dict(gain=0.8, ng=True),
                                 dict(gain=0.6, dc=True), dict(gain=1.3, ng=True, dc=True)])
def test_volume_oracle_bit_exact_vs_reference(cfg):
    L = O.oracle()
    rate, T = 48000, 40
    n = rate // 100
    t = np.arange(T * n)
    env = (np.sin(2 * np.pi * 1.5 * t / rate) > 0).astype(np.float64)
    x = (env * 9000 * np.sin(2 * np.pi * 440 * t / rate) + 300 * np.sin(2 * np.pi * 50 * t / rate) + 1200).astype(np.int16)
    x[5 * n:6 * n] = 32767
    x[6 * n:7 * n] = -32768
    g = RefGraph()
    vol = g.new("MSVolume")
    g.call_int(vol, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_float(vol, "MS_VOLUME_SET_GAIN", cfg["gain"])
    if cfg.get("ng"):
        g.call(vol, "MS_VOLUME_ENABLE_NOISE_GATE", C.c_ubyte(1))
        g.call_float(vol, "MS_VOLUME_SET_NOISE_GATE_THRESHOLD", 0.05)
        g.call_float(vol, "MS_VOLUME_SET_NOISE_GATE_FLOORGAIN", 0.02)
    if cfg.get("dc"):
        g.call_int(vol, "MS_VOLUME_REMOVE_DC", 1)
    src = g.source(x, n * 2)
    sink = g.sink()
    g.link(src, 0, vol, 0)
    g.link(vol, 0, sink, 0)
    g.run(src, T)
    y_ref, _ = g.read(sink)
    lin = C.c_float()
    g.call(vol, "MS_VOLUME_GET_LINEAR", lin)
    g.close()

    st = OrcVolumeState()
    L.orc_volume_init(C.byref(st), rate)
    st.gain = st.target_gain = st.static_gain = cfg["gain"]
    if cfg.get("ng"):
        st.noise_gate_enabled = 1
        st.ng_threshold = 0.05
        st.ng_floorgain = 0.02
        st.gain = st.target_gain = 0.005  # enable happens before the floorgain is raised (msvolume.c:352-359)
        st.gain = st.target_gain = 0.02   # then set_noise_gate_floorgain re-applies it (:367-378)
    if cfg.get("dc"):
        st.remove_dc = 1
    y = x.copy()
    for k in range(T):
        L.orc_volume_process(C.byref(st), ptr(y[k * n:(k + 1) * n]), n)
    assert np.array_equal(y, y_ref)
    # MS_VOLUME_GET_LINEAR returns the smoothed energy: compare exactly (same float ops, same order)
    assert np.float32(lin.value) == np.float32(st.energy)


@pytest.mark.parametrize("mode", [0, 1])
def test_chanadapt_oracle_bit_exact_vs_reference(mode):
    L = O.oracle()
    rate, T = 16000, 5
    n = rate // 100
    g = RefGraph()
    ad = g.new("MSChannelAdapter")
    g.call_int(ad, "MS_FILTER_SET_SAMPLE_RATE", rate)
    inch, outch = (1, 2) if mode == 0 else (2, 1)
    g.call_int(ad, "MS_FILTER_SET_NCHANNELS", inch)
    g.call_int(ad, "MS_CHANNEL_ADAPTER_SET_OUTPUT_NCHANNELS", outch)
    x = lcg_noise(3, T * n * inch, 20000)
    src = g.source(x, n * 2 * inch)
    sink = g.sink()
    g.link(src, 0, ad, 0)
    g.link(ad, 0, sink, 0)
    g.run(src, T)
    y_ref, _ = g.read(sink)
    g.close()
    out = np.zeros(T * n * outch, np.int16)
    L.orc_chanadapt(mode, 1, T * n, ptr(x), None, ptr(out))
    assert np.array_equal(out, y_ref)


def test_chanadapt_two_mono_inputs_vs_reference():
    L = O.oracle()
    rate, T = 8000, 6
    n = rate // 100
    a, b = lcg_noise(11, T * n, 15000), lcg_noise(12, T * n, 15000)
    g = RefGraph()
    ad = g.new("MSChannelAdapter")
    g.call_int(ad, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_int(ad, "MS_FILTER_SET_NCHANNELS", 2)
    g.call_int(ad, "MS_CHANNEL_ADAPTER_SET_OUTPUT_NCHANNELS", 1)
    s0, s1 = g.source(a, n * 2), g.source(b, n * 2)
    sink = g.sink()
    g.link(s0, 0, ad, 0)
    g.link(s1, 0, ad, 1)
    g.link(ad, 0, sink, 0)
    g.run(s0, T)
    y_ref, _ = g.read(sink)
    g.close()
    out = np.zeros(T * n * 2, np.int16)
    L.orc_chanadapt(2, 1, T * n, ptr(a), ptr(b), ptr(out))
    assert len(y_ref) == len(out)
    assert np.array_equal(out, y_ref)


def test_fir_oracle_bit_exact_vs_reference_function():
    """orc_fir_mem16 == ms_fir_mem16 (dsptools.c:253-268) on random taps and blocks, state carried across calls."""
    L, R = O.oracle(), O.ref()
    rng = np.random.default_rng(5)
    for ord_ in (128, 256, 512):
        taps = (rng.standard_normal(ord_) / 16).astype(np.float32)
        mem_a, mem_b = np.zeros(ord_, np.float32), np.zeros(ord_, np.float32)
        for n in (480, 160, 333):
            x = rng.integers(-32768, 32768, n).astype(np.float32)
            ya, yb = np.zeros(n, np.float32), np.zeros(n, np.float32)
            L.orc_fir_mem16(ptr(x), ptr(taps), ptr(ya), n, ord_, ptr(mem_a))
            R.ref_fir_mem16(ptr(x), ptr(taps), ptr(yb), n, ord_, ptr(mem_b))
            assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32))


@pytest.mark.parametrize("rate", [8000, 16000, 48000])
def test_equalizer_oracle_vs_reference_filter(rate):
    """Full MSEqualizer (gain table -> taps -> FIR) through the ticker. The inverse transform is the bit-exact restatement
    of the reference's float kiss_fft (oracle_plc.c), the FIR keeps ms_fir_mem16's order: every output sample is equal."""
    L = O.oracle()
    T = 12
    n = rate // 100
    t = np.arange(T * n)
    x = (6000 * np.sin(2 * np.pi * 300 * t / rate) + 5000 * np.sin(2 * np.pi * 1000 * t / rate) +
         3000 * np.sin(2 * np.pi * 3000 * t / rate)).astype(np.int16)
    g = RefGraph()
    eq = g.new("MSEqualizer")
    g.call_int(eq, "MS_FILTER_SET_SAMPLE_RATE", rate)
    for (f, gn, w) in [(1000, 2.0, 200), (3000, 0.4, 400)]:
        g.call(eq, "MS_EQUALIZER_SET_GAIN", EqualizerGain(f, gn, w))
    src = g.source(x, n * 2)
    sink = g.sink()
    g.link(src, 0, eq, 0)
    g.link(eq, 0, sink, 0)
    g.run(src, T)
    y_ref, _ = g.read(sink)
    gg = EqualizerGain(1000, 0, 0)
    g.call(eq, "MS_EQUALIZER_GET_GAIN", gg)
    g.close()
    e = L.orc_equalizer_new(rate)
    for (f, gn, w) in [(1000, 2.0, 200), (3000, 0.4, 400)]:
        L.orc_equalizer_set_gain(e, f, gn, w)
    assert abs(L.orc_equalizer_get_gain(e, 1000.0) - gg.gain) < 1e-6
    y = x.copy()
    for k in range(T):
        L.orc_equalizer_process(e, ptr(y[k * n:(k + 1) * n]), n)
    L.orc_equalizer_free(e)
    assert np.array_equal(y, y_ref)
    assert np.abs(y_ref).max() > 3000  # non-trivial signal came through


@pytest.mark.parametrize("rotation", [0, 90, 180, 270])
@pytest.mark.parametrize("down_scale", [0, 1])
def test_nv12_oracle_bit_exact_vs_reference(rotation, down_scale):
    """Same pattern as the reference's own test (tester/mediastreamer2_framework_tester.c:219-367): y[i]=i%256,
    cbcr[i]=i%256, padded strides; VGA."""
    L, R = O.oracle(), O.ref()
    f = 2 if down_scale else 1
    w, h = (640 // f, 480 // f) if rotation % 180 == 0 else (480 // f, 640 // f)
    sw, sh = 640, 480
    y_stride = sw + sw % 32 + 32
    c_stride = sw + 64
    ybuf = (np.arange(y_stride * sh) % 256).astype(np.uint8)
    cbuf = (np.arange(c_stride * sh // 2) % 256).astype(np.uint8)
    for u_first in (1, 0):
        a = np.zeros(w * h * 3 // 2 + 64, np.uint8)
        b = np.zeros_like(a)
        na = L.orc_nv12_to_i420(ptr(ybuf), ptr(cbuf), rotation, w, h, y_stride, c_stride, u_first, down_scale, ptr(a))
        nb = R.ref_nv12_to_i420(ptr(ybuf), ptr(cbuf), rotation, w, h, y_stride, c_stride, u_first, down_scale, ptr(b))
        assert na == nb == w * h * 3 // 2
        assert np.array_equal(a, b)


def _speechy(seed, n, rate, amp):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    env = (np.sin(2 * np.pi * 1.3 * t / rate + seed) > 0).astype(np.float64)
    return (env * amp * np.sin(2 * np.pi * (200 + 37 * seed) * t / rate) + 60 * rng.standard_normal(n)).astype(np.int16)


@pytest.mark.parametrize("agc,peer", [(True, False), (False, True), (True, True)])
def test_volume_chunked_mode_oracle_bit_exact_vs_reference(agc, peer):
    """msvolume.c:480-502: AGC and/or echo-limiter peer -> 10 ms chunks, echo avoider reads the peer's energy.
    Two MSVolume filters in one ticker (speaker path first, then microphone path with the speaker filter as peer)."""
    L = O.oracle()
    rate, T = 16000, 60
    n = rate // 100
    spk = _speechy(1, T * n, rate, 9000)
    mic = _speechy(2, T * n, rate, 5000)
    g = RefGraph()
    vspk, vmic = g.new("MSVolume"), g.new("MSVolume")
    for v in (vspk, vmic):
        g.call_int(v, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_float(vmic, "MS_VOLUME_SET_GAIN", 1.5)
    if agc:
        g.call_int(vmic, "MS_VOLUME_ENABLE_AGC", 1)
    if peer:
        g.call_ptr(vmic, "MS_VOLUME_SET_PEER", vspk)
        g.call_float(vmic, "MS_VOLUME_SET_EA_THRESHOLD", 0.05)
        g.call_float(vmic, "MS_VOLUME_SET_EA_SPEED", 0.3)
        g.call_float(vmic, "MS_VOLUME_SET_EA_FORCE", 6.0)
        g.call_int(vmic, "MS_VOLUME_SET_EA_SUSTAIN", 100)
    # odd block sizes on the microphone path: the chunked mode re-frames them to 10 ms
    s_spk, k_spk = g.source(spk, n * 2), g.sink()
    s_mic, k_mic = g.source(), g.sink()
    blk = 2 * n // 3
    pos = 0
    tick = 0
    while pos < len(mic):
        g.push(s_mic, tick, mic[pos:pos + blk])
        pos += blk
        tick += 1 if (tick % 3) else 0
        tick += 1 if pos % (3 * blk) == 0 else 0
    g.link(s_spk, 0, vspk, 0)
    g.link(vspk, 0, k_spk, 0)
    g.link(s_mic, 0, vmic, 0)
    g.link(vmic, 0, k_mic, 0)
    g.run([s_spk, s_mic], T + 40)
    y_spk, _ = g.read(k_spk)
    y_mic, tri = g.read(k_mic)
    g.close()
    assert set(tri[:, 1]) == {n * 2}  # every output block is one 10 ms chunk
    # oracle: replay tick by tick with the same arrival pattern (chunks are cut as soon as 10 ms are buffered)
    st_spk, st_mic = OrcVolumeState(), OrcVolumeState()
    L.orc_volume_init(C.byref(st_spk), rate)
    L.orc_volume_init(C.byref(st_mic), rate)
    st_mic.gain = st_mic.target_gain = st_mic.static_gain = 1.5
    st_mic.agc_enabled = int(agc)
    if peer:
        st_mic.ea_thres, st_mic.vol_upramp, st_mic.force, st_mic.sustain_time = 0.05, 0.3, 6.0, 100
    # arrival schedule of microphone samples per tick, mirrored from the pushes above
    sched = {}
    pos = 0
    tick = 0
    while pos < len(mic):
        sched.setdefault(tick, []).append((pos, min(pos + blk, len(mic))))
        pos += blk
        tick += 1 if (tick % 3) else 0
        tick += 1 if pos % (3 * blk) == 0 else 0
    buf = np.zeros(0, np.int16)
    out_mic = []
    exp_spk = spk.copy()
    for t in range(T + 40):
        if t < T:
            L.orc_volume_process(C.byref(st_spk), ptr(exp_spk[t * n:(t + 1) * n]), n)
        for (a, b) in sched.get(t, []):
            buf = np.concatenate([buf, mic[a:b]])
        while len(buf) >= n:
            chunk = np.ascontiguousarray(buf[:n])
            buf = buf[n:]
            pe = C.c_float(st_spk.energy)
            L.orc_volume_process_chunk(C.byref(st_mic), C.byref(pe) if peer else None, ptr(chunk), n)
            out_mic.append(chunk)
    out_mic = np.concatenate(out_mic)
    assert np.array_equal(y_spk, exp_spk)
    assert len(y_mic) == len(out_mic)
    assert np.array_equal(y_mic, out_mic)


# ---------------------------------------------------------------------------------------------------- G.711
def test_g711_oracle_exhaustive_vs_reference_functions():
    """oracle/oracle_g711.c == the UNMODIFIED Snack_* routines (g711.c:119-262) on every 16-bit sample and every code"""
    import ctypes as C
    L, R = O.oracle(), O.ref()
    pcm = np.arange(-32768, 32768, dtype=np.int16)
    codes = np.arange(256, dtype=np.uint8)
    for law, enc, dec in ((0, R.Snack_Lin2Alaw, R.Snack_Alaw2Lin), (1, R.Snack_Lin2Mulaw, R.Snack_Mulaw2Lin)):
        enc.restype, enc.argtypes = C.c_ubyte, [C.c_short]
        dec.restype, dec.argtypes = C.c_short, [C.c_ubyte]
        got_c = np.zeros(pcm.size, np.uint8)
        L.orc_g711_encode(law, ptr(pcm), ptr(got_c), pcm.size)
        exp_c = np.array([enc(int(v)) for v in pcm], np.uint8)
        assert np.array_equal(got_c, exp_c), law
        got_p = np.zeros(256, np.int16)
        L.orc_g711_decode(law, ptr(codes), ptr(got_p), 256)
        exp_p = np.array([dec(int(c)) for c in codes], np.int16)
        assert np.array_equal(got_p, exp_p), law


@pytest.mark.parametrize("name,law", [("MSAlaw", 0), ("MSUlaw", 1)])
def test_g711_oracle_vs_reference_filters_in_ticker(name, law):
    """the reference's encoder (MSBufferizer re-framing to ptime, alaw.c:56-94) and decoder (:199-211) filters in the
    unmodified MSTicker == oracle arithmetic on the same samples"""
    L = O.oracle()
    rng = np.random.default_rng(11 + law)
    n, ticks = 80, 24  # 8 kHz, 10 ms blocks
    pcm = (rng.standard_normal(n * ticks) * 9000).clip(-32768, 32767).astype(np.int16)
    pcm[:64] = np.array([-32768, 32767, 0, -1, 1, -8, 8, -9] * 8, np.int16)
    g = RefGraph()
    src, enc, dec, sink_p = g.source(pcm, n * 2), g.new(name + "Enc"), g.new(name + "Dec"), g.sink()
    g.link(src, 0, enc, 0)
    g.link(enc, 0, dec, 0)
    g.link(dec, 0, sink_p, 0)
    g.run(src, ticks + 2)
    out_p, tri = g.read(sink_p)
    g.close()
    # default ptime: 2 frames of 10 ms per packet (alaw.c:59-72) -> blocks of 160 samples
    assert len(out_p) == (n * ticks // 160) * 160 and np.all(tri[:, 1] == 320)  # (tick, bytes, timestamp) per block
    code = np.zeros(len(out_p), np.uint8)
    L.orc_g711_encode(law, ptr(pcm), ptr(code), len(out_p))
    exp = np.zeros(len(out_p), np.int16)
    L.orc_g711_decode(law, ptr(code), ptr(exp), len(out_p))
    assert np.array_equal(out_p, exp)


# ---------------------------------------------------------------------------------------------------- MSAudioFlowControl (§8f-3)
def _flowctl_signal(rng, n, ticks):
    """speech-like bursts, near-silent stretches (whole-frame drops) and flat runs (ties in the three-sample criterion)"""
    x = (rng.standard_normal(n * ticks) * 6000).clip(-32768, 32767).astype(np.int16)
    x[n * 6:n * 9] = rng.integers(-20, 21, n * 3)          # almost silent
    x[n * 14:n * 14 + 40] = 1234                             # flat: many equal minima
    x[n * 20:n * 22] = (np.arange(n * 2) % 7 - 3) * 100      # periodic: repeated minima
    return x


@pytest.mark.parametrize("strategy,drop_ms,interval_ms", [(1, 30, 200), (1, 8, 100), (0, 40, 300), (1, 120, 150)])
def test_flowcontrol_oracle_bit_exact_vs_reference_filter(strategy, drop_ms, interval_ms):
    """oracle == the unmodified MSAudioFlowControl in an MSTicker: same blocks (sizes) and samples, for both strategies,
    silent-frame drops, zero-crossing sample deletion and the too-many-samples whole-frame drop"""
    from _oracle import FlowControlConfig, FlowControlDropEvent, OrcFlowCtl
    L = O.oracle()
    rate, n, ticks = 16000, 160, 40
    x = _flowctl_signal(np.random.default_rng(drop_ms), n, ticks)
    g = RefGraph()
    src, fc, sink = g.source(x, n * 2), g.new("MSAudioFlowControl"), g.sink()
    assert g.call_int(fc, "MS_FILTER_SET_SAMPLE_RATE", rate) == 0
    assert g.call_int(fc, "MS_FILTER_SET_NCHANNELS", 1) == 0
    assert g.call(fc, "MS_AUDIO_FLOW_CONTROL_SET_CONFIG", FlowControlConfig(strategy, 0.02)) == 0
    g.link(src, 0, fc, 0)
    g.link(fc, 0, sink, 0)
    g.run(src, 2)
    assert g.call(fc, "MS_AUDIO_FLOW_CONTROL_DROP", FlowControlDropEvent(interval_ms, drop_ms)) == 0
    g.run(src, ticks)
    ref_out, tri = g.read(sink)
    g.close()
    c = OrcFlowCtl()
    L.orc_flowctl_init(C.byref(c))
    c.strategy = strategy
    out, sizes = [], []
    for t in range(ticks):
        if t == 2:
            L.orc_flowctl_set_target(C.byref(c), drop_ms * rate // 1000, interval_ms * rate // 1000)
        blk = x[t * n:(t + 1) * n].copy()
        k = L.orc_flowctl_process(C.byref(c), ptr(blk), n)
        if k:
            out.append(blk[:k])
            sizes.append(2 * k)
    out = np.concatenate(out)
    assert list(tri[:, 1]) == sizes
    assert np.array_equal(ref_out, out)
    assert len(out) < len(x)  # something was dropped


# ---------------------------------------------------------------------------------------------------- MSGenericPLC
class _CngData(C.Structure):  # MSCngData, include/mediastreamer2/mscngdtx.h:23-26
    _fields_ = [("datasize", C.c_int), ("data", C.c_uint8 * 32)]


def plc_signal(rate: int, nsamples: int, seed: int = 1) -> np.ndarray:
    t = np.arange(nsamples)
    rng = np.random.default_rng(seed)
    x = (6000 * np.sin(2 * np.pi * 440 * t / rate) + 3000 * np.sin(2 * np.pi * 1234.5 * t / rate + 1)
         + rng.normal(0, 300, nsamples))
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def plc_reference_run(rate, ticks, schedule, x, cn_at=(), plugins_dir=None, nchannels=1):
    """the UNMODIFIED MSGenericPLC in the reference's MSTicker; schedule[tick] = list of (offset, nsamples) blocks that
    arrive in that tick; cn_at = ticks before which MS_GENERIC_PLC_SET_CN is called. Returns samples and (tick, bytes)."""
    g = RefGraph(plugins_dir=plugins_dir)
    src, plc, sink = g.source(), g.new("MSGenericPLC"), g.sink()
    assert g.text(plc).startswith("B200:") == bool(plugins_dir)
    assert g.call_int(plc, "MS_FILTER_SET_SAMPLE_RATE", rate) == 0
    assert g.call_int(plc, "MS_FILTER_SET_NCHANNELS", nchannels) == 0
    for k, blocks in schedule.items():
        for off, n in blocks:
            g.push(src, k, x[off:off + n])
    g.link(src, 0, plc, 0)
    g.link(plc, 0, sink, 0)
    done = 0
    for stop in sorted(set(cn_at)) + [ticks]:
        if stop > done:
            g.run(src, stop - done)
            done = stop
        if stop < ticks:
            assert g.call(plc, "MS_GENERIC_PLC_SET_CN", _CngData()) == 0
    out, tri = g.read(sink)
    g.close()
    return out, [(int(a), int(b)) for a, b, _ in tri]


def plc_oracle_run(rate, ticks, schedule, x, cn_at=(), nchannels=1):
    L = O.oracle()
    c = L.orc_plc_create(rate)
    assert c
    out, blocks, kind = [], [], C.c_int(0)
    tick_n = nchannels * rate // 100
    for k in range(ticks):
        if k in cn_at:
            L.orc_plc_filter_set_cn(c)
        for off, n in schedule.get(k, []):
            b = x[off:off + n].copy()
            L.orc_plc_filter_packet(c, k * 10, ptr(b), n, nchannels)
            out.append(b)
            blocks.append((k, 2 * n))
        o = np.zeros(tick_n, np.int16)
        m = L.orc_plc_filter_tick(c, k * 10, 10, nchannels, ptr(o), C.byref(kind))
        if m:
            out.append(o[:m])
            blocks.append((k, 2 * m))
    L.orc_plc_destroy(c)
    return np.concatenate(out), blocks


def plc_schedule(rate, ticks, lost, block_ms=10):
    """one block of block_ms every block_ms, except the block indices in `lost`"""
    n, step = rate * block_ms // 1000, block_ms // 10
    return {k: [((k // step) * n, n)] for k in range(0, ticks, step) if (k // step) not in lost}


@pytest.mark.parametrize("rate", [8000, 16000, 32000, 48000])
def test_plc_oracle_bit_exact_vs_reference_filter(rate):
    """single losses, a 4-block burst, a 220 ms hole (fade to silence, then zeros) and a back-to-back pair: every block the
    unmodified filter emits (sizes, order, samples) equals the oracle's; transform sizes 400..4800 cover radix 2, 3, 4, 5"""
    ticks = 70
    lost = set(range(10, 14)) | {20} | set(range(30, 52)) | {60, 61}
    x = plc_signal(rate, ticks * rate // 100)
    sched = plc_schedule(rate, ticks, lost)
    ref, ref_blocks = plc_reference_run(rate, ticks, sched, x)
    out, blocks = plc_oracle_run(rate, ticks, sched, x)
    assert ref_blocks == blocks
    assert np.array_equal(ref, out)
    n = rate // 100
    assert np.abs(ref[10 * n:12 * n].astype(int)).max() > 1000  # the concealed stretch is not silence


def test_plc_oracle_long_hole_wraps_the_16_bit_counters():
    """plc_samples_used is a uint16_t in the reference (genericplc.h:46): after 65536 concealed samples it wraps and the
    filter re-emits stale generated signal; the oracle reproduces that"""
    rate, ticks = 48000, 260
    x = plc_signal(rate, ticks * 480, seed=3)
    sched = plc_schedule(rate, ticks, set(range(20, 240)))
    ref, ref_blocks = plc_reference_run(rate, ticks, sched, x)
    out, blocks = plc_oracle_run(rate, ticks, sched, x)
    assert ref_blocks == blocks and np.array_equal(ref, out)
    assert np.any(ref[160 * 480:170 * 480] != 0)  # stale signal after the wrap (would be silence with wider counters)


@pytest.mark.parametrize("rate,block_ms", [(16000, 20), (8000, 20), (16000, 30)])
def test_plc_oracle_longer_packets(rate, block_ms):
    ticks = 90
    x = plc_signal(rate, ticks * rate // 100, seed=block_ms)
    sched = plc_schedule(rate, ticks, {5, 6, 12} | set(range(20, 31)), block_ms)
    ref, ref_blocks = plc_reference_run(rate, ticks, sched, x)
    out, blocks = plc_oracle_run(rate, ticks, sched, x)
    # the recording sink counts its own invocations, which are not ticks when some ticks carry nothing: compare sizes
    assert [b for _, b in ref_blocks] == [b for _, b in blocks] and np.array_equal(ref, out)


def test_plc_oracle_comfort_noise_path():
    """MS_GENERIC_PLC_SET_CN before a hole: the hole is filled with flagged silence instead of concealment and the first
    block after it fades in from zero (msgenericplc.c:77-88, 131-141)"""
    rate, ticks = 16000, 60
    x = plc_signal(rate, ticks * 160, seed=9)
    sched = plc_schedule(rate, ticks, set(range(12, 20)) | set(range(30, 34)))
    ref, ref_blocks = plc_reference_run(rate, ticks, sched, x, cn_at=(12,))
    out, blocks = plc_oracle_run(rate, ticks, sched, x, cn_at=(12,))
    assert ref_blocks == blocks and np.array_equal(ref, out)
    assert not ref[12 * 160:20 * 160].any()


def test_plc_oracle_stereo_blocks():
    """nchannels = 2: the reference treats the interleaved block as one long mono signal at `rate` (msgenericplc.c:66-68,
    121) — twice the samples per block and per concealed tick; the oracle follows"""
    rate, ticks, nch = 16000, 60, 2
    n = nch * rate // 100
    x = plc_signal(rate, ticks * n, seed=21)
    sched = {k: [(k * n, n)] for k in range(ticks) if k not in ({7} | set(range(20, 40)))}
    ref, ref_blocks = plc_reference_run(rate, ticks, sched, x, nchannels=nch)
    out, blocks = plc_oracle_run(rate, ticks, sched, x, nchannels=nch)
    assert ref_blocks == blocks and np.array_equal(ref, out)


@pytest.mark.parametrize("seed", range(32))
def test_plc_oracle_random_schedules(seed):
    """random rate, packet length, loss bursts (1..30 blocks) and comfort-noise requests: oracle == reference filter"""
    rng = np.random.default_rng(1000 + seed)
    rate = int(rng.choice([8000, 16000, 32000, 48000]))
    block_ms = int(rng.choice([10, 10, 20]))
    ticks = 120
    nblocks = ticks * 10 // block_ms
    lost, k = set(), int(rng.integers(2, 10))
    while k < nblocks:
        burst = int(rng.choice([1, 1, 2, 3, 5, 12, 30]))
        lost |= set(range(k, min(k + burst, nblocks)))
        k += burst + int(rng.integers(1, 15))
    cn_at = tuple(sorted(int(t) for t in rng.choice(np.arange(5, ticks - 5), size=int(rng.integers(0, 3)), replace=False)))
    if rate > 16000:
        # above 16 kHz the reference's comfort-noise resume reads past its 80-sample stack buffer (msgenericplc.c:79-86:
        # `int16_t continuity_buffer[80]` used for rate * 5 / 1000 samples): undefined there, all-zero in the oracle / GPU
        cn_at = ()
    x = plc_signal(rate, ticks * rate // 100, seed=seed)
    sched = plc_schedule(rate, ticks, lost, block_ms)
    ref, ref_blocks = plc_reference_run(rate, ticks, sched, x, cn_at)
    out, blocks = plc_oracle_run(rate, ticks, sched, x, cn_at)
    assert [b for _, b in ref_blocks] == [b for _, b in blocks]
    assert np.array_equal(ref, out)


@pytest.mark.parametrize("seed", range(24))
def test_flowcontrol_oracle_random_scenarios(seed):
    """random rate, block length, strategy, silence threshold, loudness and two drop requests at random ticks (the second one
    is ignored while the first is still running, flowcontrol.c:196-207): oracle == reference filter, sizes and samples"""
    from _oracle import FlowControlConfig, FlowControlDropEvent, OrcFlowCtl
    L = O.oracle()
    rng = np.random.default_rng(seed)
    rate = int(rng.choice([8000, 16000, 48000]))
    n, ticks = rate // 100 * int(rng.choice([1, 1, 2])), 50
    strategy, thr = int(rng.integers(0, 2)), float(rng.choice([0.02, 0.0, 0.1]))
    drop_ms, interval_ms = int(rng.choice([5, 8, 15, 30, 60, 120, 250])), int(rng.choice([50, 100, 200, 400]))
    x = _flowctl_signal(rng, n, ticks)
    if rng.integers(0, 2):
        x = (x.astype(np.int32) * int(rng.integers(1, 6))).clip(-32768, 32767).astype(np.int16)
    arm = int(rng.integers(1, 10))  # after the attach: preprocess resets the controller (:163-166)
    arm2 = arm + int(rng.integers(3, 30))
    g = RefGraph()
    src, fc, sink = g.source(x, n * 2), g.new("MSAudioFlowControl"), g.sink()
    g.call_int(fc, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_int(fc, "MS_FILTER_SET_NCHANNELS", 1)
    g.call(fc, "MS_AUDIO_FLOW_CONTROL_SET_CONFIG", FlowControlConfig(strategy, thr))
    g.link(src, 0, fc, 0)
    g.link(fc, 0, sink, 0)
    done = 0
    for stop in (arm, arm2, ticks):
        if stop > done:
            g.run(src, stop - done)
            done = stop
        if stop < ticks:
            g.call(fc, "MS_AUDIO_FLOW_CONTROL_DROP", FlowControlDropEvent(interval_ms, drop_ms))
    ref_out, tri = g.read(sink)
    g.close()
    c = OrcFlowCtl()
    L.orc_flowctl_init(C.byref(c))
    c.strategy, c.silent_threshold = strategy, thr
    out, sizes = [], []
    for t in range(ticks):
        if t in (arm, arm2) and not (c.total_samples > 0 and c.target_samples > 0):
            L.orc_flowctl_set_target(C.byref(c), drop_ms * rate // 1000, interval_ms * rate // 1000)
        blk = x[t * n:(t + 1) * n].copy()
        k = L.orc_flowctl_process(C.byref(c), ptr(blk), n)
        if k:
            out.append(blk[:k])
            sizes.append(2 * k)
    assert list(tri[:, 1]) == sizes
    assert np.array_equal(ref_out, np.concatenate(out) if out else np.zeros(0, np.int16))


@pytest.mark.parametrize("seed", range(16))
def test_mixer_oracle_random_rooms(seed):
    """random rate, 2..12 pins, both modes, gains (0 and > 1 included), muted pins, a pin stuck at -32768, pins that starve
    on random ticks: every output block of the reference filter equals the oracle's"""
    L = O.oracle()
    rng = np.random.default_rng(seed)
    rate, P, T = int(rng.choice([8000, 16000, 48000])), int(rng.integers(2, 13)), 10
    nwords, conf = rate // 100, bool(rng.integers(0, 2))
    amp = int(rng.choice([3000, 9000, 30000]))
    pcm = rng.integers(-amp, amp + 1, size=(P, T * nwords)).astype(np.int16)
    if rng.integers(0, 2):
        pcm[int(rng.integers(0, P)), 2 * nwords:3 * nwords] = -32768
    gains = {int(p): float(rng.choice([0.5, 1.7, 0.0, 2.5, 1.0])) for p in rng.choice(P, size=int(rng.integers(0, P)), replace=False)}
    inactive = tuple(int(p) for p in rng.choice(P, size=int(rng.integers(0, max(1, P // 2))), replace=False))
    starve = {(int(rng.integers(0, P)), int(rng.integers(1, T))) for _ in range(int(rng.integers(0, 4)))}
    g = RefGraph()
    mix = g.new("MSAudioMixer")
    g.call_int(mix, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_int(mix, "MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE", int(conf))
    for p, gn in gains.items():
        ctl = MixerCtl(pin=p)
        ctl.param.gain = gn
        assert g.call(mix, "MS_AUDIO_MIXER_SET_INPUT_GAIN", ctl) == 0
    for p in inactive:
        ctl = MixerCtl(pin=p)
        ctl.param.active = 0
        assert g.call(mix, "MS_AUDIO_MIXER_SET_ACTIVE", ctl) == 0
    srcs, sinks = [], []
    for p in range(P):
        s = g.source()
        for t in range(T):
            if (p, t) not in starve:
                g.push(s, t, pcm[p, t * nwords:(t + 1) * nwords])
        k = g.sink()
        g.link(s, 0, mix, p)
        g.link(mix, p, k, 0)
        srcs.append(s)
        sinks.append(k)
    g.run(srcs[0], T)
    outs = [g.read(k)[0] for k in sinks]
    g.close()
    gain = np.ones(P, np.float32)
    for p, v in gains.items():
        gain[p] = v
    active = np.ones(P, np.uint8)
    active[list(inactive)] = 0
    for t in range(T):
        present = np.ones(P, np.uint8)
        for p, tt in starve:
            if tt == t:
                present[p] = 0
        blk = np.ascontiguousarray(pcm[:, t * nwords:(t + 1) * nwords])
        out = np.zeros((P, nwords) if conf else (1, nwords), np.int16)
        L.orc_mixer_process(1, P, nwords, int(conf), ptr(gain), ptr(active), ptr(blk), ptr(present), ptr(out))
        for p in range(P):
            assert np.array_equal(out[p] if conf else out[0], outs[p][t * nwords:(t + 1) * nwords]), (t, p)


@pytest.mark.parametrize("seed", range(16))
def test_volume_oracle_random_configs(seed):
    """random rate, block length, gain (0 .. 6), noise gate threshold / floor gain (below the 0.005 minimum included), DC
    removal, signal level, DC offset and noise: samples and the smoothed energy equal the reference filter's bit for bit"""
    L = O.oracle()
    rng = np.random.default_rng(seed)
    rate, T = int(rng.choice([8000, 16000, 48000])), 40
    n = rate // 100 * int(rng.choice([1, 1, 2]))
    t = np.arange(T * n)
    env = (np.sin(2 * np.pi * float(rng.uniform(0.5, 3)) * t / rate) > 0).astype(np.float64)
    amp = float(rng.choice([300, 3000, 9000, 20000, 32000]))
    x = (env * amp * np.sin(2 * np.pi * float(rng.uniform(100, 3000)) * t / rate)
         + float(rng.uniform(0, 500)) * np.sin(2 * np.pi * 50 * t / rate) + float(rng.uniform(-2000, 2000))
         + rng.normal(0, float(rng.choice([0, 30, 300])), T * n)).clip(-32768, 32767).astype(np.int16)
    if rng.integers(0, 2):
        x[5 * n:6 * n] = 32767
        x[6 * n:7 * n] = -32768
    gain = float(rng.choice([0.0, 0.1, 0.8, 1.0, 1.3, 2.5, 6.0]))
    ng, dc = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    thr, floor = float(rng.choice([0.01, 0.05, 0.2])), float(rng.choice([0.0, 0.02, 0.3]))
    g = RefGraph()
    vol = g.new("MSVolume")
    g.call_int(vol, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_float(vol, "MS_VOLUME_SET_GAIN", gain)
    if ng:
        g.call(vol, "MS_VOLUME_ENABLE_NOISE_GATE", C.c_ubyte(1))
        g.call_float(vol, "MS_VOLUME_SET_NOISE_GATE_THRESHOLD", thr)
        g.call_float(vol, "MS_VOLUME_SET_NOISE_GATE_FLOORGAIN", floor)
    if dc:
        g.call_int(vol, "MS_VOLUME_REMOVE_DC", 1)
    src, sink = g.source(x, n * 2), g.sink()
    g.link(src, 0, vol, 0)
    g.link(vol, 0, sink, 0)
    g.run(src, T)
    y_ref, _ = g.read(sink)
    lin = C.c_float()
    g.call(vol, "MS_VOLUME_GET_LINEAR", lin)
    g.close()
    st = OrcVolumeState()
    L.orc_volume_init(C.byref(st), rate)
    st.gain = st.target_gain = st.static_gain = gain
    if ng:
        st.noise_gate_enabled = 1
        st.ng_threshold = thr
        st.ng_floorgain = max(floor, 0.005)  # volume_set_noise_gate_floorgain clamps (msvolume.c:366-369)
        st.gain = st.target_gain = st.ng_floorgain  # soft start (:352, :371)
    if dc:
        st.remove_dc = 1
    y = x.copy()
    for k in range(T):
        L.orc_volume_process(C.byref(st), ptr(y[k * n:(k + 1) * n]), n)
    assert np.array_equal(y, y_ref)
    assert np.float32(lin.value) == np.float32(st.energy)


@pytest.mark.parametrize("seed", range(24))
def test_nv12_oracle_random_geometries(seed):
    """random frame sizes (8 .. 320 wide), paddings, rotations, decimation and plane order on random pixels"""
    L, R = O.oracle(), O.ref()
    rng = np.random.default_rng(seed)
    rotation, ds = int(rng.choice([0, 90, 180, 270])), int(rng.integers(0, 2))
    f = 2 if ds else 1
    sw, sh = int(rng.integers(1, 40)) * 4 * f, int(rng.integers(1, 30)) * 4 * f
    w, h = (sw // f, sh // f) if rotation % 180 == 0 else (sh // f, sw // f)
    y_stride, c_stride = sw + int(rng.choice([0, 4, 16, 33])), sw + int(rng.choice([0, 2, 8, 64]))
    ybuf = rng.integers(0, 256, y_stride * sh + 64).astype(np.uint8)
    cbuf = rng.integers(0, 256, c_stride * (sh // 2) + 64).astype(np.uint8)
    u_first = int(rng.integers(0, 2))
    a = np.zeros(w * h * 3 // 2 + 64, np.uint8)
    b = np.zeros_like(a)
    na = L.orc_nv12_to_i420(ptr(ybuf), ptr(cbuf), rotation, w, h, y_stride, c_stride, u_first, ds, ptr(a))
    nb = R.ref_nv12_to_i420(ptr(ybuf), ptr(cbuf), rotation, w, h, y_stride, c_stride, u_first, ds, ptr(b))
    assert na == nb == w * h * 3 // 2
    assert np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(12))
def test_equalizer_oracle_random_gain_tables(seed):
    """random rate (nfft 128 / 256 / 512), up to four random gain points (0.1 .. 4, widths 50 .. 1000 Hz), random multi-tone
    input with noise: bit-exact against the reference filter"""
    L = O.oracle()
    rng = np.random.default_rng(seed)
    rate, T = int(rng.choice([8000, 16000, 32000, 48000])), 12
    n = rate // 100
    t = np.arange(T * n)
    x = np.zeros(T * n)
    for _ in range(int(rng.integers(1, 5))):
        x += float(rng.uniform(500, 6000)) * np.sin(2 * np.pi * float(rng.uniform(50, rate / 2 - 100)) * t / rate + float(rng.uniform(0, 6)))
    x += rng.normal(0, float(rng.choice([0, 100, 1500])), T * n)
    x = x.clip(-32768, 32767).astype(np.int16)
    gains = [(int(rng.uniform(50, rate / 2 - 50)), float(rng.choice([0.1, 0.4, 1.0, 2.0, 4.0])), int(rng.choice([50, 200, 400, 1000])))
             for _ in range(int(rng.integers(0, 5)))]
    g = RefGraph()
    eq = g.new("MSEqualizer")
    g.call_int(eq, "MS_FILTER_SET_SAMPLE_RATE", rate)
    for f, gn, w in gains:
        g.call(eq, "MS_EQUALIZER_SET_GAIN", EqualizerGain(f, gn, w))
    src, sink = g.source(x, n * 2), g.sink()
    g.link(src, 0, eq, 0)
    g.link(eq, 0, sink, 0)
    g.run(src, T)
    y_ref, _ = g.read(sink)
    g.close()
    e = L.orc_equalizer_new(rate)
    for f, gn, w in gains:
        L.orc_equalizer_set_gain(e, f, gn, w)
    y = x.copy()
    for k in range(T):
        L.orc_equalizer_process(e, ptr(y[k * n:(k + 1) * n]), n)
    L.orc_equalizer_free(e)
    assert np.array_equal(y, y_ref)


@pytest.mark.parametrize("seed", range(16))
def test_volume_chunked_mode_oracle_random_configs(seed):
    """AGC and / or echo limiter with random rate, gain, thresholds, speed, force, sustain, loudness, and microphone blocks
    of random length arriving on random ticks (0 .. 2 ticks apart, several per tick): oracle == reference filters"""
    L = O.oracle()
    rng = np.random.default_rng(seed)
    rate, T = int(rng.choice([8000, 16000, 48000])), 60
    n = rate // 100
    agc, peer = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    if not agc and not peer:
        agc = True
    spk = _speechy(seed * 2 + 1, T * n, rate, float(rng.choice([500, 3000, 9000, 25000])))
    mic = _speechy(seed * 2 + 2, T * n, rate, float(rng.choice([300, 2000, 5000, 20000])))
    gain = float(rng.choice([0.5, 1.0, 1.5, 3.0]))
    ea = (float(rng.choice([0.01, 0.05, 0.2])), float(rng.choice([0.1, 0.3, 0.5])), float(rng.choice([2.0, 6.0, 15.0])),
          int(rng.choice([50, 100, 400])))
    blk = int(rng.choice([2 * n // 3, n, n // 2, 3 * n // 2]))
    g = RefGraph()
    vspk, vmic = g.new("MSVolume"), g.new("MSVolume")
    for v in (vspk, vmic):
        g.call_int(v, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_float(vmic, "MS_VOLUME_SET_GAIN", gain)
    if agc:
        g.call_int(vmic, "MS_VOLUME_ENABLE_AGC", 1)
    if peer:
        g.call_ptr(vmic, "MS_VOLUME_SET_PEER", vspk)
        assert g.call_float(vmic, "MS_VOLUME_SET_EA_THRESHOLD", ea[0]) == 0
        assert g.call_float(vmic, "MS_VOLUME_SET_EA_SPEED", ea[1]) == 0
        assert g.call_float(vmic, "MS_VOLUME_SET_EA_SPEED", 0.9) == -1  # out of [0, 0.5]: refused, value kept (:324-333)
        g.call_float(vmic, "MS_VOLUME_SET_EA_FORCE", ea[2])
        g.call_int(vmic, "MS_VOLUME_SET_EA_SUSTAIN", ea[3])
    s_spk, k_spk = g.source(spk, n * 2), g.sink()
    s_mic, k_mic = g.source(), g.sink()
    sched, pos, tick = {}, 0, 0
    while pos < len(mic):
        g.push(s_mic, tick, mic[pos:pos + blk])
        sched.setdefault(tick, []).append((pos, min(pos + blk, len(mic))))
        pos += blk
        tick += int(rng.integers(0, 3))
    g.link(s_spk, 0, vspk, 0)
    g.link(vspk, 0, k_spk, 0)
    g.link(s_mic, 0, vmic, 0)
    g.link(vmic, 0, k_mic, 0)
    total = max(T, tick) + 5
    g.run([s_spk, s_mic], total)
    y_spk, _ = g.read(k_spk)
    y_mic, _ = g.read(k_mic)
    g.close()
    st_spk, st_mic = OrcVolumeState(), OrcVolumeState()
    L.orc_volume_init(C.byref(st_spk), rate)
    L.orc_volume_init(C.byref(st_mic), rate)
    st_mic.gain = st_mic.target_gain = st_mic.static_gain = gain
    st_mic.agc_enabled = int(agc)
    if peer:
        st_mic.ea_thres, st_mic.vol_upramp, st_mic.force, st_mic.sustain_time = ea
    buf, out_mic, exp_spk = np.zeros(0, np.int16), [], spk.copy()
    for t in range(total):
        if t < T:
            L.orc_volume_process(C.byref(st_spk), ptr(exp_spk[t * n:(t + 1) * n]), n)
        for a, b in sched.get(t, []):
            buf = np.concatenate([buf, mic[a:b]])
        while len(buf) >= n:
            chunk = np.ascontiguousarray(buf[:n])
            buf = buf[n:]
            pe = C.c_float(st_spk.energy)
            L.orc_volume_process_chunk(C.byref(st_mic), C.byref(pe) if peer else None, ptr(chunk), n)
            out_mic.append(chunk)
    out_mic = np.concatenate(out_mic) if out_mic else np.zeros(0, np.int16)
    assert np.array_equal(y_spk, exp_spk)
    assert len(y_mic) == len(out_mic) and np.array_equal(y_mic, out_mic)
