"""The drop-in claim, end to end: the SAME graph script runs in an unmodified MSTicker once with the reference's own
filters and once with the plugin's B200 filters (picked by name through the reference factory); outputs are compared.
mixer / volume / channel adapter / equalizer: bit-exact; resampler and echo canceller (their
reference arithmetic lives in the absent speexdsp): bit-exact / <= 2 LSB against the oracle."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
from _oracle import EqualizerGain, MixerCtl, RefGraph, ptr
from chain_reference import OracleChain
from synth import cfg2_stream

pytestmark = pytest.mark.gpu


def _graphs():
    if not (O.PLUGIN_DIR / "libmsb200filters.so").exists():
        pytest.fail("plugin/lib/libmsb200filters.so missing on the GPU box (it must travel with the snapshot)")
    return RefGraph(), RefGraph(plugins_dir=str(O.PLUGIN_DIR))


def noise(seed, n, amp):
    return np.random.default_rng(seed).integers(-amp, amp + 1, size=n).astype(np.int16)


def _run_mixer(g, pcm, rate, conf):
    P = pcm.shape[0]
    nwords = rate // 100
    T = pcm.shape[1] // nwords
    mix = g.new("MSAudioMixer")
    g.call_int(mix, "MS_FILTER_SET_SAMPLE_RATE", rate)
    g.call_int(mix, "MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE", int(conf))
    ctl = MixerCtl(pin=3)
    ctl.param.gain = 0.5
    g.call(mix, "MS_AUDIO_MIXER_SET_INPUT_GAIN", ctl)
    ctl = MixerCtl(pin=7)
    ctl.param.active = 0
    g.call(mix, "MS_AUDIO_MIXER_SET_ACTIVE", ctl)
    srcs, sinks = [], []
    for p in range(P):
        s = g.source()
        for t in range(T):
            if p == 5 and t in (4, 5):
                continue  # a starving pin
            g.push(s, t, pcm[p, t * nwords:(t + 1) * nwords])
        k = g.sink()
        g.link(s, 0, mix, p)
        g.link(mix, p, k, 0)
        srcs.append(s)
        sinks.append(k)
    g.run(srcs[0], T + 1)
    outs = [g.read(k)[0] for k in sinks]
    text = g.text(mix)
    g.close()
    return outs, text


@pytest.mark.parametrize("conf", [True, False])
def test_mixer_plugin_bit_exact_vs_reference_filter_in_ticker(conf):
    ref_g, b200_g = _graphs()
    rate, P, T = 48000, 16, 10
    n = rate // 100
    pcm = np.stack([noise(p, T * n, 7000) for p in range(P)])
    pcm[:, 2 * n:3 * n] = np.where(pcm[:, 2 * n:3 * n] > 0, 30000, -30000)
    a, ta = _run_mixer(ref_g, pcm, rate, conf)
    b, tb = _run_mixer(b200_g, pcm, rate, conf)
    assert not ta.startswith("B200:") and tb.startswith("B200:")
    for p in range(P):
        assert len(a[p]) == len(b[p]) and np.array_equal(a[p], b[p]), p


def test_mixer_plugin_bypass_mode_matches_reference():
    """a silent second pin stops counting as active after BYPASS_MODE_TIMEOUT (1000 ms of ticker time): from then on the
    single active input's packets are forwarded untouched (audiomixer.c:244-286). 130 ticks cover both regimes."""
    outs = []
    for g in _graphs():
        mix = g.new("MSAudioMixer")
        g.call_int(mix, "MS_FILTER_SET_SAMPLE_RATE", 16000)
        g.call_int(mix, "MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE", 1)
        x = noise(3, 160 * 130, 9000)
        s0, s1 = g.source(x, 320), g.source()
        k0, k1 = g.sink(), g.sink()
        g.link(s0, 0, mix, 0)
        g.link(s1, 0, mix, 1)
        g.link(mix, 0, k0, 0)
        g.link(mix, 1, k1, 0)
        g.run(s0, 130)
        outs.append((g.read(k0)[0], g.read(k1)[0]))
        g.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert len(outs[0][1]) == 160 * 130          # pin 1 hears pin 0 throughout
    assert 0 < len(outs[0][0]) < 160 * 130      # pin 0 gets (silent) mixes only until bypass mode starts


def test_volume_and_chanadapt_plugin_bit_exact_vs_reference():
    rate, T = 48000, 20
    n = rate // 100
    x = (noise(9, T * n, 12000).astype(np.int32) + 900).clip(-32768, 32767).astype(np.int16)
    x[3 * n:4 * n] = 32767
    res = []
    for g in _graphs():
        vol = g.new("MSVolume")
        g.call_int(vol, "MS_FILTER_SET_SAMPLE_RATE", rate)
        g.call_float(vol, "MS_VOLUME_SET_GAIN", 0.8)
        g.call_int(vol, "MS_VOLUME_REMOVE_DC", 1)
        ad = g.new("MSChannelAdapter")
        g.call_int(ad, "MS_FILTER_SET_SAMPLE_RATE", rate)
        g.call_int(ad, "MS_FILTER_SET_NCHANNELS", 1)
        g.call_int(ad, "MS_CHANNEL_ADAPTER_SET_OUTPUT_NCHANNELS", 2)
        src, sink = g.source(x, n * 2), g.sink()
        g.link(src, 0, vol, 0)
        g.link(vol, 0, ad, 0)
        g.link(ad, 0, sink, 0)
        g.run(src, T)
        lin = C.c_float()
        g.call(vol, "MS_VOLUME_GET_LINEAR", lin)
        res.append((g.read(sink)[0], lin.value))
        g.close()
    assert len(res[0][0]) == T * n * 2
    assert np.array_equal(res[0][0], res[1][0])
    assert np.float32(res[0][1]) == np.float32(res[1][1])


def test_equalizer_plugin_equals_reference_filter():
    rate, T = 16000, 12
    n = rate // 100
    t = np.arange(T * n)
    x = (6000 * np.sin(2 * np.pi * 300 * t / rate) + 5000 * np.sin(2 * np.pi * 1000 * t / rate)).astype(np.int16)
    res = []
    for g in _graphs():
        eq = g.new("MSEqualizer")
        g.call_int(eq, "MS_FILTER_SET_SAMPLE_RATE", rate)
        g.call(eq, "MS_EQUALIZER_SET_GAIN", EqualizerGain(1000, 2.0, 200))
        src, sink = g.source(x, n * 2), g.sink()
        g.link(src, 0, eq, 0)
        g.link(eq, 0, sink, 0)
        g.run(src, T)
        res.append(g.read(sink)[0])
        g.close()
    assert len(res[0]) == len(res[1]) == T * n
    assert np.array_equal(res[0], res[1])


def test_resample_plugin_cfg1_matches_oracle_and_stamps_timestamps():
    """BASELINE cfg1: one 8 kHz stream -> 48 kHz through the plugin's MSResample in the ticker."""
    L = O.oracle()
    _, g = _graphs()
    T = 30
    x = (8000 * np.sin(2 * np.pi * 440 * np.arange(80 * T) / 8000)).astype(np.int16)
    rs = g.new("MSResample")
    g.call_int(rs, "MS_FILTER_SET_SAMPLE_RATE", 8000)
    g.call_int(rs, "MS_FILTER_SET_OUTPUT_SAMPLE_RATE", 48000)
    src, sink = g.source(x, 160), g.sink()
    g.link(src, 0, rs, 0)
    g.link(rs, 0, sink, 0)
    g.run(src, T)
    y, tri = g.read(sink)
    g.close()
    o = L.orc_resampler_new(1, 8000, 48000, 3)
    exp = []
    for k in range(T):
        out = np.zeros(488, np.int16)
        m = L.orc_msresample_block(o, ptr(x[k * 80:(k + 1) * 80]), 80, ptr(out))
        exp.append(out[:m])
    exp = np.concatenate(exp)
    assert np.array_equal(y, exp)
    assert list(tri[:, 1]) == [960] * T            # 480 samples per 10 ms block
    assert list(tri[:, 2]) == [480 * k for k in range(T)]  # running output-sample timestamp (msresample.c:168-169)


def test_speexec_plugin_graph_matches_oracle_chain():
    """ref/mic -> MSResample x2 -> MSSpeexEC -> MSVolume in the unmodified ticker (audiostream.c:1798-1832 shape) vs the
    CPU composition of the oracle pieces."""
    _, g = _graphs()
    T = 40
    ref16, mic16, _, _ = cfg2_stream(77, 160 * T, 16000)
    r1, r2 = g.new("MSResample"), g.new("MSResample")
    for r in (r1, r2):
        g.call_int(r, "MS_FILTER_SET_SAMPLE_RATE", 16000)
        g.call_int(r, "MS_FILTER_SET_OUTPUT_SAMPLE_RATE", 48000)
    ec = g.new("MSSpeexEC")
    g.call_int(ec, "MS_FILTER_SET_SAMPLE_RATE", 48000)
    vol = g.new("MSVolume")
    g.call_int(vol, "MS_FILTER_SET_SAMPLE_RATE", 48000)
    g.call_float(vol, "MS_VOLUME_SET_GAIN", 0.8)
    s_ref, s_mic = g.source(ref16, 320), g.source(mic16, 320)
    k_ref, k_out = g.sink(), g.sink()
    g.link(s_ref, 0, r1, 0)
    g.link(s_mic, 0, r2, 0)
    g.link(r1, 0, ec, 0)   # pin 0: far-end reference
    g.link(r2, 0, ec, 1)   # pin 1: microphone
    g.link(ec, 0, k_ref, 0)
    g.link(ec, 1, vol, 0)
    g.link(vol, 0, k_out, 0)
    g.run(s_ref, T)
    # attach the second source's graph too (same ticker): both sources belong to one connected graph already
    y, tri = g.read(k_out)
    g.close()
    # oracle: the EC starts consuming the reference only once echo has started (speexec.c:239-250): the first tick's
    # reference block is flushed, so the reference is delayed by one tick w.r.t. the lockstep chain -> build that here
    oc = OracleChain(1, 16000, 48000, 250, 0.8, 0)
    exp = []
    zero = np.zeros(160, np.int16)
    for t in range(T):
        exp.append(_oracle_tick_with_ec_start(oc, ref16[t * 160:(t + 1) * 160], mic16[t * 160:(t + 1) * 160], drop_ref=(t == 0)))
    exp = np.concatenate(exp)
    assert len(y) == len(exp)
    assert np.abs(y.astype(np.int32) - exp.astype(np.int32)).max() <= 2


def _oracle_tick_with_ec_start(oc, ref_in, mic_in, drop_ref):
    """OracleChain.tick_stream with the reference's start-up rule: far-end blocks arriving before the first microphone
    frame are dropped, later underruns are filled with zeros (speexec.c:239-272)."""
    L = oc.L
    o2 = np.zeros(oc.tick + 8, np.int16)
    n2 = L.orc_msresample_block(oc.rs_mic[0], ptr(np.ascontiguousarray(mic_in)), oc.tick_in, ptr(o2))
    # the far-end resampler sits before the EC: it runs on every block, also on the one the EC drops at start-up
    o1 = np.zeros(oc.tick + 8, np.int16)
    n1 = L.orc_msresample_block(oc.rs_ref[0], ptr(np.ascontiguousarray(ref_in)), oc.tick_in, ptr(o1))
    if not drop_ref:
        oc.buf_ref[0] = np.concatenate([oc.buf_ref[0], o1[:n1]])
    oc.buf_mic[0] = np.concatenate([oc.buf_mic[0], o2[:n2]])
    outs = []
    F = oc.F
    while len(oc.buf_mic[0]) >= F:
        mic = np.ascontiguousarray(oc.buf_mic[0][:F])
        oc.buf_mic[0] = oc.buf_mic[0][F:]
        if len(oc.buf_ref[0]) < F:  # underrun: a frame of zeros is appended to the delayed reference (speexec.c:261-272),
            oc.buf_ref[0] = np.concatenate([oc.buf_ref[0], np.zeros(F, np.int16)])  # then the OLDEST frame is read
        ref = np.ascontiguousarray(oc.buf_ref[0][:F])
        oc.buf_ref[0] = oc.buf_ref[0][F:]
        out = np.zeros(F, np.int16)
        L.orc_aec_process_frame(oc.aec[0], ptr(mic), ptr(ref), ptr(out))
        L.orc_volume_process(C.byref(oc.vol[0]), ptr(out), F)
        outs.append(out)
    return np.concatenate(outs) if outs else np.zeros(0, np.int16)


def test_volume_plugin_chunked_mode_with_peer_bit_exact_vs_reference():
    """echo-limiter peer + AGC (msvolume.c:480-502) through the plugin in the unmodified ticker vs the reference filters."""
    rate, T = 16000, 50
    n = rate // 100
    rng = np.random.default_rng(5)
    spk = np.concatenate([(rng.standard_normal(n) * (9000 if (k // 6) % 2 == 0 else 40)).astype(np.int16) for k in range(T)])
    mic = (rng.standard_normal(T * n) * 3000).astype(np.int16)
    res = []
    for g in _graphs():
        vspk, vmic = g.new("MSVolume"), g.new("MSVolume")
        for v in (vspk, vmic):
            g.call_int(v, "MS_FILTER_SET_SAMPLE_RATE", rate)
        g.call_float(vmic, "MS_VOLUME_SET_GAIN", 1.5)
        g.call_int(vmic, "MS_VOLUME_ENABLE_AGC", 1)
        g.call_ptr(vmic, "MS_VOLUME_SET_PEER", vspk)
        g.call_float(vmic, "MS_VOLUME_SET_EA_THRESHOLD", 0.05)
        g.call_float(vmic, "MS_VOLUME_SET_EA_FORCE", 6.0)
        s_spk, k_spk = g.source(spk, n * 2), g.sink()
        s_mic, k_mic = g.source(mic, 2 * n * 2 // 2), g.sink()
        g.link(s_spk, 0, vspk, 0)
        g.link(vspk, 0, k_spk, 0)
        g.link(s_mic, 0, vmic, 0)
        g.link(vmic, 0, k_mic, 0)
        g.run([s_spk, s_mic], T)
        res.append((g.read(k_spk)[0], g.read(k_mic)[0]))
        g.close()
    assert len(res[0][1]) == T * n
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])


def test_volume_plugin_setters_called_mid_stream_bit_exact_vs_reference():
    """the settings that reset the gain ramp do so when THEY are called and only then (msvolume.c:262-276, 352-378): noise
    gate switched on, an unrelated setting changed, the gate switched off, a gain in dB — between runs of the same graph"""
    rate, T = 16000, 12
    n = rate // 100
    rng = np.random.default_rng(11)
    x = np.concatenate([(rng.standard_normal(n) * (6000 if (k // 5) % 2 == 0 else 30)).astype(np.int16) for k in range(5 * T)])
    res = []
    for g in _graphs():
        vol = g.new("MSVolume")
        g.call_int(vol, "MS_FILTER_SET_SAMPLE_RATE", rate)
        g.call_float(vol, "MS_VOLUME_SET_GAIN", 1.3)
        src, sink = g.source(x, n * 2), g.sink()
        g.link(src, 0, vol, 0)
        g.link(vol, 0, sink, 0)
        g.run(src, T)
        g.call_float(vol, "MS_VOLUME_SET_NOISE_GATE_FLOORGAIN", 0.1)
        g.call_float(vol, "MS_VOLUME_SET_NOISE_GATE_THRESHOLD", 0.05)
        g.call(vol, "MS_VOLUME_ENABLE_NOISE_GATE", C.c_ubyte(1))
        g.run(src, T)
        g.call_int(vol, "MS_VOLUME_REMOVE_DC", 1)  # unrelated: must not replay the gate's gain reset
        g.run(src, T)
        g.call(vol, "MS_VOLUME_ENABLE_NOISE_GATE", C.c_ubyte(0))
        g.run(src, T)
        g.call_float(vol, "MS_VOLUME_SET_DB_GAIN", -3.0)  # gain = static gain, target gain untouched
        g.run(src, T)
        lin = C.c_float()
        g.call(vol, "MS_VOLUME_GET_LINEAR", lin)
        res.append((g.read(sink)[0], lin.value))
        g.close()
    assert len(res[0][0]) == 5 * T * n
    assert np.array_equal(res[0][0], res[1][0])
    assert np.float32(res[0][1]) == np.float32(res[1][1])


@pytest.mark.parametrize("name,ptime", [("MSAlaw", 0), ("MSUlaw", 0), ("MSAlaw", 30), ("MSUlaw", 10)])
def test_plugin_g711_codecs_bit_exact_vs_reference_filters(name, ptime):
    """source -> <law>Enc -> <law>Dec -> sink in the unmodified MSTicker: the plugin's filters (GPU companding, host
    re-framing) give the same blocks — sizes, timestamps and samples — as the reference's filters"""
    import ctypes as C
    rng = np.random.default_rng(17 + ptime)
    n, ticks = 80, 30
    pcm = (rng.standard_normal(n * ticks) * 9000).clip(-32768, 32767).astype(np.int16)
    pcm[:16] = [-32768, 32767, 0, -1, 1, -8, 8, -9, 255, 256, -256, -257, 4095, 4096, -4096, -4097]

    def run(plugins_dir):
        g = RefGraph(plugins_dir=plugins_dir)
        src, enc, dec, sink_c, sink_p = g.source(pcm, n * 2), g.new(name + "Enc"), g.new(name + "Dec"), g.sink(), g.sink()
        if plugins_dir:
            assert g.text(enc).startswith("B200:") and g.text(dec).startswith("B200:")
        if ptime:
            assert g.call_ptr(enc, "MS_FILTER_ADD_FMTP", C.c_char_p(b"ptime=%d" % ptime)) == 0
        g.link(src, 0, enc, 0)
        g.link(enc, 0, dec, 0)
        g.link(dec, 0, sink_p, 0)
        g.run(src, ticks + 2)
        out, tri = g.read(sink_p)
        g.close()
        return out, tri

    ref_out, ref_tri = run(None)
    got_out, got_tri = run(str(O.PLUGIN_DIR))
    assert len(ref_out) > n * (ticks - 4)
    assert np.array_equal(got_tri, ref_tri)  # (tick, bytes, timestamp) of every block
    assert np.array_equal(got_out, ref_out)


@pytest.mark.parametrize("strategy,drop_ms,interval_ms", [(1, 30, 200), (0, 40, 300), (1, 120, 150)])
def test_plugin_flowcontrol_bit_exact_vs_reference_filter(strategy, drop_ms, interval_ms):
    """source -> MSAudioFlowControl -> sink with a drop request after two ticks: the plugin's filter forwards the same
    blocks (sizes) and samples as the reference's"""
    from _oracle import FlowControlConfig, FlowControlDropEvent
    from test_oracle_vs_reference import _flowctl_signal
    rate, n, ticks = 16000, 160, 40
    x = _flowctl_signal(np.random.default_rng(drop_ms), n, ticks)

    def run(plugins_dir):
        g = RefGraph(plugins_dir=plugins_dir)
        src, fc, sink = g.source(x, n * 2), g.new("MSAudioFlowControl"), g.sink()
        assert g.text(fc).startswith("B200:") == bool(plugins_dir)
        assert g.call_int(fc, "MS_FILTER_SET_SAMPLE_RATE", rate) == 0
        assert g.call_int(fc, "MS_FILTER_SET_NCHANNELS", 1) == 0
        assert g.call(fc, "MS_AUDIO_FLOW_CONTROL_SET_CONFIG", FlowControlConfig(strategy, 0.02)) == 0
        g.link(src, 0, fc, 0)
        g.link(fc, 0, sink, 0)
        g.run(src, 2)
        assert g.call(fc, "MS_AUDIO_FLOW_CONTROL_DROP", FlowControlDropEvent(interval_ms, drop_ms)) == 0
        g.run(src, ticks)
        out, tri = g.read(sink)
        g.close()
        return out, tri

    ref_out, ref_tri = run(None)
    got_out, got_tri = run(str(O.PLUGIN_DIR))
    assert len(ref_out) < len(x)
    assert np.array_equal(got_tri[:, 1], ref_tri[:, 1])
    assert np.array_equal(got_out, ref_out)


@pytest.mark.parametrize("rate,block_ms,cn_at", [(16000, 10, ()), (48000, 10, ()), (8000, 20, ()), (16000, 10, (12,))])
def test_plugin_generic_plc_bit_exact_vs_reference_filter(rate, block_ms, cn_at):
    """source (with holes) -> MSGenericPLC -> sink: the plugin's filter (host concealer clock + msb200_plc_* on the GPU)
    emits the same blocks as the unmodified reference filter: delayed, concealed, faded, comfort-noise silence"""
    from test_oracle_vs_reference import plc_reference_run, plc_schedule, plc_signal
    ticks = 70
    lost = set(range(10, 14)) | {20} | set(range(30, 52)) | {60, 61}
    if block_ms != 10:
        lost = {5, 6, 12} | set(range(20, 31))
    x = plc_signal(rate, ticks * rate // 100, seed=rate // 1000)
    sched = plc_schedule(rate, ticks, lost, block_ms)
    ref_out, ref_blocks = plc_reference_run(rate, ticks, sched, x, cn_at)
    got_out, got_blocks = plc_reference_run(rate, ticks, sched, x, cn_at, plugins_dir=str(O.PLUGIN_DIR))
    assert got_blocks == ref_blocks
    assert np.array_equal(got_out, ref_out)
