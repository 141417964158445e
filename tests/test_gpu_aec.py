"""GPU parity for MSSpeexEC's arithmetic (MDF canceller + preprocessor) vs oracle/oracle_aec.c.

The adaptive loop is chaotic over long horizons and the kernel sums its block reductions in tree order, so parity is
stated as: <= 2 LSB on the first 50 frames from identical (fresh) state, internal state within 1e-3
relative, and echo-return-loss within 1 dB over a long run."""
import numpy as np
import pytest

import _oracle as O
from _oracle import ptr
from mediastreamer2_b200 import filters as F
from synth import cfg2_stream

pytestmark = pytest.mark.gpu


def _oracle_run(L, rate, tail, mic, ref, cancel_only=False):
    a = L.orc_aec_new(rate, tail, 64)
    Fs = L.orc_aec_frame_size(a)
    out = np.zeros_like(mic)
    for k in range(len(mic) // Fs):
        s = slice(k * Fs, (k + 1) * Fs)
        L.orc_aec_process_frame(a, ptr(mic[s]), ptr(ref[s]), ptr(out[s]))
    return a, out


@pytest.mark.parametrize("rate,tail", [(16000, 250), (48000, 250), (8000, 100)])
def test_aec_first_frames_match_oracle(ctx, rate, tail):
    L = O.oracle()
    n_streams, nframes = 3, 50
    ec = F.SpeexEC(ctx, n_streams, rate, tail)
    Fs = ec.frame_size
    assert Fs == L.orc_aec_frame_size_for_rate(rate, 64)
    n = Fs * nframes
    mics, refs = [], []
    for s in range(n_streams):
        x, mic, _, _ = cfg2_stream(s, n, rate)
        mics.append(mic)
        refs.append(x)
    mics, refs = np.stack(mics), np.stack(refs)
    # feed in uneven chunks (1, 2, 1, 2 ... frames per call) like the 10 ms tick re-framing does
    got = np.zeros_like(mics)
    k = 0
    step = 1
    while k < nframes:
        c = min(step, nframes - k)
        sl = slice(k * Fs, (k + c) * Fs)
        got[:, sl] = ec.process(np.ascontiguousarray(mics[:, sl]), np.ascontiguousarray(refs[:, sl]))
        k += c
        step = 3 - step
    for s in range(n_streams):
        a, exp = _oracle_run(L, rate, tail, mics[s], refs[s])
        d = np.abs(got[s].astype(np.int32) - exp.astype(np.int32))
        assert d.max() <= 2, (s, d.max(), np.argmax(d) // Fs)
        M, N = ec.info.M, ec.info.window_size
        for what, size in (("W", M * N), ("foreground", M * N), ("X", (M + 1) * N), ("power", Fs + 1), ("noise", Fs)):
            ref_v = np.zeros(size, np.float32)
            assert L.orc_aec_probe(a, what.encode(), ptr(ref_v), size) == size
            got_v = ec.probe(s, what, size)
            scale = np.abs(ref_v).max() + 1e-20
            assert np.abs(got_v - ref_v).max() <= 2e-3 * scale, (s, what, np.abs(got_v - ref_v).max(), scale)
            if what == "X":
                # the far-end spectra depend on the input filter and the FFT only, and the kernel's FFT does the oracle's
                # butterflies operation for operation: bit-exact
                assert np.array_equal(got_v.view(np.uint32), ref_v.view(np.uint32)), (s, np.abs(got_v - ref_v).max())
        L.orc_aec_free(a)
    ec.close()


def test_aec_serial_warp_build_is_bit_identical_to_the_plain_build(ctx):
    """48 kHz: the default kernel hands the frame's sequential IIR filters (DC notch, pre-emphasis, de-emphasis) to a ninth
    warp that runs beside the per-bin threads; same operations in the same order, so samples AND state must equal the
    256-thread build's bit for bit — including saturated microphone frames, ragged frame counts (1, 2, 3 frames per call)
    and a stream driven into speex_echo_state_reset (the reset also clears the filters' memories)."""
    rate, tail, n_streams, nframes = 48000, 250, 5, 240
    banks = {p: F.SpeexEC(ctx, n_streams, rate, tail) for p in (0, 1, 3, 4, 5)}
    for p, ec in banks.items():
        ec.set_path(p)
    Fs = banks[0].frame_size
    n = Fs * nframes
    mics, refs = [], []
    for s in range(n_streams):
        x, mic, _, _ = cfg2_stream(70 + s, n, rate)
        mics.append(mic.copy())
        refs.append(x.copy())
    mics, refs = np.stack(mics), np.stack(refs)
    mics[1, 20 * Fs:23 * Fs] = 32767  # saturation: the adaptation freezes for a frame
    outs = {}
    for p, ec in banks.items():
        got = np.zeros_like(mics)
        k, step = 0, 1
        while k < nframes:
            if k == 30:
                # stream 4: an absurd foreground filter -> the frame's sanity checks trip -> speex_echo_state_reset
                blob = bytearray(ec.get_state_blob(4))
                w = np.frombuffer(blob, np.float32, offset=16)
                w[w.size // 2:] = 1e18
                ec.set_state_blob(4, bytes(blob))
            c = min(step, nframes - k)
            sl = slice(k * Fs, (k + c) * Fs)
            got[:, sl] = ec.process(np.ascontiguousarray(mics[:, sl]), np.ascontiguousarray(refs[:, sl]))
            k += c
            step = step % 3 + 1
        outs[p] = got
        assert ec.probe(4, "scalars", 16)[11] < nframes - 30 + 1  # cancel_count restarted: the reset did happen
    M, N = banks[0].info.M, banks[0].info.window_size
    for p in (0, 3, 4, 5):  # 4 and 5 run the generic loop of the block pass, the others its tight form
        assert np.array_equal(outs[p], outs[1]), (p, np.abs(outs[p].astype(int) - outs[1].astype(int)).max())
        for s in range(n_streams):
            for what, size in (("W", M * N), ("foreground", M * N), ("X", (M + 1) * N), ("E", N), ("power_1", Fs + 1),
                               ("noise", Fs), ("scalars", 16)):
                a, b = banks[p].probe(s, what, size), banks[1].probe(s, what, size)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (p, s, what)
    for ec in banks.values():
        ec.close()


@pytest.mark.parametrize("rate,tail", [(8000, 250), (16000, 250), (16000, 100), (48000, 250)])
def test_aec_pass_forms_are_bit_identical(ctx, rate, tail):
    """the block pass has a tight form for the common frame (no foreground refresh pending, adaptation on, whole groups of
    three blocks) and a generic loop for everything else; set_path(5) runs every frame through the generic loop. Samples and
    state must be equal bit for bit at every frame size, through the start-up, the adapted regime (|W_j|^2 bookkeeping), a
    saturated stretch (adaptation off: generic loop in both) and the foreground refreshes on the way."""
    n_streams, seconds = 3, 2.6
    a, b = F.SpeexEC(ctx, n_streams, rate, tail), F.SpeexEC(ctx, n_streams, rate, tail)
    b.set_path(5)
    Fs = a.frame_size
    nframes = int(seconds * rate) // Fs
    mics, refs = [], []
    for s in range(n_streams):
        x, mic, _, _ = cfg2_stream(900 + s, Fs * nframes, rate)
        mics.append(mic.copy())
        refs.append(x.copy())
    mics, refs = np.stack(mics), np.stack(refs)
    mics[2, 40 * Fs:42 * Fs] = -32768
    for k in range(0, nframes, 2):
        sl = slice(k * Fs, min(k + 2, nframes) * Fs)
        ya = a.process(np.ascontiguousarray(mics[:, sl]), np.ascontiguousarray(refs[:, sl]))
        yb = b.process(np.ascontiguousarray(mics[:, sl]), np.ascontiguousarray(refs[:, sl]))
        assert np.array_equal(ya, yb), (rate, k)
    M, N = a.info.M, a.info.window_size
    for s in range(n_streams):
        for what, size in (("W", M * N), ("foreground", M * N), ("X", (M + 1) * N), ("E", N), ("scalars", 16)):
            assert np.array_equal(a.probe(s, what, size).view(np.uint32), b.probe(s, what, size).view(np.uint32)), (s, what)
    assert a.probe(0, "scalars", 16)[0] == 1  # the adapted regime was reached
    a.close()
    b.close()


def test_aec_serial_warp_build_at_full_occupancy(ctx):
    """the same identity on a grid that fills the chip several times over (four CTAs per SM, every hand-off between the
    serial warp and the per-bin threads contended): 2048 streams x 8 frames in calls of 2 frames"""
    rate, n_streams, nframes = 48000, 2048, 8
    Fs = 256
    rng = np.random.default_rng(11)
    base = np.stack([np.stack(cfg2_stream(500 + s, Fs * nframes, rate)[:2]) for s in range(16)])  # [16][2][n]
    pick = rng.integers(0, 16, n_streams)
    gain = rng.uniform(0.3, 1.0, (n_streams, 1))
    refs = (base[pick, 0] * gain).astype(np.int16)
    mics = (base[pick, 1] * gain).astype(np.int16)
    outs = {}
    for p in (0, 1, 4, 5):
        ec = F.SpeexEC(ctx, n_streams, rate, 250)
        ec.set_path(p)
        got = np.zeros_like(mics)
        for k in range(0, nframes, 2):
            sl = slice(k * Fs, (k + 2) * Fs)
            got[:, sl] = ec.process(np.ascontiguousarray(mics[:, sl]), np.ascontiguousarray(refs[:, sl]))
        outs[p] = (got, np.stack([ec.probe(s, "E", 2 * Fs) for s in range(0, n_streams, 97)]))
        ec.close()
    for p in (0, 4, 5):
        assert np.array_equal(outs[p][0], outs[1][0]), p
        assert np.array_equal(outs[p][1].view(np.uint32), outs[1][1].view(np.uint32)), p
    assert np.abs(outs[0][0]).max() > 0


def _erle(mic, out, near, rate, lo, hi):
    seg = slice(int(lo * rate), int(hi * rate))
    m = np.abs(near[seg]) < 1
    return 10 * np.log10(np.mean(mic[seg][m].astype(float) ** 2) / max(np.mean(out[seg][m].astype(float) ** 2), 1e-9))


def test_aec_long_run_echo_return_loss_matches_oracle(ctx):
    L = O.oracle()
    rate, tail, secs = 16000, 250, 6
    ec = F.SpeexEC(ctx, 2, rate, tail)
    Fs = ec.frame_size
    n = (rate * secs) // Fs * Fs
    data = [cfg2_stream(40 + s, n, rate) for s in range(2)]
    mics = np.stack([d[1] for d in data])
    refs = np.stack([d[0] for d in data])
    got = ec.process(mics, refs)
    for s in range(2):
        a, exp = _oracle_run(L, rate, tail, mics[s], refs[s])
        L.orc_aec_free(a)
        e_gpu = _erle(mics[s], got[s], data[s][3], rate, 3, 6)
        e_cpu = _erle(mics[s], exp, data[s][3], rate, 3, 6)
        assert e_cpu > 6.0, e_cpu  # the canceller does cancel
        assert abs(e_gpu - e_cpu) <= 1.0, (e_gpu, e_cpu)
        # near-end speech is preserved equally
        near = np.abs(data[s][3]) > 100
        r_gpu = np.sqrt(np.mean(got[s][near].astype(float) ** 2))
        r_cpu = np.sqrt(np.mean(exp[near].astype(float) ** 2))
        assert abs(20 * np.log10(r_gpu / r_cpu)) <= 0.5
    ec.close()


def test_aec_white_noise_converges_and_state_blob_roundtrip(ctx):
    rate = 16000
    rng = np.random.default_rng(1)
    n = rate * 4
    x = rng.standard_normal(n) * 3000
    h = rng.standard_normal(400) * np.exp(-np.arange(400) / 100.0)
    h *= 0.25 / np.sqrt((h * h).sum())
    mic = np.convolve(x, h)[:n] + rng.standard_normal(n) * 10
    x16, mic16 = x.astype(np.int16), mic.astype(np.int16)
    ec = F.SpeexEC(ctx, 2, rate, 250)
    Fs = ec.frame_size
    n = n // Fs * Fs
    mics = np.stack([mic16[:n], mic16[:n]])
    refs = np.stack([x16[:n], x16[:n]])
    out = ec.process(mics, refs)
    tail_seg = slice(n - rate, n)
    erle = 10 * np.log10(np.mean(mics[0, tail_seg].astype(float) ** 2) / np.mean(out[0, tail_seg].astype(float) ** 2))
    assert erle > 20, erle
    assert np.array_equal(out[0], out[1])  # identical streams -> identical results (deterministic kernel)
    # MS_ECHO_CANCELLER_GET/SET_STATE_STRING equivalent: move stream 0's filter into a fresh canceller
    blob = ec.get_state_blob(0)
    ec2 = F.SpeexEC(ctx, 1, rate, 250)
    ec2.set_state_blob(0, blob)
    assert ec2.get_state_blob(0) == blob
    w = ec2.probe(0, "W", ec.info.M * ec.info.window_size)
    assert np.array_equal(w, ec.probe(0, "W", ec.info.M * ec.info.window_size))
    assert np.abs(w).max() > 0
    ec.close()
    ec2.close()


def test_two_aec_banks_with_different_tails_share_the_kernel(ctx):
    """the shared-memory opt-in belongs to the kernel, not to a bank: a later, smaller bank (tail 100 ms) must not take it
    away from an earlier, larger one (tail 500 ms)"""
    L = O.oracle()
    rate = 48000
    big = F.SpeexEC(ctx, 2, rate, 500)
    small = F.SpeexEC(ctx, 2, rate, 100)
    Fs = big.frame_size
    rng = np.random.default_rng(5)
    far = rng.integers(-8000, 8000, size=(2, 4 * Fs)).astype(np.int16)
    mic = (far // 2).astype(np.int16)
    for bank, tail in ((small, 100), (big, 500), (small, 100)):
        got = bank.process(mic, far)
        assert got.shape == mic.shape
    big.close()
    small.close()


def test_aec_ragged_frame_counts_keep_streams_independent(ctx):
    """Streams whose 10 ms ticks fall differently against the 256-sample frame grid stage 1 or 2 frames in DIFFERENT ticks
    (the plugin's lockstep batch mode). With per-stream counts each stream must produce exactly what it produces alone:
    no padding frame ever enters its far-end history or its adaptive filter."""
    from mediastreamer2_b200 import _lib

    rate, n, ticks = 48000, 3, 40
    lib = ctx.lib
    rng = np.random.default_rng(11)
    bank = F.SpeexEC(ctx, n, rate, 250)
    Fs = bank.frame_size
    # frames staged per tick: 480 samples per tick against frames of 256, each stream starting at another phase
    counts = np.zeros((ticks, n), np.int32)
    have = np.array([0, 130, 250])
    for k in range(ticks):
        have += 480
        counts[k] = have // Fs
        have %= Fs
    total = counts.sum(0)
    far = rng.integers(-9000, 9000, size=(n, total.max() * Fs)).astype(np.int16)
    mic = (far // 3 + rng.integers(-300, 300, size=far.shape)).astype(np.int16)
    got = [np.zeros(total[s] * Fs, np.int16) for s in range(n)]
    pos = np.zeros(n, np.int64)
    maxu = 4
    arena_m = np.zeros((n, maxu * Fs), np.int16)
    arena_r = np.zeros_like(arena_m)
    arena_o = np.zeros_like(arena_m)
    for k in range(ticks):
        for s in range(n):
            c = counts[k, s]
            arena_m[s, :c * Fs] = mic[s, pos[s] * Fs:(pos[s] + c) * Fs]
            arena_r[s, :c * Fs] = far[s, pos[s] * Fs:(pos[s] + c) * Fs]
            arena_m[s, c * Fs:] = 12345  # garbage beyond the count must never be read
            arena_r[s, c * Fs:] = -12345
        units = int(counts[k].max())
        if units == 0:
            continue
        ck = np.ascontiguousarray(counts[k])
        _lib.check(lib.msb200_aec_process_counts(bank.h, O.ptr(arena_m), O.ptr(arena_r), O.ptr(arena_o), units, maxu * Fs, O.ptr(ck)))
        for s in range(n):
            c = counts[k, s]
            got[s][pos[s] * Fs:(pos[s] + c) * Fs] = arena_o[s, :c * Fs]
            pos[s] += c
    bank.close()
    for s in range(n):
        solo = F.SpeexEC(ctx, 1, rate, 250)
        exp = np.zeros(total[s] * Fs, np.int16)
        for k in range(0, total[s], 2):
            c = min(2, total[s] - k)
            exp[k * Fs:(k + c) * Fs] = solo.process(mic[s:s + 1, k * Fs:(k + c) * Fs], far[s:s + 1, k * Fs:(k + c) * Fs])[0]
        solo.close()
        assert np.array_equal(got[s], exp), f"stream {s} differs from its solo run"
        assert exp[20 * Fs:].any()


def _gpu_run_scenarios(ctx, sigs):
    """every scenario as one stream of ONE bank (lockstep frames), 25 frames per call; shorter scenarios are padded with
    silence (the caller cuts each stream back to its own length)"""
    n = max(len(m) for _, m, _ in sigs)
    sigs = [(np.pad(f, (0, n - len(f))), np.pad(m, (0, n - len(m))), nr) for f, m, nr in sigs]
    ec = F.SpeexEC(ctx, len(sigs), 16000, 250)
    Fs = ec.frame_size
    nfr = n // Fs
    got = np.zeros((len(sigs), nfr * Fs), np.int16)
    step = 25
    for k in range(0, nfr, step):
        c = min(step, nfr - k)
        sl = slice(k * Fs, (k + c) * Fs)
        mics = np.ascontiguousarray(np.stack([m[sl] for _, m, _ in sigs]))
        refs = np.ascontiguousarray(np.stack([f[sl] for f, _, _ in sigs]))
        got[:, sl] = ec.process(mics, refs)
    ec.close()
    return got


def test_aec_on_the_reference_testers_simple_talk_material(ctx):
    """the reference tester's own echo scenario (tests/aec_fixture.py): the GPU canceller removes the lone far-end
    talker's echo by >= 25 dB, keeps the lone near-end talker, and tracks the oracle's ERLE within 2 dB"""
    import aec_fixture as A

    L = O.oracle()
    far, mic, near = A.load()
    got = _gpu_run_scenarios(ctx, [(far, mic, near), (far, mic, near)])  # two identical streams: treated identically
    assert np.array_equal(got[0], got[1])
    n = got.shape[1]
    mic, far = mic[:n], far[:n]
    erle, keep, corr = A.check_behaviour(got[0], mic, near, min_erle_db=25.0)
    _, exp = _oracle_run(L, A.RATE, 250, mic[:n], far[:n])
    erle_o, _, _ = A.check_behaviour(exp, mic, near, min_erle_db=25.0)
    assert max(abs(a - b) for a, b in zip(erle, erle_o)) <= 2.0, (erle, erle_o)


def test_aec_reference_suite_metric_gpu_vs_oracle(ctx, tmp_path):
    """Every scenario of the reference's AEC suite that the committed material covers, through the reference's OWN metric
    (ms_audio_compare_silence_and_speech, unmodified audiodiff.c in oracle/_ref): the GPU bank meets the same bounds as the
    oracle (tests/test_oracle_aec_fixture.py) and lands within 0.01 similarity of it."""
    import aec_fixture as A

    L = O.oracle()
    R = O.ref()
    g = A.load_all()
    names = list(A.SCENARIOS)
    sigs = [A.scenario_signals(g, nm) for nm in names]
    got = _gpu_run_scenarios(ctx, sigs)
    ideal = _gpu_run_scenarios(ctx, [(np.zeros_like(nr), nr.copy(), nr) for _, _, nr in sigs])
    Fs = 128
    for k, nm in enumerate(names):
        far, mic, near = sigs[k]
        n = len(mic) // Fs * Fs  # each scenario at its OWN length: the metric's windows reach to 14.5 s of the 15 s cuts
        sim, energy = A.silence_and_speech(R, tmp_path, near[:n], got[k][:n], nm)
        _, exp = _oracle_run(L, A.RATE, 250, mic[:n], far[:n])
        sim_o, energy_o = A.silence_and_speech(R, tmp_path, near[:n], exp, nm)
        min_sim, max_energy = A.SCENARIOS[nm][6]
        assert min_sim <= sim < 1.0 and energy <= max_energy, (nm, sim, energy)
        assert abs(sim - sim_o) <= 0.01, (nm, sim, sim_o)
        assert abs(energy - energy_o) <= 0.25 * energy_o + 0.05, (nm, energy, energy_o)
        if far.any():
            sim_i, _ = A.silence_and_speech(R, tmp_path, ideal[k][:n], got[k][:n], nm)
            assert sim_i >= A.ISOLATED_MIN[nm], (nm, sim_i)
