"""The echo-canceller oracle (oracle/oracle_aec.c) on its own, CPU only: its real transform against numpy's, and its
behaviour as a canceller — it must converge on a synthetic echo path, leave near-end speech alone and stay silent on
silence. (speexdsp is not in the reference tree: parity with the library stays unpinned, see the file's header; the
reference suite's own material and metric are in tests/test_oracle_aec_fixture.py.)"""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
from _oracle import ptr
from synth import cfg2_stream


def _lib():
    L = O.oracle()
    for name in ("orc_test_rfft", "orc_test_irfft"):
        fn = getattr(L, name)
        fn.restype, fn.argtypes = None, [C.c_int, C.c_void_p, C.c_void_p]
    return L


@pytest.mark.parametrize("n", [128, 256, 512])
def test_real_transform_matches_numpy(n):
    """packed spx_fft format [r0, r1, i1, ..., r(N/2-1), i(N/2-1), r(N/2)], forward scaled by 1/N, inverse unscaled"""
    L = _lib()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32) * 1000
    out = np.zeros(n, np.float32)
    L.orc_test_rfft(n, ptr(x), ptr(out))
    ref = np.fft.rfft(x.astype(np.float64)) / n
    exp = np.zeros(n)
    exp[0], exp[n - 1] = ref[0].real, ref[n // 2].real
    exp[1:n - 1:2], exp[2:n - 1:2] = ref[1:n // 2].real, ref[1:n // 2].imag
    assert np.abs(out - exp).max() <= 2e-6 * np.abs(exp).max() + 1e-4
    back = np.zeros(n, np.float32)
    L.orc_test_irfft(n, ptr(out), ptr(back))
    assert np.abs(back - x).max() <= 1e-3  # forward (1/N) then inverse (unscaled) is the identity to float rounding


def _run(rate, mic, ref, tail=250):
    L = O.oracle()
    a = L.orc_aec_new(rate, tail, 64)
    F = L.orc_aec_frame_size(a)
    n = len(mic) // F * F
    out = np.zeros(n, np.int16)
    for k in range(0, n, F):
        m, r, o = np.ascontiguousarray(mic[k:k + F]), np.ascontiguousarray(ref[k:k + F]), np.zeros(F, np.int16)
        L.orc_aec_process_frame(a, ptr(m), ptr(r), ptr(o))
        out[k:k + F] = o
    L.orc_aec_free(a)
    return out


def _db(x):
    return 10 * np.log10(max(np.mean(np.asarray(x, float) ** 2), 1e-9))


@pytest.mark.parametrize("rate", [8000, 16000])
def test_oracle_converges_on_a_synthetic_echo_path(rate):
    """echo only (no near end): after 3 s the residual sits well below the echo (ERLE), and it keeps improving"""
    n = 6 * rate
    x, _, echo, _ = cfg2_stream(3, n, rate)
    mic = np.clip(np.round(echo), -32768, 32767).astype(np.int16)
    out = _run(rate, mic, x)
    active = np.abs(echo) > 50
    seg1, seg2 = slice(1 * rate, 2 * rate), slice(4 * rate, 6 * rate)
    erle_early = _db(mic[seg1][active[seg1]]) - _db(out[seg1][active[seg1]])
    erle_late = _db(mic[seg2][active[seg2]]) - _db(out[seg2][active[seg2]])
    assert erle_late > 30.0, erle_late          # measured: 36.5 dB at 8 kHz, 35.6 dB at 16 kHz
    assert erle_early > 12.0, erle_early        # 16.9 / 17.9 dB in the second second
    assert erle_late > erle_early + 10.0, (erle_early, erle_late)


def test_oracle_leaves_near_end_speech_and_silence_alone():
    """silent far end: the output is the near-end signal through the DC notch and the denoiser — close in level, never
    louder; all-zero input gives all-zero output"""
    rate, n = 16000, 3 * 16000
    _, _, _, near = cfg2_stream(5, n, rate)
    mic = np.clip(np.round(near), -32768, 32767).astype(np.int16)
    out = _run(rate, mic, np.zeros(n, np.int16))
    talk = np.abs(near[:len(out)]) > 100
    assert talk.sum() > rate // 4
    d = _db(out[talk]) - _db(mic[:len(out)][talk])
    assert -1.5 < d < 0.2, d  # measured -0.28 dB
    z = _run(rate, np.zeros(n, np.int16), np.zeros(n, np.int16))
    assert not z.any()
