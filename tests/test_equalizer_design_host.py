"""The equalizer's tap design is host code in the product library (msb200_equalizer_design: ms_ifft of the packed gain table
through a restatement of the float kiss_fft, time shift, Hamming window). It runs without a device, so it is checked here:
bit for bit against the oracle's taps, which tests/test_oracle_vs_reference.py shows to reproduce the reference filter's
output exactly. With the FIR kernel bit-exact for equal taps (tests/test_gpu_audio.py), MSEqualizer is bit-exact end to end."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
from _oracle import ptr
from mediastreamer2_b200 import _lib


class _OrcEq(C.Structure):  # oracle/msb200_oracle.h orc_equalizer
    _fields_ = [("rate", C.c_int), ("nfft", C.c_int), ("needs_update", C.c_int), ("active", C.c_int),
                ("fft_cpx", C.POINTER(C.c_float)), ("fir", C.POINTER(C.c_float)), ("mem", C.POINTER(C.c_float))]


@pytest.mark.parametrize("seed", range(24))
def test_host_tap_design_equals_oracle_taps_bit_for_bit(seed):
    L, lib = O.oracle(), _lib.load()
    rng = np.random.default_rng(seed)
    rate = int(rng.choice([8000, 16000, 32000, 48000]))
    e = L.orc_equalizer_new(rate)
    for _ in range(int(rng.integers(0, 6))):
        L.orc_equalizer_set_gain(e, float(rng.uniform(50, rate / 2 - 50)), float(rng.choice([0.1, 0.4, 1.0, 2.0, 4.0])),
                                 float(rng.choice([50, 200, 400, 1000])))
    st = C.cast(e, C.POINTER(_OrcEq)).contents
    nfft = st.nfft
    assert nfft == (128 if rate < 16000 else (256 if rate < 32000 else 512))
    table = np.ctypeslib.as_array(st.fft_cpx, shape=(nfft,)).copy()
    want = np.ctypeslib.as_array(L.orc_equalizer_taps(e), shape=(nfft,)).copy()
    got = np.zeros(nfft, np.float32)
    assert lib.msb200_equalizer_design(nfft, ptr(table), ptr(got)) == 0
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(want).max() > 0
    L.orc_equalizer_free(e)


def test_host_tap_design_rejects_other_sizes():
    lib = _lib.load()
    t = np.zeros(1024, np.float32)
    assert lib.msb200_equalizer_design(100, ptr(t), ptr(t)) != 0
    assert lib.msb200_equalizer_design(1024, ptr(t), ptr(t)) != 0
