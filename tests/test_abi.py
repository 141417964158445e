"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/msb200dsp.h
declares, and refuses to run without a GPU (no CPU fallback) — no compute calls here."""
import ctypes as C

import pytest

from mediastreamer2_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(_lib.LIB_PATH))
    declared = _lib.declared_symbols()
    assert len(declared) >= 70
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/msb200dsp.h but not exported: {missing}"


def test_binding_table_matches_header():
    assert set(_lib._SIGS) == set(_lib.declared_symbols())


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.msb200_version() >= 100
    assert isinstance(lib.msb200_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    """On a box without CUDA the context cannot be created and the error says why."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.msb200_ctx_create(0, C.byref(h))
    assert rc == _lib.ENODEV
    assert b"no CPU fallback" in lib.msb200_last_error()
    assert not h.value


def test_frame_size_helper_matches_reference_rule():
    # adjust_framesize(), speexec.c:171-180: largest power of two <= 64*rate/8000
    lib = _lib.load()
    assert lib.msb200_aec_frame_size_for_rate(8000, 64) == 64
    assert lib.msb200_aec_frame_size_for_rate(16000, 64) == 128
    assert lib.msb200_aec_frame_size_for_rate(48000, 64) == 256
    assert lib.msb200_aec_frame_size_for_rate(44100, 64) == 256
    assert lib.msb200_aec_frame_size_for_rate(32000, 64) == 256
