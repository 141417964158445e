"""Test-side loaders for the two CHECKERS (never imported by the product):

* ``oracle()``  -> oracle/libmsb200oracle.so, our CPU restatements (oracle/*.c)
* ``ref()``     -> oracle/_ref/libms2ref.so, the UNMODIFIED reference runtime + in-tree filters + oracle/ref_harness.c
                  (built here from /root/reference; travels prebuilt to the GPU box; may be absent -> tests skip)
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "libmsb200oracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libms2ref.so"
PLUGIN_DIR = ROOT / "plugin" / "lib"

_P, _I, _F = C.c_void_p, C.c_int, C.c_float


class OrcVolumeState(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("energy", "level_pk", "instant_energy", "gain", "static_gain", "target_gain",
                                         "ng_gain", "ng_threshold", "ng_floorgain")] + \
               [(n, C.c_int32) for n in ("dc_offset", "ng_noise_dur", "noise_gate_enabled", "remove_dc", "sample_rate",
                                         "fast_upramp")] + \
               [(n, C.c_float) for n in ("lt_speaker_en", "ea_thres", "ea_transmit_thres", "force", "vol_upramp")] + \
               [(n, C.c_int32) for n in ("sustain_time", "sustain_dur", "agc_enabled", "peer")]


_oracle = None
_ref = None


def _build_oracle():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "oracle"], check=True)


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is not None:
        return _oracle
    srcs = list((ROOT / "oracle").glob("oracle_*.c")) + [ROOT / "oracle" / "msb200_oracle.h"]
    if not ORACLE_SO.exists() or any(s.stat().st_mtime > ORACLE_SO.stat().st_mtime for s in srcs):
        _build_oracle()
    L = C.CDLL(str(ORACLE_SO))
    sig = {
        "orc_mixer_process": (None, [_I, _I, _I, _I, _P, _P, _P, _P, _P]),
        "orc_mixer_partial": (None, [_I, _I, _I, _P, _P, _P, _P, _P]),
        "orc_volume_init": (None, [C.POINTER(OrcVolumeState), _I]),
        "orc_volume_process": (None, [C.POINTER(OrcVolumeState), _P, _I]),
        "orc_volume_process_chunk": (None, [C.POINTER(OrcVolumeState), C.POINTER(C.c_float), _P, _I]),
        "orc_chanadapt": (None, [_I, _I, _I, _P, _P, _P]),
        "orc_equalizer_new": (_P, [_I]),
        "orc_equalizer_free": (None, [_P]),
        "orc_equalizer_set_gain": (None, [_P, _F, _F, _F]),
        "orc_equalizer_get_gain": (_F, [_P, _F]),
        "orc_equalizer_taps": (C.POINTER(C.c_float), [_P]),
        "orc_equalizer_process": (None, [_P, _P, _I]),
        "orc_fir_s16": (None, [_P, _I, _P, _P, _I]),
        "orc_resampler_new": (_P, [_I, _I, _I, _I]),
        "orc_resampler_free": (None, [_P]),
        "orc_msresample_block": (_I, [_P, _P, _I, _P]),
        "orc_resampler_filt_len": (_I, [_P]),
        "orc_resampler_use_direct": (_I, [_P]),
    }
    optional = {
        "orc_aec_frame_size_for_rate": (_I, [_I, _I]),
        "orc_aec_new": (_P, [_I, _I, _I]),
        "orc_aec_free": (None, [_P]),
        "orc_aec_frame_size": (_I, [_P]),
        "orc_aec_M": (_I, [_P]),
        "orc_aec_process_frame": (None, [_P, _P, _P, _P]),
        "orc_aec_cancel_frame": (None, [_P, _P, _P, _P]),
        "orc_aec_preprocess_frame": (None, [_P, _P]),
        "orc_aec_probe": (_I, [_P, C.c_char_p, _P, _I]),
        "orc_nv12_to_i420": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
        "orc_scaler_new": (_P, [_I, _I, _I, _I, _I, _I]),
        "orc_scaler_free": (None, [_P]),
        "orc_scaler_src_bytes": (C.c_size_t, [_P]),
        "orc_scaler_dst_bytes": (C.c_size_t, [_P]),
        "orc_scaler_process": (_I, [_P, _P, _P]),
        "orc_scaler_get_filter": (_I, [_P, _I, _P, _P, _I]),
        "orc_scaler_set_x86_vertical": (None, [_P, _I]),
        "orc_flowctl_init": (None, [_P]),
        "orc_flowctl_set_target": (None, [_P, C.c_uint32, C.c_uint32]),
        "orc_flowctl_process": (_I, [_P, _P, _I]),
        "orc_g711_encode": (None, [_I, _P, _P, C.c_size_t]),
        "orc_g711_decode": (None, [_I, _P, _P, C.c_size_t]),
        "orc_plc_rate_supported": (_I, [_I]),
        "orc_plc_create": (_P, [_I]),
        "orc_plc_destroy": (None, [_P]),
        "orc_plc_history_len": (_I, [_P]),
        "orc_plc_packet": (None, [_P, _P, _I, _I]),
        "orc_plc_conceal": (None, [_P, _P, _I]),
        "orc_plc_filter_set_cn": (None, [_P]),
        "orc_plc_filter_packet": (None, [_P, C.c_uint64, _P, _I, _I]),
        "orc_plc_filter_tick": (_I, [_P, C.c_uint64, _I, _I, _P, _P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    for name, (res, args) in optional.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
    _oracle = L
    return L


def have_ref() -> bool:
    return REF_SO.exists()


def ref() -> C.CDLL:
    global _ref
    if _ref is not None:
        return _ref
    if not REF_SO.exists():
        pytest.skip("oracle/_ref/libms2ref.so not built (needs /root/reference)")
    # RTLD_GLOBAL: plugins dlopen'ed by the reference's loader resolve ms_*/ortp symbols against this library
    L = C.CDLL(str(REF_SO), mode=C.RTLD_GLOBAL)
    sig = {
        "ref_factory_new": (_P, [C.c_char_p]),
        "ref_factory_destroy": (None, [_P]),
        "ref_filter_new": (_P, [_P, C.c_char_p]),
        "ref_filter_text": (C.c_char_p, [_P]),
        "ref_filter_destroy": (None, [_P]),
        "ref_filter_call": (_I, [_P, C.c_uint, _P]),
        "ref_link": (_I, [_P, _I, _P, _I]),
        "ref_unlink": (_I, [_P, _I, _P, _I]),
        "ref_source_push": (None, [_P, _I, _P, _I]),
        "ref_source_push_stream": (None, [_P, _I, _P, _I, _I]),
        "ref_source_push_video": (None, [_P, _I, _P, _I, _I, _I, C.c_uint]),
        "ref_sink_read_dims": (None, [_P, _P]),
        "ref_sink_set_discard": (None, [_P, _I]),
        "ref_set_scaler_callbacks": (None, [_P, _P, _P]),
        "ref_set_scaler_desc": (None, [_P]),
        "ref_sink_size": (C.c_long, [_P]),
        "ref_sink_nblocks": (_I, [_P]),
        "ref_sink_read": (None, [_P, _P, _P]),
        "ref_ticker_new": (_P, []),
        "ref_ticker_attach": (_I, [_P, _P]),
        "ref_ticker_detach": (_I, [_P, _P]),
        "ref_ticker_run": (None, [_P, _I]),
        "ref_ticker_release": (None, [_P, _I]),
        "ref_ticker_wait": (None, [_P]),
        "ref_ticker_time": (C.c_ulonglong, [_P]),
        "ref_ticker_destroy": (None, [_P]),
        "ref_method_id": (C.c_uint, [C.c_char_p]),
        "ref_nv12_to_i420": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
        "ref_fir_mem16": (None, [_P, _P, _P, _I, _I, _P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _ref = L
    return L


class MixerCtl(C.Structure):  # MSAudioMixerCtl, include/mediastreamer2/msaudiomixer.h:26-43
    class _U(C.Union):
        _fields_ = [("gain", C.c_float), ("active", C.c_int), ("enabled", C.c_int)]

    _fields_ = [("pin", C.c_int), ("param", _U)]


class EqualizerGain(C.Structure):  # MSEqualizerGain, include/mediastreamer2/msequalizer.h:26-31
    _fields_ = [("frequency", C.c_float), ("gain", C.c_float), ("width", C.c_float)]


class OrcFlowCtl(C.Structure):
    _fields_ = [("strategy", C.c_int32), ("silent_threshold", C.c_float), ("target_samples", C.c_uint32),
                ("total_samples", C.c_uint32), ("current_pos", C.c_uint32), ("current_dropped", C.c_uint32)]


class FlowControlConfig(C.Structure):  # MSAudioFlowControlConfig
    _fields_ = [("strategy", C.c_int), ("silent_threshold", C.c_float)]


class FlowControlDropEvent(C.Structure):  # MSAudioFlowControlDropEvent
    _fields_ = [("flow_control_interval_ms", C.c_uint32), ("drop_ms", C.c_uint32)]


class RefGraph:
    """A reference MSFactory + gated MSTicker; filters by name; scripted sources and recording sinks.

    ``plugins_dir`` makes the reference's own loader dlopen libmsb200dsp's plugin so that the same graph script
    runs against the drop-in filters.
    """

    def __init__(self, plugins_dir: str | None = None):
        self.L = ref()
        self.fac = self.L.ref_factory_new((plugins_dir or "").encode())
        self.ticker = None
        self.filters = []
        self.attached = []

    def new(self, name: str):
        f = self.L.ref_filter_new(self.fac, name.encode())
        assert f, f"filter {name} not found"
        self.filters.append(f)
        return f

    def text(self, f) -> str:
        return self.L.ref_filter_text(f).decode()

    def call(self, f, method: str, arg) -> int:
        mid = self.L.ref_method_id(method.encode())
        assert mid != 0, method
        return self.L.ref_filter_call(f, mid, C.byref(arg) if arg is not None else None)

    def call_ptr(self, f, method: str, ptr_value) -> int:
        """methods whose argument IS a pointer (e.g. MS_VOLUME_SET_PEER takes the peer MSFilter*)"""
        mid = self.L.ref_method_id(method.encode())
        assert mid != 0, method
        return self.L.ref_filter_call(f, mid, ptr_value)

    def call_int(self, f, method: str, value: int) -> int:
        return self.call(f, method, C.c_int(value))

    def call_float(self, f, method: str, value: float) -> int:
        return self.call(f, method, C.c_float(value))

    def link(self, f1, p1, f2, p2):
        assert self.L.ref_link(f1, p1, f2, p2) == 0

    def source(self, pcm: np.ndarray | None = None, block_bytes: int = 0, tick0: int = 0):
        s = self.new("HarnessSource")
        if pcm is not None:
            self.push_stream(s, pcm, block_bytes, tick0)
        return s

    def push_stream(self, s, data: np.ndarray, block_bytes: int, tick0: int = 0):
        data = np.ascontiguousarray(data)
        nblocks = data.nbytes // block_bytes
        self.L.ref_source_push_stream(s, tick0, data.ctypes.data_as(C.c_void_p), block_bytes, nblocks)

    def push(self, s, tick: int, data: np.ndarray):
        data = np.ascontiguousarray(data)
        self.L.ref_source_push(s, tick, data.ctypes.data_as(C.c_void_p), data.nbytes)

    def push_video(self, s, tick: int, frame: np.ndarray, w: int, h: int, ts: int):
        """w > 0: an I420 frame in a header-carrying video mblk (ms_yuv_buf_alloc); w == 0: a raw packed frame"""
        frame = np.ascontiguousarray(frame)
        self.L.ref_source_push_video(s, tick, frame.ctypes.data_as(C.c_void_p), frame.nbytes, w, h, ts)

    def read_dims(self, sink) -> np.ndarray:
        nb = self.L.ref_sink_nblocks(sink)
        d = np.zeros((nb, 2), np.int32)
        if nb:
            self.L.ref_sink_read_dims(sink, d.ctypes.data_as(C.c_void_p))
        return d

    def sink(self):
        return self.new("HarnessSink")

    def read(self, sink, dtype=np.int16):
        n = self.L.ref_sink_size(sink)
        nb = self.L.ref_sink_nblocks(sink)
        buf = np.zeros(n, np.uint8)
        tri = np.zeros((nb, 3), np.int32)
        self.L.ref_sink_read(sink, buf.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p))
        return buf.view(dtype), tri

    def run(self, attach_to, nticks: int):
        """attach_to: one source filter, or a list of sources of disconnected graphs (executed in this order)"""
        if self.ticker is None:
            self.ticker = self.L.ref_ticker_new()
        for src in (attach_to if isinstance(attach_to, (list, tuple)) else [attach_to]):
            if src not in self.attached:
                assert self.L.ref_ticker_attach(self.ticker, src) == 0
                self.attached.append(src)
        self.L.ref_ticker_run(self.ticker, nticks)

    def close(self):
        for f in self.attached:
            self.L.ref_ticker_detach(self.ticker, f)
        self.attached = []
        if self.ticker:
            self.L.ref_ticker_destroy(self.ticker)
            self.ticker = None


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
