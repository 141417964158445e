"""ctypes binding of libmsb200dsp.so (the C ABI declared in include/msb200dsp.h).

This module is plumbing for tests and bench.py: the product is the shared library. It never imports anything from
``oracle/`` and raises loudly when the library is missing or cannot be loaded (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ.get("MSB200_LIB", ROOT / "mediastreamer2_b200" / "lib" / "libmsb200dsp.so"))
HEADER_PATH = ROOT / "include" / "msb200dsp.h"

OK, EINVAL, ENODEV, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5

CHAN_MONO_TO_STEREO, CHAN_STEREO_TO_MONO, CHAN_2MONO_TO_STEREO = 0, 1, 2
PIX_YUV420P, PIX_YUYV, PIX_RGB24, PIX_RGB24_REV, PIX_UYVY, PIX_YUY2, PIX_RGBA32, PIX_RGBA32_REV = 0, 1, 2, 3, 5, 6, 7, 11
PIX_NV12, PIX_NV21 = 100, 101
PIX_RGB565 = 8


class Msb200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"msb200 error {code}: {msg}")
        self.code = code


class VolumeState(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("energy", "level_pk", "instant_energy", "gain", "static_gain", "target_gain",
                                         "ng_gain", "ng_threshold", "ng_floorgain")] + \
               [(n, C.c_int32) for n in ("dc_offset", "ng_noise_dur", "noise_gate_enabled", "remove_dc", "sample_rate",
                                         "fast_upramp")] + \
               [(n, C.c_float) for n in ("lt_speaker_en", "ea_thres", "ea_transmit_thres", "force", "vol_upramp")] + \
               [(n, C.c_int32) for n in ("sustain_time", "sustain_dur", "agc_enabled", "peer")]


class AecInfo(C.Structure):
    _fields_ = [("frame_size", C.c_int32), ("window_size", C.c_int32), ("M", C.c_int32), ("sample_rate", C.c_int32),
                ("filter_length", C.c_int32), ("state_bytes_per_stream", C.c_size_t)]


class Rect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32)]


class YuvLayout(C.Structure):
    _fields_ = [("plane_offset", C.c_size_t * 3), ("row_stride", C.c_int32 * 3), ("pix_stride", C.c_int32 * 3),
                ("frame_bytes", C.c_size_t)]


class RtpMeta(C.Structure):
    _fields_ = [("timestamp", C.c_uint32), ("ssrc", C.c_uint32), ("payload_len", C.c_int32), ("seq", C.c_uint16),
                ("marker", C.c_uint8), ("payload_type", C.c_uint8)]


class ChainParams(C.Structure):
    _fields_ = [("n_streams", C.c_int32), ("in_rate", C.c_int32), ("rate", C.c_int32), ("tail_length_ms", C.c_int32),
                ("framesize_at_8000", C.c_int32), ("volume_gain", C.c_float), ("mixer_pins", C.c_int32),
                ("use_cuda_graph", C.c_int32)]


def declared_symbols() -> list[str]:
    """Every function name declared MSB200_API in the public header."""
    text = HEADER_PATH.read_text()
    return sorted(set(re.findall(r"MSB200_API\s+[\w\s\*]+?\b(msb200_\w+)\s*\(", text)))


_P, _I, _F, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_PP = C.POINTER(C.c_void_p)
_PI = C.POINTER(C.c_int)

_SIGS = {
    "msb200_version": (C.c_int, []),
    "msb200_last_error": (C.c_char_p, []),
    "msb200_ctx_create": (_I, [_I, _PP]),
    "msb200_ctx_create_on_stream": (_I, [_I, _P, _PP]),
    "msb200_ctx_destroy": (None, [_P]),
    "msb200_ctx_sync": (_I, [_P]),
    "msb200_ctx_set_deferred_sync": (_I, [_P, _I]),
    "msb200_ctx_launch_count": (C.c_uint64, [_P]),
    "msb200_dev_alloc": (_I, [_P, _SZ, _PP]),
    "msb200_dev_free": (_I, [_P, _P]),
    "msb200_host_alloc_pinned": (_I, [_P, _SZ, _PP]),
    "msb200_host_free_pinned": (_I, [_P, _P]),
    "msb200_memcpy_h2d": (_I, [_P, _P, _P, _SZ]),
    "msb200_memcpy_d2h": (_I, [_P, _P, _P, _SZ]),
    "msb200_memset_dev": (_I, [_P, _P, _I, _SZ]),
    "msb200_timer_start": (_I, [_P]),
    "msb200_timer_stop_ms": (_I, [_P, C.POINTER(C.c_float)]),
    "msb200_flush_l2": (_I, [_P]),
    "msb200_mixer_create": (_I, [_P, _I, _I, _I, _I, _PP]),
    "msb200_mixer_destroy": (None, [_P]),
    "msb200_mixer_set_input_gain": (_I, [_P, _I, _I, _F]),
    "msb200_mixer_set_active": (_I, [_P, _I, _I, _I]),
    "msb200_mixer_process": (_I, [_P, _P, _P, _P]),
    "msb200_mixer_process_dev": (_I, [_P, _P, _P, _P]),
    "msb200_mixer_partial_dev": (_I, [_P, _P, _P, _P]),
    "msb200_mixer_finish_dev": (_I, [_P, _P, _P, _P, _P]),
    "msb200_ipc_export": (_I, [_P, _P, _P]),
    "msb200_ipc_import": (_I, [_P, _P, _PP]),
    "msb200_ipc_close": (_I, [_P, _P]),
    "msb200_comm_available": (_I, []),
    "msb200_comm_nccl_version": (_I, []),
    "msb200_comm_unique_id": (_I, [_P]),
    "msb200_comm_create": (_I, [_P, _P, _I, _I, _PP]),
    "msb200_comm_destroy": (None, [_P]),
    "msb200_comm_allreduce_sum_i32_dev": (_I, [_P, _P, _SZ]),
    "msb200_mixer_process_striped_dev": (_I, [_P, _P, _P, _P, _P, _P]),
    "msb200_mixer_xchg_create": (_I, [_P, _I, _I, _PP]),
    "msb200_mixer_xchg_export": (_I, [_P, _P]),
    "msb200_mixer_xchg_connect": (_I, [_P, _P]),
    "msb200_mixer_xchg_connect_local": (_I, [_P, _PP]),
    "msb200_mixer_xchg_process_dev": (_I, [_P, _P, _P, _P]),
    "msb200_mixer_xchg_status": (_I, [_P, C.POINTER(C.c_uint32)]),
    "msb200_mixer_xchg_wire_bytes_per_tick": (_SZ, [_P]),
    "msb200_mixer_xchg_destroy": (None, [_P]),
    "msb200_volume_create": (_I, [_P, _I, _I, _I, _PP]),
    "msb200_volume_destroy": (None, [_P]),
    "msb200_ctx_make_current": (_I, [_P]),
    "msb200_flowcontrol_create": (_I, [_P, _I, _I, _PP]),
    "msb200_flowcontrol_destroy": (None, [_P]),
    "msb200_flowcontrol_set_config": (_I, [_P, _I, _I, _F]),
    "msb200_flowcontrol_set_target": (_I, [_P, _I, C.c_uint32, C.c_uint32]),
    "msb200_flowcontrol_reset": (_I, [_P, _I]),
    "msb200_flowcontrol_get_state": (_I, [_P, _I, _P]),
    "msb200_flowcontrol_process": (_I, [_P, _P, _I, _P]),
    "msb200_flowcontrol_process_dev": (_I, [_P, _P, _I, _I, _P]),
    "msb200_plc_create": (_I, [_P, _I, _I, _I, _PP]),
    "msb200_plc_destroy": (None, [_P]),
    "msb200_plc_history_samples": (_I, [_P]),
    "msb200_plc_reset_stream": (_I, [_P, _I]),
    "msb200_plc_process": (_I, [_P, _P, _I, _P]),
    "msb200_plc_process_strided": (_I, [_P, _P, _I, _I, _P]),
    "msb200_plc_set_live": (_I, [_P, _I]),
    "msb200_plc_process_dev": (_I, [_P, _P, _I, _I, _P]),
    "msb200_g711_decode": (_I, [_P, _I, _P, _P, _SZ]),
    "msb200_g711_encode": (_I, [_P, _I, _P, _P, _SZ]),
    "msb200_g711_decode_dev": (_I, [_P, _I, _P, _P, _SZ]),
    "msb200_g711_encode_dev": (_I, [_P, _I, _P, _P, _SZ]),
    "msb200_volume_reset_stream": (_I, [_P, _I]),
    "msb200_volume_set_live": (_I, [_P, _I]),
    "msb200_volume_set_kernel": (_I, [_P, _I]),
    "msb200_mixer_set_live": (_I, [_P, _I]),
    "msb200_resample_set_live": (_I, [_P, _I]),
    "msb200_aec_set_live": (_I, [_P, _I]),
    "msb200_aec_set_path": (_I, [_P, _I]),
    "msb200_volume_set_gain": (_I, [_P, _I, _F]),
    "msb200_volume_set_db_gain": (_I, [_P, _I, _F]),
    "msb200_volume_enable_noise_gate": (_I, [_P, _I, _I]),
    "msb200_volume_set_noise_gate_threshold": (_I, [_P, _I, _F]),
    "msb200_volume_set_noise_gate_floorgain": (_I, [_P, _I, _F]),
    "msb200_volume_remove_dc": (_I, [_P, _I, _I]),
    "msb200_volume_enable_agc": (_I, [_P, _I, _I]),
    "msb200_volume_set_peer": (_I, [_P, _I, _P, _I]),
    "msb200_volume_set_ea_threshold": (_I, [_P, _I, _F]),
    "msb200_volume_set_ea_speed": (_I, [_P, _I, _F]),
    "msb200_volume_set_ea_force": (_I, [_P, _I, _F]),
    "msb200_volume_set_ea_sustain": (_I, [_P, _I, _I]),
    "msb200_volume_set_ea_transmit_threshold": (_I, [_P, _I, _F]),
    "msb200_volume_get_state": (_I, [_P, _I, C.POINTER(VolumeState)]),
    "msb200_volume_process": (_I, [_P, _P, _I]),
    "msb200_volume_process_blocks": (_I, [_P, _P, _I, _I, _I, _P]),
    "msb200_volume_process_dev": (_I, [_P, _P, _I, _I]),
    "msb200_chanadapt_process": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "msb200_chanadapt_process_dev": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "msb200_equalizer_create": (_I, [_P, _I, _I, _I, _PP]),
    "msb200_equalizer_destroy": (None, [_P]),
    "msb200_equalizer_nfft": (_I, [_P]),
    "msb200_equalizer_set_gain": (_I, [_P, _I, _F, _F, _F]),
    "msb200_equalizer_get_gain": (_I, [_P, _I, _F, C.POINTER(C.c_float)]),
    "msb200_equalizer_set_active": (_I, [_P, _I, _I]),
    "msb200_equalizer_design": (_I, [_I, _P, _P]),
    "msb200_equalizer_set_taps": (_I, [_P, _I, _P]),
    "msb200_equalizer_get_taps": (_I, [_P, _I, _P]),
    "msb200_equalizer_process": (_I, [_P, _P, _I]),
    "msb200_equalizer_process_dev": (_I, [_P, _P, _I, _I]),
    "msb200_resample_create": (_I, [_P, _I, _I, _I, _I, _I, _PP]),
    "msb200_resample_destroy": (None, [_P]),
    "msb200_resample_max_out": (_I, [_P, _I]),
    "msb200_resample_reset": (_I, [_P]),
    "msb200_resample_reset_stream": (_I, [_P, _I]),
    "msb200_resample_process": (_I, [_P, _P, _I, _P, _I, _PI]),
    "msb200_resample_process_dev": (_I, [_P, _P, _I, _I, _P, _I, _PI]),
    "msb200_aec_frame_size_for_rate": (_I, [_I, _I]),
    "msb200_aec_create": (_I, [_P, _I, _I, _I, _I, _PP]),
    "msb200_aec_destroy": (None, [_P]),
    "msb200_aec_get_info": (_I, [_P, C.POINTER(AecInfo)]),
    "msb200_aec_reset": (_I, [_P, _I]),
    "msb200_aec_process": (_I, [_P, _P, _P, _P, _I]),
    "msb200_aec_process_dev": (_I, [_P, _P, _P, _P, _I, _I]),
    "msb200_aec_process_strided": (_I, [_P, _P, _P, _P, _I, _I]),
    "msb200_aec_process_counts": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "msb200_aec_process_counts_dev": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "msb200_aec_state_blob_size": (_SZ, [_P]),
    "msb200_aec_get_state_blob": (_I, [_P, _I, _P, _SZ]),
    "msb200_aec_set_state_blob": (_I, [_P, _I, _P, _SZ]),
    "msb200_aec_probe": (_I, [_P, _I, C.c_char_p, _P, _I]),
    "msb200_rtp_parse": (_I, [_P, _SZ, C.POINTER(RtpMeta), C.POINTER(C.c_size_t)]),
    "msb200_rtp_rx_create": (_I, [_P, _I, _I, _I, _PP]),
    "msb200_rtp_rx_destroy": (None, [_P]),
    "msb200_rtp_rx_row_samples": (_I, [_P]),
    "msb200_rtp_rx_begin_tick": (_I, [_P]),
    "msb200_rtp_rx_push": (_I, [_P, _I, _P, _SZ, _I]),
    "msb200_rtp_rx_push_payload": (_I, [_P, _I, _P, _I, C.POINTER(RtpMeta)]),
    "msb200_rtp_rx_decode": (_I, [_P, _P, _P]),
    "msb200_rtp_rx_decode_dev": (_I, [_P, _PP, _PP]),
    "msb200_rtp_tx_create": (_I, [_P, _I, _I, _I, _PP]),
    "msb200_rtp_tx_destroy": (None, [_P]),
    "msb200_rtp_tx_set_stream": (_I, [_P, _I, C.c_uint32, _I, C.c_uint16, C.c_uint32]),
    "msb200_rtp_tx_packet_bytes": (_SZ, [_P]),
    "msb200_rtp_tx_encode": (_I, [_P, _P, _P, _P, _PP]),
    "msb200_rtp_tx_encode_dev": (_I, [_P, _P, _P, _P, _PP]),
    "msb200_chain_create": (_I, [_P, C.POINTER(ChainParams), _PP]),
    "msb200_chain_destroy": (None, [_P]),
    "msb200_chain_next_out_samples": (_I, [_P]),
    "msb200_chain_max_out_samples": (_I, [_P]),
    "msb200_chain_tick": (_I, [_P, _P, _P, _P, _PI]),
    "msb200_chain_submit": (_I, [_P, _P, _P, _P, _PI]),
    "msb200_chain_wait": (_I, [_P]),
    "msb200_chain_tick_dev": (_I, [_P, _P, _P, _P, _PI]),
    "msb200_chain_set_overlap": (_I, [_P, _I]),
    "msb200_chain_join": (_I, [_P]),
    "msb200_chain_launches_per_tick": (_I, [_P]),
    "msb200_chain_aec": (_P, [_P]),
    "msb200_chain_enable_kernel_timing": (_I, [_P, _I]),
    "msb200_chain_get_kernel_timing": (_I, [_P, C.POINTER(C.c_float), _PI, _PI]),
    "msb200_nv12_to_i420": (_I, [_P, _I, _P, _SZ, _SZ, _I, _I, _I, _I, _I, _I, _I, _P]),
    "msb200_nv12_to_i420_dev": (_I, [_P, _I, _P, _SZ, _SZ, _I, _I, _I, _I, _I, _I, _I, _P]),
    "msb200_yuv_copy_strided": (_I, [_P, _I, _P, C.POINTER(YuvLayout), Rect, _P, C.POINTER(YuvLayout), Rect]),
    "msb200_yuv_copy_strided_dev": (_I, [_P, _I, _P, C.POINTER(YuvLayout), Rect, _P, C.POINTER(YuvLayout), Rect]),
    "msb200_scaler_create": (_I, [_P, _I, _I, _I, _I, _I, _I, _PP]),
    "msb200_scaler_destroy": (None, [_P]),
    "msb200_scaler_src_frame_bytes": (_SZ, [_P]),
    "msb200_scaler_dst_frame_bytes": (_SZ, [_P]),
    "msb200_scaler_process": (_I, [_P, _I, _P, _P]),
    "msb200_scaler_process_dev": (_I, [_P, _I, _P, _P]),
    "msb200_scaler_process_frames": (_I, [_P, _I, _P, _P]),
    "msb200_scaler_set_canvas": (_I, [_P, _I, _I, _I, _P]),
    "msb200_scaler_set_x86_vertical": (_I, [_P, _I]),
    "msb200_scaler_canvas_bytes": (_SZ, [_P]),
    "msb200_scaler_set_path": (_I, [_P, _I]),
    "msb200_scaler_get_schedule": (_I, [_P, _PI, _PI]),
    "msb200_scaler_get_path": (_I, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the product library; fail loudly if it is absent (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(nvcc, sm_100a). mediastreamer2_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree: let it propagate
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        raise Msb200Error(code, load().msb200_last_error().decode("utf-8", "replace"))
