"""Thin numpy-facing wrappers over the C ABI, named after the reference filters they batch.

Used by tests/ and bench.py only. Each class owns one bank object of libmsb200dsp.so; method names follow the
reference's MS_* method ids (e.g. MS_AUDIO_MIXER_SET_INPUT_GAIN -> set_input_gain). Arrays are numpy int16/uint8,
C-contiguous; nothing here computes anything on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _req(a, dtype) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Context:
    """msb200_ctx: one per (process, GPU)."""

    def __init__(self, device: int = 0, cuda_stream: int | None = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if cuda_stream is None:
            check(self.lib.msb200_ctx_create(device, C.byref(h)))
        else:  # launch on a stream owned by the caller (e.g. torch.cuda.current_stream().cuda_stream)
            check(self.lib.msb200_ctx_create_on_stream(device, C.c_void_p(cuda_stream), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.msb200_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        check(self.lib.msb200_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.msb200_ctx_launch_count(self.h))

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self.lib.msb200_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p: int):
        check(self.lib.msb200_dev_free(self.h, C.c_void_p(p)))

    def pinned(self, shape, dtype) -> np.ndarray:
        """numpy view over pinned host memory (kept alive by the returned array's .base chain)."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        check(self.lib.msb200_host_alloc_pinned(self.h, n, C.byref(p)))
        buf = (C.c_uint8 * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt).reshape(shape)
        return arr

    def h2d(self, dev: int, a: np.ndarray):
        check(self.lib.msb200_memcpy_h2d(self.h, C.c_void_p(dev), _ptr(a), a.nbytes))

    def d2h(self, a: np.ndarray, dev: int):
        check(self.lib.msb200_memcpy_d2h(self.h, _ptr(a), C.c_void_p(dev), a.nbytes))

    def timer_start(self):
        check(self.lib.msb200_timer_start(self.h))

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        check(self.lib.msb200_timer_stop_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        check(self.lib.msb200_flush_l2(self.h))


class AudioMixer:
    """n_rooms x MSAudioMixer (audiomixer.c:288-346)."""

    def __init__(self, ctx: Context, n_rooms: int, n_pins: int, nwords: int, conference_mode: bool):
        self.ctx, self.lib = ctx, ctx.lib
        self.shape = (n_rooms, n_pins, nwords)
        self.conf = bool(conference_mode)
        h = C.c_void_p()
        check(self.lib.msb200_mixer_create(ctx.h, n_rooms, n_pins, nwords, int(self.conf), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.msb200_mixer_destroy(self.h)
            self.h = None

    def set_input_gain(self, room: int, pin: int, gain: float):
        check(self.lib.msb200_mixer_set_input_gain(self.h, room, pin, gain))

    def set_active(self, room: int, pin: int, active: bool):
        check(self.lib.msb200_mixer_set_active(self.h, room, pin, int(active)))

    def process(self, pcm: np.ndarray, present: np.ndarray | None = None) -> np.ndarray:
        r, p, n = self.shape
        pcm = _req(pcm, np.int16).reshape(r, p, n)
        present = np.ones((r, p), np.uint8) if present is None else _req(present, np.uint8).reshape(r, p)
        out = np.empty((r, p, n) if self.conf else (r, n), np.int16)
        check(self.lib.msb200_mixer_process(self.h, _ptr(pcm), _ptr(present), _ptr(out)))
        return out


class Volume:
    """n x MSVolume light path (msvolume.c:503-513)."""

    def __init__(self, ctx: Context, n_streams: int, sample_rate: int, max_block: int = 960):
        self.ctx, self.lib, self.n = ctx, ctx.lib, n_streams
        h = C.c_void_p()
        check(self.lib.msb200_volume_create(ctx.h, n_streams, sample_rate, max_block, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.msb200_volume_destroy(self.h)
            self.h = None

    def set_kernel(self, choice: int):
        """0: by bank size (default), 1: one warp per stream, 2: one lane per stream in the sequential part"""
        check(self.lib.msb200_volume_set_kernel(self.h, choice))

    def set_gain(self, stream: int, gain: float):
        check(self.lib.msb200_volume_set_gain(self.h, stream, gain))

    def set_db_gain(self, stream: int, db: float):
        check(self.lib.msb200_volume_set_db_gain(self.h, stream, db))

    def enable_noise_gate(self, stream: int, enabled: bool):
        check(self.lib.msb200_volume_enable_noise_gate(self.h, stream, int(enabled)))

    def set_noise_gate_threshold(self, stream: int, thr: float):
        check(self.lib.msb200_volume_set_noise_gate_threshold(self.h, stream, thr))

    def set_noise_gate_floorgain(self, stream: int, g: float):
        check(self.lib.msb200_volume_set_noise_gate_floorgain(self.h, stream, g))

    def remove_dc(self, stream: int, enabled: bool):
        check(self.lib.msb200_volume_remove_dc(self.h, stream, int(enabled)))

    def enable_agc(self, stream: int, enabled: bool):
        check(self.lib.msb200_volume_enable_agc(self.h, stream, int(enabled)))

    def set_peer(self, stream: int, peer_bank: "Volume", peer_stream: int):
        check(self.lib.msb200_volume_set_peer(self.h, stream, peer_bank.h, peer_stream))

    def set_ea(self, stream: int, threshold=None, speed=None, force=None, sustain=None, transmit_threshold=None):
        if threshold is not None:
            check(self.lib.msb200_volume_set_ea_threshold(self.h, stream, threshold))
        if speed is not None:
            check(self.lib.msb200_volume_set_ea_speed(self.h, stream, speed))
        if force is not None:
            check(self.lib.msb200_volume_set_ea_force(self.h, stream, force))
        if sustain is not None:
            check(self.lib.msb200_volume_set_ea_sustain(self.h, stream, sustain))
        if transmit_threshold is not None:
            check(self.lib.msb200_volume_set_ea_transmit_threshold(self.h, stream, transmit_threshold))

    def state(self, stream: int) -> _lib.VolumeState:
        st = _lib.VolumeState()
        check(self.lib.msb200_volume_get_state(self.h, stream, C.byref(st)))
        return st

    def process(self, pcm: np.ndarray) -> np.ndarray:
        io = _req(pcm, np.int16).reshape(self.n, -1).copy()
        check(self.lib.msb200_volume_process(self.h, _ptr(io), io.shape[1]))
        return io


class ChannelAdapter:
    """MSChannelAdapter (chanadapt.c:68-131), stateless."""

    def __init__(self, ctx: Context):
        self.ctx, self.lib = ctx, ctx.lib

    def process(self, mode: int, a: np.ndarray | None, b: np.ndarray | None = None) -> np.ndarray:
        ref = a if a is not None else b
        ref = _req(ref, np.int16)
        if mode == _lib.CHAN_STEREO_TO_MONO:
            n, frames = ref.shape[0], ref.shape[1] // 2
            out = np.empty((n, frames), np.int16)
        else:
            n, frames = ref.shape
            out = np.empty((n, frames * 2), np.int16)
        a = None if a is None else _req(a, np.int16)
        b = None if b is None else _req(b, np.int16)
        check(self.lib.msb200_chanadapt_process(self.ctx.h, mode, n, frames, _ptr(a), _ptr(b), _ptr(out)))
        return out


class Equalizer:
    """n x MSEqualizer (equalizer.c:279-288 -> dsptools.c:253-268)."""

    def __init__(self, ctx: Context, n_streams: int, sample_rate: int, max_block: int = 960):
        self.ctx, self.lib, self.n = ctx, ctx.lib, n_streams
        h = C.c_void_p()
        check(self.lib.msb200_equalizer_create(ctx.h, n_streams, sample_rate, max_block, C.byref(h)))
        self.h = h
        self.nfft = self.lib.msb200_equalizer_nfft(h)

    def close(self):
        if self.h:
            self.lib.msb200_equalizer_destroy(self.h)
            self.h = None

    def set_gain(self, stream: int, frequency: float, gain: float, width: float):
        check(self.lib.msb200_equalizer_set_gain(self.h, stream, frequency, gain, width))

    def get_gain(self, stream: int, frequency: float) -> float:
        g = C.c_float()
        check(self.lib.msb200_equalizer_get_gain(self.h, stream, frequency, C.byref(g)))
        return float(g.value)

    def set_active(self, stream: int, active: bool):
        check(self.lib.msb200_equalizer_set_active(self.h, stream, int(active)))

    def set_taps(self, stream: int, taps: np.ndarray):
        taps = _req(taps, np.float32)
        assert taps.size == self.nfft
        check(self.lib.msb200_equalizer_set_taps(self.h, stream, _ptr(taps)))

    def get_taps(self, stream: int) -> np.ndarray:
        t = np.empty(self.nfft, np.float32)
        check(self.lib.msb200_equalizer_get_taps(self.h, stream, _ptr(t)))
        return t

    def process(self, pcm: np.ndarray) -> np.ndarray:
        io = _req(pcm, np.int16).reshape(self.n, -1).copy()
        check(self.lib.msb200_equalizer_process(self.h, _ptr(io), io.shape[1]))
        return io


class Resample:
    """n x MSResample (msresample.c:122-179; speexdsp quality 3)."""

    def __init__(self, ctx: Context, n_streams: int, in_rate: int, out_rate: int, nchannels: int = 1,
                 max_in_frames: int = 960):
        self.ctx, self.lib, self.n, self.nch = ctx, ctx.lib, n_streams, nchannels
        h = C.c_void_p()
        check(self.lib.msb200_resample_create(ctx.h, n_streams, in_rate, out_rate, nchannels, max_in_frames, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.msb200_resample_destroy(self.h)
            self.h = None

    def max_out(self, in_frames: int) -> int:
        return self.lib.msb200_resample_max_out(self.h, in_frames)

    def process(self, pcm: np.ndarray) -> np.ndarray:
        """pcm [n][in_frames*nch] -> [n][out_frames*nch]"""
        x = _req(pcm, np.int16).reshape(self.n, -1)
        in_frames = x.shape[1] // self.nch
        cap = self.max_out(in_frames)
        out = np.zeros((self.n, cap * self.nch), np.int16)
        got = C.c_int()
        check(self.lib.msb200_resample_process(self.h, _ptr(x), in_frames, _ptr(out), cap, C.byref(got)))
        return out[:, : got.value * self.nch].copy()


class SpeexEC:
    """n x MSSpeexEC arithmetic (speexec.c:297-298: speex_echo_cancellation + speex_preprocess_run per frame)."""

    def __init__(self, ctx: Context, n_streams: int, sample_rate: int, tail_length_ms: int = 250,
                 framesize_at_8000: int = 64):
        self.ctx, self.lib, self.n = ctx, ctx.lib, n_streams
        h = C.c_void_p()
        check(self.lib.msb200_aec_create(ctx.h, n_streams, sample_rate, tail_length_ms, framesize_at_8000, C.byref(h)))
        self.h = h
        info = _lib.AecInfo()
        check(self.lib.msb200_aec_get_info(h, C.byref(info)))
        self.info = info
        self.frame_size = info.frame_size

    def close(self):
        if self.h:
            self.lib.msb200_aec_destroy(self.h)
            self.h = None

    def process(self, mic: np.ndarray, ref: np.ndarray) -> np.ndarray:
        mic = _req(mic, np.int16).reshape(self.n, -1)
        ref = _req(ref, np.int16).reshape(self.n, -1)
        assert mic.shape == ref.shape and mic.shape[1] % self.frame_size == 0
        out = np.empty_like(mic)
        check(self.lib.msb200_aec_process(self.h, _ptr(mic), _ptr(ref), _ptr(out), mic.shape[1] // self.frame_size))
        return out

    def set_path(self, path: int):
        check(self.lib.msb200_aec_set_path(self.h, path))

    def probe(self, stream: int, what: str, max_floats: int) -> np.ndarray:
        buf = np.zeros(max_floats, np.float32)
        n = self.lib.msb200_aec_probe(self.h, stream, what.encode(), _ptr(buf), max_floats)
        if n < 0:
            check(n)
        return buf[:n]

    def get_state_blob(self, stream: int) -> bytes:
        n = self.lib.msb200_aec_state_blob_size(self.h)
        buf = (C.c_uint8 * n)()
        check(self.lib.msb200_aec_get_state_blob(self.h, stream, buf, n))
        return bytes(buf)

    def set_state_blob(self, stream: int, blob: bytes):
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(self.lib.msb200_aec_set_state_blob(self.h, stream, buf, len(blob)))


class AudioChain:
    """BASELINE cfg2 graph resident on the device: 2x MSResample -> MSSpeexEC -> MSVolume [-> MSAudioMixer]."""

    def __init__(self, ctx: Context, n_streams: int, in_rate: int = 16000, rate: int = 48000, tail_length_ms: int = 250,
                 volume_gain: float = 0.8, mixer_pins: int = 0, use_cuda_graph: bool = False):
        self.ctx, self.lib, self.n = ctx, ctx.lib, n_streams
        p = _lib.ChainParams(n_streams, in_rate, rate, tail_length_ms, 64, volume_gain, mixer_pins, int(use_cuda_graph))
        h = C.c_void_p()
        check(self.lib.msb200_chain_create(ctx.h, C.byref(p), C.byref(h)))
        self.h = h
        self.in_frames = in_rate // 100
        self.max_out = self.lib.msb200_chain_max_out_samples(h)

    def close(self):
        if self.h:
            self.lib.msb200_chain_destroy(self.h)
            self.h = None

    def tick(self, ref_in: np.ndarray, mic_in: np.ndarray, out: np.ndarray | None = None):
        ref_in = _req(ref_in, np.int16)
        mic_in = _req(mic_in, np.int16)
        if out is None:
            out = np.zeros((self.n, self.max_out), np.int16)
        got = C.c_int()
        check(self.lib.msb200_chain_tick(self.h, _ptr(ref_in), _ptr(mic_in), _ptr(out), C.byref(got)))
        return out, got.value

    def submit(self, ref_in: np.ndarray, mic_in: np.ndarray, out: np.ndarray) -> int:
        """pipelined tick (pinned buffers; valid in `out` after the matching wait()): returns the sample count"""
        got = C.c_int()
        check(self.lib.msb200_chain_submit(self.h, _ptr(ref_in), _ptr(mic_in), _ptr(out), C.byref(got)))
        return got.value

    def wait(self):
        check(self.lib.msb200_chain_wait(self.h))

    def enable_kernel_timing(self, enabled: bool = True):
        check(self.lib.msb200_chain_enable_kernel_timing(self.h, int(enabled)))

    def kernel_timing(self):
        ms, nl, nf = C.c_float(), C.c_int(), C.c_int()
        check(self.lib.msb200_chain_get_kernel_timing(self.h, C.byref(ms), C.byref(nl), C.byref(nf)))
        return float(ms.value), nl.value, nf.value

    def set_overlap(self, enabled: bool):
        check(self.lib.msb200_chain_set_overlap(self.h, int(enabled)))

    def join(self):
        check(self.lib.msb200_chain_join(self.h))

    def tick_dev(self, d_ref: int, d_mic: int, d_out: int) -> int:
        got = C.c_int()
        check(self.lib.msb200_chain_tick_dev(self.h, C.c_void_p(d_ref), C.c_void_p(d_mic), C.c_void_p(d_out), C.byref(got)))
        return got.value


class Scaler:
    """MSScaler (msvideo.h:473-479) batched: pixel-format conversion + bilinear scale (swscale SWS_BILINEAR design)."""

    def __init__(self, ctx: Context, src_w: int, src_h: int, src_fmt: int, dst_w: int, dst_h: int, dst_fmt: int):
        self.ctx, self.lib = ctx, ctx.lib
        h = C.c_void_p()
        check(self.lib.msb200_scaler_create(ctx.h, src_w, src_h, src_fmt, dst_w, dst_h, dst_fmt, C.byref(h)))
        self.h = h
        self.src_bytes = self.lib.msb200_scaler_src_frame_bytes(h)
        self.dst_bytes = self.lib.msb200_scaler_dst_frame_bytes(h)

    def close(self):
        if self.h:
            self.lib.msb200_scaler_destroy(self.h)
            self.h = None

    def process(self, frames: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """host frames in, host frames out (pass pinned `frames` / `out` from Context.pinned for full PCIe speed)"""
        frames = _req(frames, np.uint8).reshape(-1, self.src_bytes)
        if out is None:
            out = np.empty((frames.shape[0], self.dst_bytes), np.uint8)
        check(self.lib.msb200_scaler_process(self.h, frames.shape[0], _ptr(frames), _ptr(out)))
        return out

    def process_dev(self, n_frames: int, d_src: int, d_dst: int):
        check(self.lib.msb200_scaler_process_dev(self.h, n_frames, C.c_void_p(d_src), C.c_void_p(d_dst)))

    def set_path(self, path: int):
        """tests/profiling: 0 best available, 1 persistent tile kernel, 2 generic tile kernel, 3 strip kernel, 4 streaming kernel (all bit-exact)"""
        check(self.lib.msb200_scaler_set_path(self.h, path))

    @property
    def path(self) -> int:
        """4 streaming kernel, 3 strip kernel, 2 persistent tile kernel, 1 generic tile kernel, 0 plane / packed kernels"""
        return self.lib.msb200_scaler_get_path(self.h)


    @property
    def schedule(self) -> tuple[int, int, int]:
        """(static row schedule index or -1, strips on the schedule, strips per frame column) of the strip kernel"""
        a, b = C.c_int(), C.c_int()
        k = self.lib.msb200_scaler_get_schedule(self.h, C.byref(a), C.byref(b))
        return k, a.value, b.value


def nv12_to_i420(ctx: Context, frames: np.ndarray, w: int, h: int, rotation: int = 0, y_stride: int | None = None,
                 cbcr_stride: int | None = None, u_first: bool = True, down_scale: bool = False,
                 cbcr_offset: int | None = None) -> np.ndarray:
    """Batched copy_ycbcrbiplanar_to_true_yuv_with_rotation_and_down_scale_by_2 (msvideo.c:787-919)."""
    frames = _req(frames, np.uint8)
    frames = frames.reshape(frames.shape[0], -1)
    n, fb = frames.shape
    f = 2 if down_scale else 1
    sw, sh = (w * f, h * f) if rotation % 180 == 0 else (h * f, w * f)
    y_stride = sw if y_stride is None else y_stride
    cbcr_stride = sw if cbcr_stride is None else cbcr_stride
    cbcr_offset = y_stride * sh if cbcr_offset is None else cbcr_offset
    out = np.empty((n, w * h * 3 // 2), np.uint8)
    check(ctx.lib.msb200_nv12_to_i420(ctx.h, n, _ptr(frames), fb, cbcr_offset, rotation, w, h, y_stride, cbcr_stride,
                                      int(u_first), int(down_scale), _ptr(out)))
    return out


G711_ALAW, G711_ULAW = 0, 1


def g711_decode(ctx: Context, law: int, code: np.ndarray) -> np.ndarray:
    """MSAlawDec / MSUlawDec arithmetic (alaw.c:199-211; g711.c:152-172, 249-262) over a flat batch of code words."""
    code = _req(code, np.uint8)
    pcm = np.empty(code.shape, np.int16)
    check(ctx.lib.msb200_g711_decode(ctx.h, law, _ptr(code), _ptr(pcm), code.size))
    return pcm


def g711_encode(ctx: Context, law: int, pcm: np.ndarray) -> np.ndarray:
    """MSAlawEnc / MSUlawEnc arithmetic (alaw.c:84-87; g711.c:119-146, 208-238) over a flat batch of samples."""
    pcm = _req(pcm, np.int16)
    code = np.empty(pcm.shape, np.uint8)
    check(ctx.lib.msb200_g711_encode(ctx.h, law, _ptr(pcm), _ptr(code), pcm.size))
    return code


def g711_decode_dev(ctx: Context, law: int, d_code: int, d_pcm: int, n: int):
    check(ctx.lib.msb200_g711_decode_dev(ctx.h, law, C.c_void_p(d_code), C.c_void_p(d_pcm), n))


def g711_encode_dev(ctx: Context, law: int, d_pcm: int, d_code: int, n: int):
    check(ctx.lib.msb200_g711_encode_dev(ctx.h, law, C.c_void_p(d_pcm), C.c_void_p(d_code), n))


class FlowControlState(C.Structure):  # msb200_flowcontrol_state
    _fields_ = [("strategy", C.c_int32), ("silent_threshold", C.c_float), ("target_samples", C.c_uint32),
                ("total_samples", C.c_uint32), ("current_pos", C.c_uint32), ("current_dropped", C.c_uint32)]


class FlowControl:
    """n x MSAudioFlowControl (flowcontrol.c:110-150): blocks pass, are dropped, or lose a few well chosen samples."""

    BASIC, SOFT = 0, 1

    def __init__(self, ctx: Context, n_streams: int, max_block: int = 960):
        self.ctx, self.lib, self.n = ctx, ctx.lib, n_streams
        h = C.c_void_p()
        check(self.lib.msb200_flowcontrol_create(ctx.h, n_streams, max_block, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.msb200_flowcontrol_destroy(self.h)
            self.h = None

    def set_config(self, stream: int, strategy: int, silent_threshold: float = 0.02):
        check(self.lib.msb200_flowcontrol_set_config(self.h, stream, strategy, silent_threshold))

    def set_target(self, stream: int, samples_to_drop: int, total_samples: int):
        check(self.lib.msb200_flowcontrol_set_target(self.h, stream, samples_to_drop, total_samples))

    def state(self, stream: int) -> FlowControlState:
        st = FlowControlState()
        check(self.lib.msb200_flowcontrol_get_state(self.h, stream, C.byref(st)))
        return st

    def process(self, pcm: np.ndarray):
        """pcm [n][nsamples] s16 -> (pcm processed in place copy, remaining sample count per stream)"""
        io = np.array(_req(pcm, np.int16).reshape(self.n, -1), copy=True)
        out_n = np.zeros(self.n, np.int32)
        check(self.lib.msb200_flowcontrol_process(self.h, _ptr(io), io.shape[1], _ptr(out_n)))
        return io, out_n


class GenericPLC:
    """n x MSGenericPLC (msgenericplc.c:61-157): per tick and stream either a received block (delayed by 5 ms, cross-faded
    out of a concealed stretch) or a concealed block stretched from the last 50 ms; the concealer clock is the caller's."""

    IDLE, PACKET, CONCEAL, AFTER_CNG = 0, 1, 2, 4

    def __init__(self, ctx: Context, n_streams: int, sample_rate: int, max_block: int | None = None):
        self.ctx, self.lib, self.n, self.rate = ctx, ctx.lib, n_streams, sample_rate
        h = C.c_void_p()
        check(self.lib.msb200_plc_create(ctx.h, n_streams, sample_rate, max_block or sample_rate // 50, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.lib.msb200_plc_destroy(self.h)
            self.h = None

    @property
    def history_samples(self) -> int:
        return self.lib.msb200_plc_history_samples(self.h)

    def reset_stream(self, stream: int):
        check(self.lib.msb200_plc_reset_stream(self.h, stream))

    def process(self, pcm: np.ndarray, mode) -> np.ndarray:
        """pcm [n][nsamples] s16 (rows of streams in CONCEAL / IDLE mode are ignored on input), mode [n] -> output rows"""
        io = np.array(_req(pcm, np.int16).reshape(self.n, -1), copy=True)
        md = np.ascontiguousarray(np.asarray(mode, np.uint8).reshape(self.n))
        check(self.lib.msb200_plc_process(self.h, _ptr(io), io.shape[1], _ptr(md)))
        return io
