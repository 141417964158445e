"""The video drop-in boundary as one graph script: HarnessSource -> MSPixConv -> MSSizeConv -> HarnessSink in an UNMODIFIED
MSTicker, run with
  * the reference's own filters (src/videofilters/pixconv.c, sizeconv.c compiled unmodified into oracle/_ref) over a scaler
    that calls the oracle (an MSScalerDesc built from ctypes callbacks: the "expected" frames),
  * the reference's own filters over the plugin's GPU MSScalerDesc (ms_video_set_scaler_impl(msb200_ms_scaler_desc())),
  * the plugin's own MSPixConv / MSSizeConv descs (picked by name through the reference factory).
Frames, their 16-byte video headers (w, h) and their timestamps must agree."""
from __future__ import annotations

import ctypes as C

import numpy as np

import _oracle as O
from _oracle import RefGraph

# MSPixFmt (include/mediastreamer2/msvideo.h:267-280) -> oracle / MSB200_PIX_* constants
MS_YUV420P, MS_YUYV, MS_RGB24, MS_RGB24_REV, MS_UYVY, MS_YUY2, MS_RGBA32, MS_RGBA32_REV = 1, 2, 3, 4, 6, 7, 8, 11
MS_RGB565 = 9
MS_NV12, MS_NV21 = 12, 13  # include/msb200_ms2.h
_TO_ORC = {MS_RGB565: 8, MS_YUV420P: 0, MS_YUYV: 1, MS_RGB24: 2, MS_RGB24_REV: 3, MS_UYVY: 5, MS_YUY2: 6, MS_RGBA32: 7, MS_RGBA32_REV: 11,
           MS_NV12: 100, MS_NV21: 101}


class VideoSize(C.Structure):  # MSVideoSize
    _fields_ = [("width", C.c_int), ("height", C.c_int)]


_CREATE = C.CFUNCTYPE(C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_PROCESS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_int))
_FREE = C.CFUNCTYPE(None, C.c_void_p)


class OracleScalerDesc:
    """MSScalerDesc whose context_process runs oracle/oracle_video.c (planes packed / unpacked in numpy). Keep the object
    alive while the reference may call it."""

    def __init__(self):
        self.L = O.oracle()
        self.ctxs = {}
        self.next = 1
        self.calls = 0

        def create(sw, sh, sf, dw, dh, df, flags):
            o = self.L.orc_scaler_new(sw, sh, _TO_ORC.get(sf, -1), dw, dh, _TO_ORC.get(df, -1))
            if not o:
                return None
            k = self.next
            self.next += 1
            self.ctxs[k] = (o, sw, sh, sf, dw, dh, df)
            return k

        def rows(ptr, stride, row, n):
            out = np.empty((n, row), np.uint8)
            for y in range(n):
                out[y] = np.ctypeslib.as_array((C.c_uint8 * row).from_address(ptr + y * stride))
            return out.reshape(-1)

        def process(k, src, sstr, dst, dstr):
            o, sw, sh, sf, dw, dh, df = self.ctxs[k]
            self.calls += 1
            if sf == MS_YUV420P:
                cw, ch = (sw + 1) // 2, (sh + 1) // 2
                frame = np.concatenate([rows(src[0], sstr[0], sw, sh), rows(src[1], sstr[1], cw, ch), rows(src[2], sstr[2], cw, ch)])
            else:
                bpp = {MS_YUYV: 2, MS_YUY2: 2, MS_UYVY: 2, MS_RGB565: 2, MS_RGB24: 3, MS_RGB24_REV: 3, MS_RGBA32: 4, MS_RGBA32_REV: 4}[sf]
                frame = rows(src[0], sstr[0], sw * bpp, sh)
            assert frame.nbytes == self.L.orc_scaler_src_bytes(o)
            out = np.zeros(self.L.orc_scaler_dst_bytes(o), np.uint8)
            self.L.orc_scaler_process(o, O.ptr(np.ascontiguousarray(frame)), O.ptr(out))
            assert df == MS_YUV420P
            cw, ch = (dw + 1) // 2, (dh + 1) // 2
            off = 0
            for p, (w, h) in enumerate(((dw, dh), (cw, ch), (cw, ch))):
                for y in range(h):
                    C.memmove(dst[p] + y * dstr[p], out[off + y * w:].ctypes.data, w)
                off += w * h
            return 0

        def free(k):
            o = self.ctxs.pop(k)[0]
            self.L.orc_scaler_free(o)

        self.cb = (_CREATE(create), _PROCESS(process), _FREE(free))

    def install(self, R):
        R.ref_set_scaler_callbacks(C.cast(self.cb[0], C.c_void_p), C.cast(self.cb[1], C.c_void_p), C.cast(self.cb[2], C.c_void_p))


def run_pixconv_sizeconv(g: RefGraph, frames, in_fmt: int, w: int, h: int, target=None, fps: float | None = None,
                         ticks_per_frame: int = 1, extra_ticks: int = 3, want_b200: bool | None = None):
    """frames: list of tight frames (uint8). MSPixConv(in_fmt, w x h) [-> MSSizeConv(target)] -> sink.
    Returns (frames out as one uint8 array, (tick, nbytes, timestamp) triples, (w, h) of every output header)."""
    src = g.source()
    pix = g.new("MSPixConv")
    if want_b200 is not None:
        assert g.text(pix).startswith("B200:") == want_b200, g.text(pix)
    g.call(pix, "MS_FILTER_SET_VIDEO_SIZE", VideoSize(w, h))
    g.call_int(pix, "MS_FILTER_SET_PIX_FMT", in_fmt)
    g.link(src, 0, pix, 0)
    last = pix
    if target is not None:
        sz = g.new("MSSizeConv")
        g.call(sz, "MS_FILTER_SET_VIDEO_SIZE", VideoSize(*target))
        if fps is not None:
            g.call_float(sz, "MS_FILTER_SET_FPS", fps)
        g.link(pix, 0, sz, 0)
        last = sz
    sink = g.sink()
    g.link(last, 0, sink, 0)
    for k, fr in enumerate(frames):
        g.push_video(src, k * ticks_per_frame, fr, w if in_fmt == MS_YUV420P else 0, h if in_fmt == MS_YUV420P else 0, 1000 + 90 * k)
    g.run(src, len(frames) * ticks_per_frame + extra_ticks)
    data, tri = g.read(sink, np.uint8)
    dims = g.read_dims(sink)
    g.close()
    return data, tri, dims


def synth_frame(fmt: int, w: int, h: int, t: int, seed: int = 0) -> np.ndarray:
    """a moving test pattern with per-pixel dither in the given packed / planar format (tight)"""
    rng = np.random.default_rng(seed * 1000 + t)
    yy, xx = np.mgrid[0:h, 0:w]
    if fmt == MS_YUV420P:
        y = ((xx + 2 * yy + 3 * t) % 256 + rng.integers(-3, 4, (h, w))).clip(0, 255).astype(np.uint8)
        cy, cx = np.mgrid[0:h // 2, 0:w // 2]
        u = ((cx + t) % 256).astype(np.uint8)
        v = ((cy + 2 * t) % 256).astype(np.uint8)
        return np.concatenate([y.reshape(-1), u.reshape(-1), v.reshape(-1)])
    if fmt in (MS_YUYV, MS_YUY2, MS_UYVY):
        y = ((xx * 3 + yy + 5 * t) % 256).astype(np.uint8)
        c = ((xx // 2 * 7 + yy * 2 + t) % 256).astype(np.uint8)
        out = np.empty((h, w, 2), np.uint8)
        out[..., 0 if fmt != MS_UYVY else 1] = y
        out[..., 1 if fmt != MS_UYVY else 0] = c
        return out.reshape(-1)
    bpp = 3 if fmt in (MS_RGB24, MS_RGB24_REV) else (2 if fmt == MS_RGB565 else 4)
    out = rng.integers(0, 256, (h, w, bpp)).astype(np.uint8)
    out[..., 0] = (xx * 2 + t * 9) % 256
    out[..., 1] = (yy * 3 + t) % 256
    return out.reshape(-1)
