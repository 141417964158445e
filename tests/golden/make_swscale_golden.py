"""Generates golden frames with the REAL libswscale (9.1.100, FFmpeg 8) found inside the opencv wheel of this container
— the library the reference's ffmpeg scaler back-end calls (/root/reference/src/voip/msvideo.c:651-681:
sws_getContext(..., SWS_BILINEAR, NULL, NULL, NULL) + sws_scale()). Run here (the wheel may not exist on the GPU box);
the small .npz files it writes are committed and pin oracle/oracle_video.c (tests/test_oracle_video.py).

    python tests/golden/make_swscale_golden.py
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
AV_PIX = {"yuv420p": 0, "yuyv422": 1, "rgb24": 2, "bgr24": 3, "uyvy422": 15, "nv12": 23, "nv21": 24, "rgba": 26, "bgra": 28,
          "rgb565le": 37}
SWS_BILINEAR = 2
SWS_BITEXACT = 0x80000  # same algorithm, the library's C reference functions instead of its approximate x86 SIMD ones


def load_swscale():
    libdir = None
    for p in sys.path:
        cand = glob.glob(os.path.join(p, "opencv_python_headless.libs", "libswscale-*.so*"))
        if cand:
            libdir = os.path.dirname(cand[0])
            break
    if libdir is None:
        raise RuntimeError("libswscale not found (opencv_python_headless.libs)")
    # the wheel's private dependencies (libavutil, libdrm, ...) resolve through LD_LIBRARY_PATH: re-exec once with it
    if libdir not in os.environ.get("LD_LIBRARY_PATH", "").split(":"):
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = libdir + ":" + env.get("LD_LIBRARY_PATH", "")
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    L = C.CDLL(glob.glob(os.path.join(libdir, "libswscale-*.so*"))[0])
    L.sws_getContext.restype = C.c_void_p
    L.sws_getContext.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sws_scale.restype = C.c_int
    L.sws_scale.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.sws_freeContext.argtypes = [C.c_void_p]
    L.swscale_version.restype = C.c_uint
    return L


def planes(fmt: str, w: int, h: int, buf: np.ndarray):
    base = buf.ctypes.data
    cw, ch = (w + 1) // 2, (h + 1) // 2
    if fmt in ("rgb24", "bgr24"):
        return [base, 0, 0, 0], [w * 3, 0, 0, 0]
    if fmt in ("rgba", "bgra"):
        return [base, 0, 0, 0], [w * 4, 0, 0, 0]
    if fmt in ("yuyv422", "uyvy422", "rgb565le"):
        return [base, 0, 0, 0], [w * 2, 0, 0, 0]
    if fmt in ("nv12", "nv21"):
        return [base, base + w * h, 0, 0], [w, cw * 2, 0, 0]
    return [base, base + w * h, base + w * h + cw * ch, 0], [w, cw, cw, 0]


def nbytes(fmt: str, w: int, h: int) -> int:
    if fmt in ("rgb24", "bgr24"):
        return w * h * 3
    if fmt in ("rgba", "bgra"):
        return w * h * 4
    if fmt in ("yuyv422", "uyvy422", "rgb565le"):
        return w * h * 2
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


def sws_convert(L, src: np.ndarray, sfmt: str, sw: int, sh: int, dfmt: str, dw: int, dh: int, flags: int = SWS_BILINEAR) -> np.ndarray:
    ctx = L.sws_getContext(sw, sh, AV_PIX[sfmt], dw, dh, AV_PIX[dfmt], flags, None, None, None)
    assert ctx
    dst = np.zeros(nbytes(dfmt, dw, dh) + 64, np.uint8)
    sp, ss = planes(sfmt, sw, sh, src)
    dp, ds = planes(dfmt, dw, dh, dst)
    r = L.sws_scale(ctx, (C.c_void_p * 4)(*sp), (C.c_int * 4)(*ss), 0, sh, (C.c_void_p * 4)(*dp), (C.c_int * 4)(*ds))
    assert r == dh, r
    L.sws_freeContext(ctx)
    return dst[: nbytes(dfmt, dw, dh)].copy()


def test_frame(fmt: str, w: int, h: int, t: int, seed: int) -> np.ndarray:
    """BASELINE cfg4 pattern: Y=(x+2y+3t)%256, U=(x/2+t)%256, V=(y/2+2t)%256 plus +-3 dither."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    Y = ((xx + 2 * yy + 3 * t) % 256 + rng.integers(-3, 4, size=(h, w))).clip(0, 255).astype(np.uint8)
    ch, cw = (h + 1) // 2, (w + 1) // 2
    cy, cx = np.mgrid[0:ch, 0:cw]
    U = ((cx + t) % 256 + rng.integers(-3, 4, size=(ch, cw))).clip(0, 255).astype(np.uint8)
    V = ((cy + 2 * t) % 256 + rng.integers(-3, 4, size=(ch, cw))).clip(0, 255).astype(np.uint8)
    if fmt in ("rgb24", "bgr24", "rgba", "bgra"):  # MSPixConv's RGB inputs: a colour ramp plus full-range noise in one quadrant
        nc = 4 if fmt in ("rgba", "bgra") else 3
        rgb = np.stack([(3 * xx + t) % 256, (2 * yy + 5 * t) % 256, (xx + yy) % 256, (xx * yy) % 256][:nc], axis=-1).astype(np.uint8)
        rgb[: h // 2, : w // 2] = rng.integers(0, 256, size=(h // 2, w // 2, nc), dtype=np.uint8)
        return rgb.ravel()
    if fmt == "yuv420p":
        return np.concatenate([Y.ravel(), U.ravel(), V.ravel()])
    if fmt in ("yuyv422", "uyvy422"):  # 4:2:2 packed: chroma at full vertical resolution
        U2 = rng.integers(0, 256, size=(h, cw), dtype=np.uint8)
        V2 = rng.integers(0, 256, size=(h, cw), dtype=np.uint8)
        out = np.zeros((h, w * 2), np.uint8)
        yo, uo = (0, 1) if fmt == "yuyv422" else (1, 0)
        out[:, yo::2] = Y
        out[:, uo::4] = U2
        out[:, uo + 2::4] = V2
        return out.ravel()
    a, b = (U, V) if fmt == "nv12" else (V, U)
    return np.concatenate([Y.ravel(), np.stack([a, b], axis=-1).ravel()])


CASES = [
    # (src fmt, sw, sh, dst fmt, dw, dh)
    ("nv12", 192, 108, "rgb24", 128, 72),      # cfg4 shape at 1/10 scale: 1.5x down both ways
    ("nv12", 160, 120, "rgb24", 64, 48),       # 2.5x down
    ("nv12", 64, 48, "rgb24", 96, 72),         # 1.5x up
    ("nv21", 96, 64, "bgr24", 64, 48),
    ("yuv420p", 192, 108, "yuv420p", 128, 72), # MSSizeConv path (I420 -> I420)
    ("yuv420p", 64, 48, "yuv420p", 160, 120),
    ("yuv420p", 176, 144, "rgb24", 128, 96),
    ("nv12", 128, 72, "rgb24", 128, 72),       # same size through the generic (unscaled) path
    ("nv12", 130, 74, "rgb24", 86, 50),        # odd chroma geometry
    ("yuyv422", 96, 64, "yuv420p", 96, 64),    # MSPixConv: packed 4:2:2 -> I420, same size
    ("uyvy422", 64, 48, "yuv420p", 64, 48),
    ("rgb24", 96, 64, "yuv420p", 96, 64),      # MSPixConv: MS_RGB24 -> I420 (generic scaler path with the RGB input stage)
    ("bgr24", 64, 48, "yuv420p", 64, 48),      # MSPixConv: MS_RGB24_REV -> I420 (unscaled special converter rgb24toyv12)
    ("rgb24", 132, 70, "yuv420p", 132, 70),    # width % 8 != 0
    ("rgba", 96, 64, "yuv420p", 96, 64),       # MSPixConv: MS_RGBA32 -> I420
    ("bgra", 64, 48, "yuv420p", 64, 48),       # MSPixConv: MS_RGBA32_REV -> I420
]


def main():
    L = load_swscale()
    ver = L.swscale_version()
    print(f"libswscale {ver >> 16}.{(ver >> 8) & 255}.{ver & 255}")
    out = {}
    for k, (sf, sw, sh, df, dw, dh) in enumerate(CASES):
        src = test_frame(sf, sw, sh, t=k, seed=1000 + k)
        dst = sws_convert(L, src, sf, sw, sh, df, dw, dh)
        out[f"case{k}_src"] = src
        out[f"case{k}_dst"] = dst
        out[f"case{k}_dst_bitexact"] = sws_convert(L, src, sf, sw, sh, df, dw, dh, SWS_BILINEAR | SWS_BITEXACT)
        out[f"case{k}_meta"] = np.array([AV_PIX[sf], sw, sh, AV_PIX[df], dw, dh], np.int32)
    # full-size BASELINE cfg4 frame (NV12 1080p -> RGB24 720p): too large to commit, pinned by its SHA-256
    import hashlib
    src = test_frame("nv12", 1920, 1080, t=0, seed=4242)
    dst = sws_convert(L, src, "nv12", 1920, 1080, "rgb24", 1280, 720)
    out["cfg4_sha256"] = np.frombuffer(hashlib.sha256(dst.tobytes()).digest(), np.uint8)
    out["swscale_version"] = np.array([ver], np.uint32)
    np.savez_compressed(HERE / "swscale_bilinear.npz", **out)
    print("wrote", HERE / "swscale_bilinear.npz", (HERE / "swscale_bilinear.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
