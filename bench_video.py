"""pixconv half of the headline metric (BASELINE cfg4): NV12 1080p -> RGB24 720p, fused MSPixConv+MSSizeConv arithmetic,
batched over concurrent video streams on one B200. Imported by bench.py; runnable alone for profiling:

    python bench_video.py [--frames 512] [--iters 20]
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SRC_W, SRC_H, DST_W, DST_H = 1920, 1080, 1280, 720
SRC_BYTES = SRC_W * SRC_H * 3 // 2   # 3,110,400
DST_BYTES = DST_W * DST_H * 3        # 2,764,800
ALGO_BYTES = SRC_BYTES + DST_BYTES   # 5,875,200 per frame


def _load_libswscale():
    """The REAL libswscale (what the reference's ffmpeg scaler back-end calls, src/voip/msvideo.c:651-681) from the opencv
    wheel of this image, with its private dependency chain loaded by absolute path. None when the wheel is absent."""
    import ctypes as C
    import glob
    import os

    d = None
    for p in sys.path:
        c = glob.glob(os.path.join(p, "opencv_python_headless.libs", "libswscale-*.so*"))
        if c:
            d = os.path.dirname(c[0])
            break
    if d is None:
        return None
    try:
        for name in ("libcrypto", "libssl", "libdrm", "libavutil"):
            for f in sorted(glob.glob(os.path.join(d, name + "-*.so*"))):
                C.CDLL(f, mode=C.RTLD_GLOBAL)
        L = C.CDLL(glob.glob(os.path.join(d, "libswscale-*.so*"))[0])
    except OSError:
        return None
    L.sws_getContext.restype = C.c_void_p
    L.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3
    L.sws_scale.restype = C.c_int
    L.sws_scale.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p),
                            C.POINTER(C.c_int)]
    L.sws_freeContext.argtypes = [C.c_void_p]
    L.swscale_version.restype = C.c_uint
    return L


def pixconv_cpu_baseline(frame: np.ndarray, budget_s: float = 8.0):
    """cfg4 on the host cores through the reference's own library call (sws_getContext(..., SWS_BILINEAR) + sws_scale),
    one context per thread, all cores, bounded to ~budget_s of wall time. Returns (dict, first output frame) or None."""
    import ctypes as C
    import os
    import threading
    import time

    L = _load_libswscale()
    if L is None:
        return None
    threads = os.cpu_count() or 1
    ver = L.swscale_version()
    outs = [None] * threads

    def work(k, n):
        ctx = L.sws_getContext(SRC_W, SRC_H, 23, DST_W, DST_H, 2, 2, None, None, None)  # NV12 -> RGB24, SWS_BILINEAR
        dst = np.zeros(DST_BYTES + 64, np.uint8)
        sp = (C.c_void_p * 4)(frame.ctypes.data, frame.ctypes.data + SRC_W * SRC_H, 0, 0)
        ss = (C.c_int * 4)(SRC_W, SRC_W, 0, 0)
        dp = (C.c_void_p * 4)(dst.ctypes.data, 0, 0, 0)
        ds = (C.c_int * 4)(DST_W * 3, 0, 0, 0)
        for _ in range(n):
            L.sws_scale(ctx, sp, ss, 0, SRC_H, dp, ds)
        L.sws_freeContext(ctx)
        outs[k] = dst[:DST_BYTES]

    def run(n):
        th = [threading.Thread(target=work, args=(k, n)) for k in range(threads)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    dt = run(2)  # calibrate
    n = max(2, min(200, int(budget_s / max(dt / 2, 1e-3))))
    dt = run(n)
    fps = threads * n / dt
    return ({"mpix_per_s_in": SRC_W * SRC_H * fps / 1e6, "frames_per_s": fps, "cores": threads,
             "kind": f"reference: libswscale {ver >> 16}.{(ver >> 8) & 255}.{ver & 255} (the library the reference's ffmpeg scaler back-end calls)",
             "sample": f"{threads} threads x {n} frames, one SwsContext per thread, NV12 {SRC_W}x{SRC_H} -> RGB24 {DST_W}x{DST_H}, SWS_BILINEAR"},
            outs[0])


def pixconv_bench(ctx, hbm_peak_gbs: float, n_frames: int = 512, iters: int = 10, cpu_baseline: bool = True) -> dict:
    from mediastreamer2_b200 import _lib
    from mediastreamer2_b200 import filters as F

    sc = F.Scaler(ctx, SRC_W, SRC_H, _lib.PIX_NV12, DST_W, DST_H, _lib.PIX_RGB24)
    # cfg4 synthetic frames: 8 distinct frames tiled over the batch (1.6 GB in, 1.4 GB out: far larger than L2)
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:SRC_H, 0:SRC_W]
    base = []
    for t in range(8):
        Y = ((xx + 2 * yy + 3 * t) % 256 + rng.integers(-3, 4, size=(SRC_H, SRC_W))).clip(0, 255).astype(np.uint8)
        cy, cx = np.mgrid[0:SRC_H // 2, 0:SRC_W // 2]
        U = ((cx + t) % 256).astype(np.uint8)
        V = ((cy + 2 * t) % 256).astype(np.uint8)
        base.append(np.concatenate([Y.ravel(), np.stack([U, V], axis=-1).ravel()]))
    base = np.stack(base)
    d_src = ctx.dev_alloc(n_frames * SRC_BYTES)
    d_dst = ctx.dev_alloc(n_frames * DST_BYTES)
    for i in range(0, n_frames, 8):
        k = min(8, n_frames - i)
        ctx.h2d(d_src + i * SRC_BYTES, base[:k])
    for _ in range(3):
        sc.process_dev(n_frames, d_src, d_dst)
    ctx.sync()
    ctx.timer_start()
    for _ in range(iters):
        sc.process_dev(n_frames, d_src, d_dst)
    ms = ctx.timer_stop_ms() / iters
    # end to end for a smaller batch through host buffers (pinned)
    ne = min(64, n_frames)
    src_pin = ctx.pinned((ne, SRC_BYTES), np.uint8)
    src_pin[...] = base[np.arange(ne) % 8]
    import time

    dst_pin = ctx.pinned((ne, DST_BYTES), np.uint8)
    sc.process(src_pin, dst_pin)
    t0 = time.perf_counter()
    out = sc.process(src_pin, dst_pin)
    e2e_s = time.perf_counter() - t0
    ctx.dev_free(d_src)
    ctx.dev_free(d_dst)
    sc.close()
    achieved = ALGO_BYTES * n_frames / (ms / 1000.0) / 1e9
    cpu = None
    if cpu_baseline:
        got = pixconv_cpu_baseline(np.ascontiguousarray(base[0]))
        if got is not None:
            cpu, ref_frame = got
            cpu["gpu_frame_equals_library_frame"] = bool(np.array_equal(out[0], ref_frame))  # same input frame: bit-exact
    res = {
        "workload": f"cfg4: {n_frames} concurrent streams, NV12 {SRC_W}x{SRC_H} -> RGB24 {DST_W}x{DST_H}, one frame each per launch",
        "mpix_per_s_in": SRC_W * SRC_H * n_frames / (ms / 1000.0) / 1e6,
        "mpix_per_s_out": DST_W * DST_H * n_frames / (ms / 1000.0) / 1e6,
        "ms_per_batch": ms, "frames_per_s": n_frames / (ms / 1000.0),
        "roofline": {"bound": "hbm", "kernel": "scale_rgb_strip_kernel (static row schedule)", "achieved": achieved, "peak": hbm_peak_gbs,
                     "unit": "GB/s", "frac": achieved / hbm_peak_gbs, "bytes_per_frame": ALGO_BYTES},
        "e2e": {"frames": ne, "mpix_per_s_in": SRC_W * SRC_H * ne / e2e_s / 1e6,
                "h2d_bytes": ne * SRC_BYTES, "d2h_bytes": ne * DST_BYTES},
        "checksum": int(out[:2].astype(np.uint64).sum()),
    }
    if cpu is not None:
        res["cpu_baseline"] = cpu
    return res


if __name__ == "__main__":
    import argparse

    from mediastreamer2_b200 import filters as F

    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    c = F.Context(0)
    print(json.dumps(pixconv_bench(c, 6577.0, a.frames, a.iters)))
    c.close()
