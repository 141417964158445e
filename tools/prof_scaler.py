"""One scaler geometry on device-resident frames, timed; the command ncu wraps for per-kernel captures.

    python tools/prof_scaler.py SRC_FMT SW SH DST_FMT DW DH [N_FRAMES] [PATH]     (formats: nv12 nv21 i420 rgb24 bgr24)
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

from mediastreamer2_b200 import _lib, filters as F

FMT = {"nv12": _lib.PIX_NV12, "nv21": _lib.PIX_NV21, "i420": _lib.PIX_YUV420P, "rgb24": _lib.PIX_RGB24, "bgr24": _lib.PIX_RGB24_REV}
a = sys.argv[1:]
sf, sw, sh, df, dw, dh = FMT[a[0]], int(a[1]), int(a[2]), FMT[a[3]], int(a[4]), int(a[5])
n = int(a[6]) if len(a) > 6 else 128
ctx = F.Context(0)
sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
if len(a) > 7:
    sc.set_path(int(a[7]))
d_src, d_dst = ctx.dev_alloc(n * sc.src_bytes), ctx.dev_alloc(n * sc.dst_bytes)
fr = np.random.default_rng(0).integers(0, 256, size=(8, sc.src_bytes), dtype=np.uint8)
for i in range(0, n, 8):
    ctx.h2d(d_src + i * sc.src_bytes, fr[: min(8, n - i)])
for _ in range(3):
    sc.process_dev(n, d_src, d_dst)
ctx.sync()
ctx.timer_start()
for _ in range(10):
    sc.process_dev(n, d_src, d_dst)
ms = ctx.timer_stop_ms() / 10
b = n * (sc.src_bytes + sc.dst_bytes)
print(f"path {sc.path}: {ms:.4f} ms per {n} frames, {b / (ms / 1e3) / 1e9:.1f} GB/s on {b} algorithmic bytes")
