import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from mediastreamer2_b200 import filters as F, _lib
ctx = F.Context(0)
n = 512
sc = F.Scaler(ctx, 1920, 1080, _lib.PIX_NV12, 1280, 720, _lib.PIX_RGB24)
d_src = ctx.dev_alloc(n * sc.src_bytes); d_dst = ctx.dev_alloc(n * sc.dst_bytes)
rng = np.random.default_rng(0)
fr = rng.integers(0, 256, size=(8, sc.src_bytes), dtype=np.uint8)
for i in range(0, n, 8): ctx.h2d(d_src + i * sc.src_bytes, fr)
for path in [int(a) for a in sys.argv[1:]] or (0, 3, 1):
    sc.set_path(path)
    for _ in range(3): sc.process_dev(n, d_src, d_dst)
    ctx.sync(); ctx.timer_start()
    for _ in range(10): sc.process_dev(n, d_src, d_dst)
    ms = ctx.timer_stop_ms() / 10
    print(f"path {path} kind {sc.path}: {ms:.4f} ms  frac {5875200*n/(ms/1e3)/1e9/6455.9:.3f}")
