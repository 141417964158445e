"""G.711 block of the bench line (the decode / encode stubs of BASELINE cfg5 made real): A-law decode and
encode over (a) one 10 ms tick of 4096 x 8 kHz streams (80 samples each: the media-server shape, launch-latency bound)
and (b) a batch far larger than L2 (roofline: 3 algorithmic bytes per sample). Imported by bench.py; runnable alone:

    python bench_g711.py
"""
from __future__ import annotations

import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def g711_bench(ctx, hbm_peak_gbs: float, cpu_baseline: bool = True) -> dict:
    from mediastreamer2_b200 import filters as F

    out = {"bytes_per_sample": 3, "law": "alaw"}
    rng = np.random.default_rng(7)
    for tag, n, iters in (("tick_4096_streams", 4096 * 80, 200), ("batch_256M_samples", 256 << 20, 10)):
        d_code, d_pcm = ctx.dev_alloc(n), ctx.dev_alloc(2 * n)
        chunk = rng.integers(0, 256, min(n, 1 << 22)).astype(np.uint8)
        for off in range(0, n, chunk.size):
            ctx.h2d(d_code + off, chunk[: min(chunk.size, n - off)])
        res = {"samples": n}
        for op, fn, a, b in (("decode", F.g711_decode_dev, d_code, d_pcm), ("encode", F.g711_encode_dev, d_pcm, d_code)):
            for _ in range(3):
                fn(ctx, F.G711_ALAW, a, b, n)
            ctx.sync()
            ctx.timer_start()
            for _ in range(iters):
                fn(ctx, F.G711_ALAW, a, b, n)
            ms = ctx.timer_stop_ms() / iters
            gbs = 3.0 * n / (ms / 1000.0) / 1e9
            res[op] = {"ms": ms, "gsamples_per_s": n / (ms / 1000.0) / 1e9, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak_gbs}
        ctx.dev_free(d_code)
        ctx.dev_free(d_pcm)
        out[tag] = res
    if cpu_baseline:  # the oracle's scalar loop on one host core, bounded sample
        import _oracle as O

        L = O.oracle()
        n = 8 << 20
        pcm = rng.integers(-32768, 32768, n).astype(np.int16)
        code = np.zeros(n, np.uint8)
        t0 = time.perf_counter()
        L.orc_g711_encode(0, pcm.ctypes.data_as(C.c_void_p), code.ctypes.data_as(C.c_void_p), n)
        L.orc_g711_decode(0, code.ctypes.data_as(C.c_void_p), pcm.ctypes.data_as(C.c_void_p), n)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"gsamples_per_s": 2 * n / dt / 1e9, "cores": 1, "kind": "port",
                               "sample": "8 Mi samples encode + decode, oracle/oracle_g711.c, one thread"}
    return out


if __name__ == "__main__":
    from mediastreamer2_b200 import filters as F

    c = F.Context(0)
    print(json.dumps(g711_bench(c, 6455.9)))
    c.close()
