"""512-camera shape of the video boundary in the UNMODIFIED reference MSTicker, built from the plugin's B200 filters:

    per stream:  camera source (YUY2 w x h, one frame every `period` ticks) -> MSPixConv -> MSSizeConv(w/2 x h/2) -> sink

    MSB200_BATCH=0|<slots> [MSB200_VIDEO_BATCH=<frames per flush>] python tests/video_runner.py --streams 512 --tickers 8

Prints one JSON line: wall time per tick of the free-running (gated) tickers, frames converted, lane flushes (= batched
launch sequences). In batch mode the MSPixConv / MSSizeConv instances of one ticker share a conversion lane each: a tick's
frames go up in one copy, through one launch sequence, and come back straight into the pinned blocks handed downstream."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import _oracle as O  # noqa: E402
import video_graph as V  # noqa: E402
from _oracle import RefGraph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--tickers", type=int, default=4)
    ap.add_argument("--ticks", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--period", type=int, default=3, help="a frame every `period` ticks per stream (3 = 33 fps at 10 ms)")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    a = ap.parse_args()
    w, h = a.width, a.height
    g = RefGraph(plugins_dir=str(O.PLUGIN_DIR))
    frames = [V.synth_frame(V.MS_YUY2, w, h, t) for t in range(3)]
    sources, sinks = [], []
    for s in range(a.streams):
        src, pix, sz, sink = g.source(), g.new("MSPixConv"), g.new("MSSizeConv"), g.sink()
        assert g.text(pix).startswith("B200:") and g.text(sz).startswith("B200:")
        g.call(pix, "MS_FILTER_SET_VIDEO_SIZE", V.VideoSize(w, h))
        g.call_int(pix, "MS_FILTER_SET_PIX_FMT", V.MS_YUY2)
        g.call(sz, "MS_FILTER_SET_VIDEO_SIZE", V.VideoSize(w // 2, h // 2))
        g.link(src, 0, pix, 0)
        g.link(pix, 0, sz, 0)
        g.link(sz, 0, sink, 0)
        for k, t in enumerate(range(s % a.period, a.ticks, a.period)):  # cameras are not in phase
            g.push_video(src, t, frames[(s + k) % len(frames)], 0, 0, 90 * t)
        g.L.ref_sink_set_discard(sink, 1)
        sources.append(src)
        sinks.append(sink)
    tickers = [g.L.ref_ticker_new() for _ in range(max(1, a.tickers))]
    for r, src in enumerate(sources):
        assert g.L.ref_ticker_attach(tickers[r % len(tickers)], src) == 0

    def run(n):
        for t in tickers:
            g.L.ref_ticker_release(t, n)
        for t in tickers:
            g.L.ref_ticker_wait(t)

    run(a.warmup)
    per_tick = []
    for _ in range(a.ticks - a.warmup):
        t0 = time.perf_counter()
        run(1)
        per_tick.append(time.perf_counter() - t0)
    run(4)  # drain the lanes
    ms = np.array(per_tick) * 1000.0
    out_frames = sum(g.L.ref_sink_nblocks(k) for k in sinks)
    stats = {"mode": "batch" if int(os.environ.get("MSB200_BATCH", "0") or 0) > 0 else "sync", "streams": a.streams,
             "tickers": len(tickers), "source": f"YUY2 {w}x{h}", "target": f"I420 {w // 2}x{h // 2}", "period_ticks": a.period,
             "ticks_timed": len(ms), "tick_ms_mean": float(ms.mean()), "tick_ms_p50": float(np.percentile(ms, 50)),
             "tick_ms_p99": float(np.percentile(ms, 99)), "tick_ms_max": float(ms.max()), "late_ticks_10ms": int((ms > 10.0).sum()),
             "frames_out": int(out_frames), "frames_per_s": float(a.streams / a.period / (ms.mean() / 1000.0))}
    try:
        plug = C.CDLL(str(O.PLUGIN_DIR / "libmsb200filters.so"))
        fl, fr = C.c_ulonglong(), C.c_ulonglong()
        plug.msb200_filters_video_stats(C.byref(fl), C.byref(fr))
        stats.update({"lane_flushes": fl.value, "frames_converted": fr.value})
    except (OSError, AttributeError):
        pass
    print(json.dumps(stats), flush=True)
    for r, src in enumerate(sources):
        g.L.ref_ticker_detach(tickers[r % len(tickers)], src)
    for t in tickers:
        g.L.ref_ticker_destroy(t)
    g.close()


if __name__ == "__main__":
    main()
