"""N lossy streams, each `source (with holes) -> MSGenericPLC -> sink`, on one unmodified reference MSTicker, built from the
plugin's B200 filter. The plugin's execution mode is fixed per process by MSB200_BATCH, so tests spawn this script twice:

    MSB200_BATCH=0|<slots> python tests/plc_runner.py --streams 12 --ticks 80 --rate 16000 --dump out.npz

Stream s loses the blocks LOSS[s % len(LOSS)] (tests/test_gpu_plc.py); stream 1 also gets MS_GENERIC_PLC_SET_CN before
its hole. --dump writes every sink's sample stream and block sizes."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import _oracle as O  # noqa: E402
from _oracle import RefGraph  # noqa: E402

LOSS = [
    set(range(10, 14)) | {20} | set(range(30, 52)) | {60, 61},
    {5} | set(range(40, 46)),
    set(),
    set(range(3, 70)),
    {8, 10, 12, 14, 16, 18},
    set(range(25, 36)) | set(range(40, 44)),
]


def main():
    from test_oracle_vs_reference import _CngData, plc_signal
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=12)
    ap.add_argument("--ticks", type=int, default=80)
    ap.add_argument("--rate", type=int, default=16000)
    ap.add_argument("--dump", default="")
    a = ap.parse_args()
    n = a.rate // 100
    g = RefGraph(plugins_dir=str(O.PLUGIN_DIR))
    sources, plcs, sinks = [], [], []
    for s in range(a.streams):
        x = plc_signal(a.rate, a.ticks * n, seed=50 + s)
        src, plc, sink = g.source(), g.new("MSGenericPLC"), g.sink()
        assert g.text(plc).startswith("B200:")
        g.call_int(plc, "MS_FILTER_SET_SAMPLE_RATE", a.rate)
        g.call_int(plc, "MS_FILTER_SET_NCHANNELS", 1)
        for k in range(a.ticks):
            if k not in LOSS[s % len(LOSS)]:
                g.push(src, k, x[k * n:(k + 1) * n])
        g.link(src, 0, plc, 0)
        g.link(plc, 0, sink, 0)
        sources.append(src)
        plcs.append(plc)
        sinks.append(sink)
    g.run(sources, 40)
    if a.streams > 1:
        assert g.call(plcs[1], "MS_GENERIC_PLC_SET_CN", _CngData()) == 0  # stream 1's hole at 40..45 becomes comfort noise
    g.run(sources, a.ticks - 40)
    if a.dump:
        out = {}
        for i, k in enumerate(sinks):
            pcm, tri = g.read(k)
            out[f"pcm{i}"] = pcm
            out[f"sizes{i}"] = tri[:, 1]
        np.savez(a.dump, **out)
    stats = {}
    try:
        plug = C.CDLL(str(O.PLUGIN_DIR / "libmsb200filters.so"))
        gr, fl, un = C.c_int(), C.c_ulonglong(), C.c_ulonglong()
        plug.msb200_filters_batch_stats(C.byref(gr), C.byref(fl), C.byref(un))
        stats = {"batch_groups": gr.value, "batch_launches": fl.value, "batch_units": un.value}
    except (OSError, AttributeError):
        pass
    print(json.dumps(stats), flush=True)
    g.close()


if __name__ == "__main__":
    main()
