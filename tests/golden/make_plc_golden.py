"""Golden vectors for MSGenericPLC: the UNMODIFIED reference filter (oracle/_ref/libms2ref.so = the reference's own
msgenericplc.c, genericplc.c, kiss_fft.c, kiss_fftr.c, dsptools.c compiled from /root/reference) run in the reference's
MSTicker over seeded lossy streams. The outputs are committed so that the pin travels with the repository (the GPU box and a
fresh clone have no reference tree). Run in the build container:

    python tests/golden/make_plc_golden.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))

from test_oracle_vs_reference import plc_reference_run, plc_schedule, plc_signal  # noqa: E402

CASES = {  # name: (rate, ticks, lost blocks, block_ms, comfort-noise requests before these ticks, signal seed)
    "r8000": (8000, 70, sorted(set(range(10, 14)) | {20} | set(range(30, 52)) | {60, 61}), 10, (), 1),
    "r16000": (16000, 70, sorted(set(range(10, 14)) | {20} | set(range(30, 52)) | {60, 61}), 10, (), 2),
    "r48000": (48000, 40, sorted({5} | set(range(12, 24))), 10, (), 3),
    "r16000_cn": (16000, 50, sorted(set(range(12, 20)) | set(range(30, 34))), 10, (12,), 4),
    "r16000_20ms": (16000, 60, sorted({3, 7, 8} | set(range(12, 19))), 20, (), 5),
}

if __name__ == "__main__":
    out = {}
    for name, (rate, ticks, lost, block_ms, cn_at, seed) in CASES.items():
        x = plc_signal(rate, ticks * rate // 100, seed=seed)
        pcm, blocks = plc_reference_run(rate, ticks, plc_schedule(rate, ticks, set(lost), block_ms), x, cn_at)
        out[f"{name}_in"] = x
        out[f"{name}_out"] = pcm
        out[f"{name}_sizes"] = np.array([b for _, b in blocks], np.int32)
    np.savez_compressed(HERE / "plc_reference.npz", **out)
    print((HERE / "plc_reference.npz").stat().st_size, "bytes")
