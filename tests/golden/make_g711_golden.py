"""Golden tables for G.711: the UNMODIFIED Snack_Lin2Alaw / Snack_Lin2Mulaw / Snack_Alaw2Lin / Snack_Mulaw2Lin of the
reference (src/audiofilters/g711.c, compiled into oracle/_ref/libms2ref.so) evaluated on EVERY 16-bit sample and every code
word. Committed so that the exhaustive pin travels with the repository. Run in the build container:

    python tests/golden/make_g711_golden.py
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))

import _oracle as O  # noqa: E402

if __name__ == "__main__":
    R = O.ref()
    out = {}
    for name, enc, dec in (("alaw", R.Snack_Lin2Alaw, R.Snack_Alaw2Lin), ("ulaw", R.Snack_Lin2Mulaw, R.Snack_Mulaw2Lin)):
        enc.restype, enc.argtypes = C.c_ubyte, [C.c_short]
        dec.restype, dec.argtypes = C.c_short, [C.c_ubyte]
        out[f"{name}_enc"] = np.array([enc(v) for v in range(-32768, 32768)], np.uint8)  # index = sample + 32768
        out[f"{name}_dec"] = np.array([dec(c) for c in range(256)], np.int16)
    np.savez_compressed(HERE / "g711_reference.npz", **out)
    print((HERE / "g711_reference.npz").stat().st_size, "bytes")
