"""The reference's "simple talk" echo-canceller scenario (tests/golden/aec_simple_talk.npz, cut from the reference's own
test sounds by tests/golden/make_aec_fixture.py): far-end and near-end talkers alternate; the microphone hears the
near-end talker plus a full-level echo of the far end delayed by ~31 ms. Shared by the CPU (oracle) and GPU tests."""
from pathlib import Path

import numpy as np

FIX = Path(__file__).resolve().parent / "golden" / "aec_simple_talk.npz"
RATE = 16000
FAR_ONLY_SECONDS = (9, 10)   # far-end talker alone, after the filter has converged (seconds 0-1 and 5-6 train it)
NEAR_ONLY_SECONDS = (3, 4)   # near-end talker alone


def load():
    g = np.load(FIX)
    far, echo, near = g["farend"], g["echo"], g["nearend"]
    mic = np.clip(near.astype(np.int32) + echo.astype(np.int32), -32768, 32767).astype(np.int16)
    return far, mic, near


def db(x):
    return 10 * np.log10(np.mean(x.astype(np.float64) ** 2) + 1.0)


def second(x, s):
    return x[s * RATE:(s + 1) * RATE]


def best_corr(a, b, max_lag=400):
    """peak normalised correlation of a against b delayed by 0..max_lag samples (the preprocessor's overlap-add delay)"""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    best = -1.0
    for lag in range(0, max_lag, 8):
        x, y = a[lag:], b[:len(b) - lag]
        c = float(np.dot(x, y) / (np.linalg.norm(x) * np.linalg.norm(y) + 1e-9))
        best = max(best, c)
    return best


def check_behaviour(out, mic, near, min_erle_db=20.0):
    """what an echo canceller must do on this material; returns the measured numbers"""
    erle = [db(second(mic, s)) - db(second(out, s)) for s in FAR_ONLY_SECONDS]
    keep = [db(second(out, s)) - db(second(near, s)) for s in NEAR_ONLY_SECONDS]
    corr = [best_corr(second(out, s), second(near, s)) for s in NEAR_ONLY_SECONDS]
    assert min(erle) >= min_erle_db, erle          # the echo of the lone far-end talker is removed
    assert max(abs(k) for k in keep) <= 4.0, keep  # the lone near-end talker passes at (almost) full level
    assert min(corr) >= 0.8, corr                  # ... and undistorted
    return erle, keep, corr
