"""GPU leg of the resampler anchors (tests/resample_anchor.py): the device bank behind msb200_resample_process gets the SAME
independent checks as the oracle — sine fit (unit gain, delay of exactly filt_len/2 input samples, rounding noise only),
correlation with scipy.signal.resample_poly and with the reference's own multi-rate recording, BASELINE cfg1 on the
hello8000 material — so the row does not rest on GPU == oracle alone (that bit-equality is tests/test_gpu_audio.py)."""
import numpy as np
import pytest

import resample_anchor as RA
from mediastreamer2_b200 import filters as F
from test_oracle_resample import CORPUS_MIN, SCIPY_MIN

pytestmark = pytest.mark.gpu


def gpu_resample(ctx, xs: np.ndarray, in_rate: int, out_rate: int) -> np.ndarray:
    """xs [streams][samples] -> [streams][out samples], fed in 10 ms blocks like the ticker does"""
    blk = in_rate // 100
    r = F.Resample(ctx, xs.shape[0], in_rate, out_rate, 1, blk)
    outs = [r.process(np.ascontiguousarray(xs[:, k:k + blk])) for k in range(0, xs.shape[1] - blk + 1, blk)]
    r.close()
    return np.concatenate(outs, axis=1)


def filt_len(in_rate: int, out_rate: int) -> int:  # the published rule, asserted for the oracle in test_oracle_resample.py
    return 48 if out_rate >= in_rate else ((48 * in_rate // out_rate - 1) & ~7) + 8


@pytest.mark.parametrize("in_rate,out_rate", RA.RATIOS)
def test_gpu_sine_fit(ctx, in_rate, out_rate):
    fracs = (0.05, 0.25, 0.5, 0.75, 0.85)
    fl = filt_len(in_rate, out_rate)
    got = {}

    def run(x):  # sine_fit builds one sine per call: collect, run as ONE bank of len(fracs) streams
        got.setdefault("x", []).append(x)
        return np.zeros(len(x) * out_rate // in_rate + 4096)

    with np.errstate(divide="ignore"):  # this pass only collects the inputs: the fit of an all-zero "output" is discarded
        for fr in fracs:
            RA.sine_fit(run, in_rate, out_rate, fl, fr)
    ys = gpu_resample(ctx, np.stack(got["x"]), in_rate, out_rate)
    for k, fr in enumerate(fracs):
        gain_db, delay_err, resid = RA.sine_fit(lambda x, k=k: ys[k], in_rate, out_rate, fl, fr)
        assert abs(delay_err) <= 0.001, (fr, delay_err)
        if fr <= 0.75:
            assert abs(gain_db) <= 0.002 and resid <= 2.0, (fr, gain_db, resid)
        else:
            lo, hi = (-0.6, -0.3) if out_rate >= in_rate else (-1.5, -1.0)
            assert lo <= gain_db <= hi, gain_db


@pytest.mark.parametrize("in_rate,out_rate", RA.RATIOS)
def test_gpu_speech_matches_scipy_and_the_reference_corpus(ctx, in_rate, out_rate):
    voice = np.load(RA.GOLDEN)
    x = voice[f"voice_{in_rate}"]
    y = gpu_resample(ctx, x[None, :], in_rate, out_rate)[0]
    fl = filt_len(in_rate, out_rate)
    a, b = RA.aligned(y, RA.scipy_resample(x, in_rate, out_rate), fl, in_rate, out_rate)
    assert RA.ncorr(a, b) >= SCIPY_MIN[(in_rate, out_rate)]
    a, c = RA.aligned(y, voice[f"voice_{out_rate}"], fl, in_rate, out_rate)
    assert RA.ncorr(a, c) >= CORPUS_MIN[(in_rate, out_rate)]


def test_gpu_cfg1_hello8000(ctx):
    x = np.load(RA.GOLDEN)["hello_8000"]
    y = gpu_resample(ctx, x[None, :], 8000, 48000)[0]
    assert len(y) == 6 * len(x)
    ys = RA.scipy_resample(x, 8000, 48000)
    a, b = RA.aligned(y, ys, 48, 8000, 48000)
    assert RA.ncorr(a, b) >= 0.9995
    assert np.abs(a.astype(np.float64) - b).max() <= 0.05 * np.abs(b).max()
