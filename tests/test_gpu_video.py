"""GPU parity for the video rows: NV12 -> I420 (bit-exact vs oracle, itself pinned bit-exact vs the reference)."""
import numpy as np
import pytest

import _oracle as O
from _oracle import ptr
from mediastreamer2_b200 import filters as F

pytestmark = pytest.mark.gpu


def _frames(n, y_stride, c_stride, sh, seed):
    rng = np.random.default_rng(seed)
    fb = y_stride * sh + c_stride * (sh // 2)
    return rng.integers(0, 256, size=(n, fb), dtype=np.uint8)


@pytest.mark.parametrize("rotation", [0, 90, 180, 270])
@pytest.mark.parametrize("down_scale", [False, True])
@pytest.mark.parametrize("padded", [False, True])
def test_nv12_to_i420_bit_exact(ctx, rotation, down_scale, padded):
    L = O.oracle()
    f = 2 if down_scale else 1
    sw, sh = 640, 480
    w, h = (sw // f, sh // f) if rotation % 180 == 0 else (sh // f, sw // f)
    ys = sw + (sw % 32 + 32 if padded else 0)
    cs = sw + (64 if padded else 0)
    frames = _frames(3, ys, cs, sh, 7)
    for u_first in (True, False):
        got = F.nv12_to_i420(ctx, frames, w, h, rotation, ys, cs, u_first, down_scale)
        for i in range(frames.shape[0]):
            exp = np.zeros(w * h * 3 // 2, np.uint8)
            y = frames[i, :ys * sh]
            c = frames[i, ys * sh:]
            L.orc_nv12_to_i420(ptr(y), ptr(c), rotation, w, h, ys, cs, int(u_first), int(down_scale), ptr(exp))
            assert np.array_equal(got[i], exp), (i, u_first)


def test_nv12_reference_test_pattern_1080p(ctx):
    """The reference's own pattern (y[i]=i%256, cbcr[i]=i%256; framework_tester.c:219-367) at 1080p, fast path."""
    L = O.oracle()
    w, h = 1920, 1080
    y = (np.arange(w * h) % 256).astype(np.uint8)
    c = (np.arange(w * h // 2) % 256).astype(np.uint8)
    frame = np.concatenate([y, c])[None, :]
    got = F.nv12_to_i420(ctx, frame, w, h)
    exp = np.zeros(w * h * 3 // 2, np.uint8)
    L.orc_nv12_to_i420(ptr(y), ptr(c), 0, w, h, w, w, 1, 0, ptr(exp))
    assert np.array_equal(got[0], exp)
