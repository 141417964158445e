"""oracle/oracle_video.c against the LIVE libswscale of this image (the library the reference's ffmpeg scaler back-end calls,
src/voip/msvideo.c:651-681) on random geometries and formats — a wider net than the 16 committed golden frames. Skipped
where the library is absent (the GPU box). Also pins, as a known limit, the one place where the library leaves the algorithm
the oracle restates."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import _oracle as O  # noqa: E402
from _oracle import ptr  # noqa: E402
from bench_video import _load_libswscale  # noqa: E402
from make_swscale_golden import AV_PIX, SWS_BILINEAR, SWS_BITEXACT, sws_convert  # noqa: E402
from make_swscale_golden import test_frame as make_frame  # noqa: E402

AV2MS = {0: 0, 1: 1, 2: 2, 3: 3, 15: 5, 23: 100, 24: 101, 26: 7, 28: 11, 37: 8}  # AVPixelFormat -> MSB200_PIX_*
_SWS = None


def _sws():
    global _SWS
    if _SWS is None:
        _SWS = _load_libswscale() or False
    if not _SWS:
        pytest.skip("libswscale (opencv wheel) not available here")
    return _SWS


def _oracle_convert(src, sf, sw, sh, df, dw, dh):
    L = O.oracle()
    s = L.orc_scaler_new(sw, sh, AV2MS[AV_PIX[sf]], dw, dh, AV2MS[AV_PIX[df]])
    assert s
    out = np.zeros(L.orc_scaler_dst_bytes(s) + 64, np.uint8)
    assert L.orc_scaler_process(s, ptr(np.ascontiguousarray(src)), ptr(out)) == 0
    L.orc_scaler_free(s)
    return out[:-64]


@pytest.mark.parametrize("seed", range(60))
def test_scaler_oracle_vs_live_libswscale_random(seed):
    """random source sizes (16 .. 300), scale factors 0.25 .. 3 (independent per axis), all supported format pairs.
    Bit-exact against the library's C reference functions (SWS_BITEXACT); where that flag itself changes the FILTER (it zeroes
    the taps the alignment padding would keep, utils.c initFilter) the oracle must equal the plain-flag output the reference
    really asks for: exactly for RGB24, within the library's approximate x86 SIMD for planar (<= 1) and BGR24 (<= 5)."""
    sws = _sws()
    rng = np.random.default_rng(seed)
    kind = int(rng.integers(0, 4))
    if kind == 0:
        sf, df = str(rng.choice(["nv12", "nv21", "yuv420p"])), str(rng.choice(["rgb24", "bgr24"]))
    elif kind == 1:
        sf = df = "yuv420p"
    elif kind == 2:
        sf, df = str(rng.choice(["nv12", "nv21"])), "yuv420p"
    else:
        sf, df = str(rng.choice(["yuyv422", "uyvy422", "rgb24", "bgr24", "rgba", "bgra"])), "yuv420p"
    sw, sh = int(rng.integers(8, 150)) * 2, int(rng.integers(8, 100)) * 2
    if kind == 3:
        dw, dh = sw, sh
    else:
        r = float(rng.choice([0.25, 0.4, 0.5, 0.6667, 0.75, 0.9, 1.0, 1.25, 1.5, 2.0, 3.0]))
        r2 = r if rng.integers(0, 3) else float(rng.choice([0.5, 0.75, 1.0, 1.5]))
        dw, dh = max(8, int(sw * r) // 2 * 2), max(8, int(sh * r2) // 2 * 2)
        if sf == "yuv420p" and df in ("rgb24", "bgr24") and (dw, dh) == (sw, sh):
            dw += 2  # see test_yuv420p_to_rgb_same_size_known_limit
    src = make_frame(sf, sw, sh, t=seed, seed=seed)
    out = _oracle_convert(src, sf, sw, sh, df, dw, dh)
    exact = sws_convert(sws, src, sf, sw, sh, df, dw, dh, SWS_BILINEAR | SWS_BITEXACT)
    if np.array_equal(out, exact):
        return
    plain = sws_convert(sws, src, sf, sw, sh, df, dw, dh, SWS_BILINEAR)
    d = np.abs(out.astype(int) - plain.astype(int))
    tol = 0 if df == "rgb24" else (5 if df == "bgr24" else 1)
    assert d.max() <= tol, (sf, sw, sh, df, dw, dh, int(d.max()))


@pytest.mark.parametrize("fmt", ["yuyv422", "uyvy422"])
@pytest.mark.parametrize("w", [46, 66, 90, 104, 148, 186, 260, 96])
def test_packed422_every_width_matches_the_library(fmt, w):
    """YUYV / UYVY -> I420: the library (x86) averages the two chroma lines with a rounding SIMD average over whole groups of
    8 chroma samples and a truncating scalar loop over the rest of the row; the oracle (and the GPU kernel) reproduce exactly
    that, so widths that are not multiples of 16 are bit-exact too"""
    sws = _sws()
    h = 16
    src = make_frame(fmt, w, h, t=1, seed=w)
    out = _oracle_convert(src, fmt, w, h, "yuv420p", w, h)
    assert np.array_equal(out, sws_convert(sws, src, fmt, w, h, "yuv420p", w, h, SWS_BILINEAR))


def test_yuv420p_to_rgb_same_size_known_limit():
    """YUV420P -> RGB24 at IDENTICAL size: the library leaves the scaler for its table-driven unscaled converter (yuv2rgb.c);
    the oracle (and the GPU) stay on the scaler path the reference's display filters get at every other size. NV12 input has
    no such converter and is exact at identical size too."""
    sws = _sws()
    w, h = 128, 72
    src = make_frame("yuv420p", w, h, t=5, seed=5)
    out = _oracle_convert(src, "yuv420p", w, h, "rgb24", w, h)
    ref = sws_convert(sws, src, "yuv420p", w, h, "rgb24", w, h, SWS_BILINEAR)
    d = np.abs(out.astype(int) - ref.astype(int))
    assert 0 < d.max() <= 6 and d.mean() < 1.0
    src = make_frame("nv12", w, h, t=5, seed=5)
    assert np.array_equal(_oracle_convert(src, "nv12", w, h, "rgb24", w, h),
                          sws_convert(sws, src, "nv12", w, h, "rgb24", w, h, SWS_BILINEAR))


@pytest.mark.parametrize("seed", range(40))
def test_scaler_oracle_vs_live_libswscale_extreme_ratios_and_odd_sizes(seed):
    """scale factors 0.1 .. 6 (independent per axis), odd source / destination sizes (odd chroma geometry, last-row and
    last-column folding of the filters): same acceptance as above. Output widths below 16 are left out: the library itself
    writes past the end of such narrow tight planes."""
    sws = _sws()
    rng = np.random.default_rng(10000 + seed)
    sf, df = str(rng.choice(["nv12", "nv21", "yuv420p"])), str(rng.choice(["rgb24", "bgr24", "yuv420p"]))
    odd = bool(rng.integers(0, 2))
    sw, sh = int(rng.integers(16, 400)), int(rng.integers(16, 300))
    r = float(rng.choice([0.1, 0.15, 0.2, 0.3, 0.45, 0.55, 0.8, 1.1, 1.7, 2.5, 4.0, 6.0]))
    r2 = float(rng.choice([r, r, 0.3, 1.0, 2.2]))
    dw, dh = max(16, int(sw * r)), max(8, int(sh * r2))
    if not odd:
        sw, sh, dw, dh = sw & ~1, sh & ~1, dw & ~1, dh & ~1
    if df != "yuv420p":
        dw &= ~1  # odd RGB widths are refused: test_odd_rgb_width_is_refused
    if (dw, dh) == (sw, sh):
        dw += 2
    if dw > 1200 or dh > 900:
        dw, dh = min(dw, 1200) & ~1, min(dh, 900) & ~1
    src = make_frame(sf, sw, sh, t=seed, seed=seed)
    out = _oracle_convert(src, sf, sw, sh, df, dw, dh)
    exact = sws_convert(sws, src, sf, sw, sh, df, dw, dh, SWS_BILINEAR | SWS_BITEXACT)
    if np.array_equal(out, exact):
        return
    plain = sws_convert(sws, src, sf, sw, sh, df, dw, dh, SWS_BILINEAR)
    d = np.abs(out.astype(int) - plain.astype(int))
    tol = 0 if df == "rgb24" else (5 if df == "bgr24" else 1)
    assert d.max() <= tol, (sf, sw, sh, df, dw, dh, int(d.max()))


def test_odd_rgb_width_is_refused():
    """for an odd RGB output width the library forces SWS_FULL_CHR_H_INT (full horizontal chroma interpolation, another
    algorithm): the oracle refuses it, and so does msb200_scaler_create"""
    L = O.oracle()
    assert not L.orc_scaler_new(64, 48, 0, 33, 24, 2)
    s = L.orc_scaler_new(64, 48, 0, 33, 24, 0)  # planar output: fine
    assert s
    L.orc_scaler_free(s)


@pytest.mark.parametrize("seed", range(40))
def test_planar_output_x86_vertical_mode_equals_plain_flag_library_exactly(seed):
    """what the reference's MSSizeConv really gets on x86: sws_getContext(..., SWS_BILINEAR) without SWS_BITEXACT runs the
    library's SIMD vertical scaler for planar output (x86/yuv2yuvX.asm: 16-bit accumulators, one pmulhw per tap, rounder
    (64 + 8 (taps - 1)) >> 4, final >> 3; the last two output lines by the C functions). The oracle's x86 mode restates it:
    bit-exact against the live library on random geometries. The GPU kernels compute the C arithmetic (<= 1 away) this
    round; switching them is the next step for row a8."""
    sws = _sws()
    L = O.oracle()
    rng = np.random.default_rng(seed)
    sf = str(rng.choice(["yuv420p", "nv12", "nv21"]))
    sw, sh = int(rng.integers(8, 200)) * 2, int(rng.integers(8, 120)) * 2
    r = float(rng.choice([0.25, 0.4, 0.5, 0.6667, 0.75, 0.9, 1.25, 1.5, 2.0, 3.0]))
    r2 = r if rng.integers(0, 3) else float(rng.choice([0.5, 0.75, 1.5]))
    dw, dh = max(16, int(sw * r) // 2 * 2), max(8, int(sh * r2) // 2 * 2)
    src = make_frame(sf, sw, sh, t=seed, seed=seed)
    s = L.orc_scaler_new(sw, sh, AV2MS[AV_PIX[sf]], dw, dh, AV2MS[AV_PIX["yuv420p"]])
    assert s
    L.orc_scaler_set_x86_vertical(s, 1)
    out = np.zeros(L.orc_scaler_dst_bytes(s) + 64, np.uint8)
    assert L.orc_scaler_process(s, ptr(np.ascontiguousarray(src)), ptr(out)) == 0
    L.orc_scaler_free(s)
    assert np.array_equal(out[:-64], sws_convert(sws, src, sf, sw, sh, "yuv420p", dw, dh, SWS_BILINEAR))


@pytest.mark.parametrize("seed", range(24))
def test_rgb_sources_x86_vertical_mode_equals_plain_flag_library_exactly(seed):
    """MSPixConv's RGB24 / RGBA / BGRA inputs go through the scaler's 2:1 vertical chroma filter, hence through the same SIMD
    vertical scaler on x86 (border lines carry folded coefficients there, which matters once every tap is truncated);
    BGR24 takes the library's special converter and has no such mode. x86 mode == live plain-flag library, bit for bit."""
    sws = _sws()
    L = O.oracle()
    rng = np.random.default_rng(seed)
    sf = str(rng.choice(["rgb24", "rgba", "bgra", "bgr24"]))
    w, h = int(rng.integers(4, 100)) * 4, int(rng.integers(4, 80)) * 2
    src = make_frame(sf, w, h, t=seed, seed=seed)
    s = L.orc_scaler_new(w, h, AV2MS[AV_PIX[sf]], w, h, AV2MS[AV_PIX["yuv420p"]])
    assert s
    L.orc_scaler_set_x86_vertical(s, 1)
    out = np.zeros(L.orc_scaler_dst_bytes(s) + 64, np.uint8)
    assert L.orc_scaler_process(s, ptr(np.ascontiguousarray(src)), ptr(out)) == 0
    L.orc_scaler_free(s)
    assert np.array_equal(out[:-64], sws_convert(sws, src, sf, w, h, "yuv420p", w, h, SWS_BILINEAR))


@pytest.mark.parametrize("seed", range(16))
@pytest.mark.parametrize("flags_x86", [(SWS_BILINEAR | SWS_BITEXACT, 0), (SWS_BILINEAR, 1)])
def test_rgb565_source_equals_the_library_exactly(seed, flags_x86):
    """MS_RGB565 (AV_PIX_FMT_RGB565LE) -> YUV420P at the same size, MSPixConv's remaining RGB input: the library's 16-bit
    reader is the RGB24 reader on r5 << 3, g6 << 2, b5 << 3; both roundings of the vertical chroma filter"""
    sws = _sws()
    L = O.oracle()
    flags, x86 = flags_x86
    rng = np.random.default_rng(100 + seed)
    w, h = int(rng.integers(4, 100)) * 4, int(rng.integers(4, 80)) * 2
    src = rng.integers(0, 256, w * h * 2).astype(np.uint8)
    if seed % 3 == 0:  # saturated colours and pure ramps as well as noise
        px = (np.arange(w * h, dtype=np.uint32) * 2654435761 >> 7).astype(np.uint16)
        px[: w * h // 3] = 0xFFFF
        px[w * h // 3: w * h // 2] = 0x0000
        src = px.view(np.uint8).copy()
    s = L.orc_scaler_new(w, h, 8, w, h, 0)
    assert s
    L.orc_scaler_set_x86_vertical(s, x86)
    out = np.zeros(L.orc_scaler_dst_bytes(s) + 64, np.uint8)
    assert L.orc_scaler_process(s, ptr(np.ascontiguousarray(src)), ptr(out)) == 0
    L.orc_scaler_free(s)
    assert np.array_equal(out[:-64], sws_convert(sws, src, "rgb565le", w, h, "yuv420p", w, h, flags))
