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


# ------------------------------------------------------------------------------------------------ scaler (MSPixConv / MSSizeConv arithmetic)
from pathlib import Path

from mediastreamer2_b200 import _lib

GOLD = Path(__file__).resolve().parent / "golden" / "swscale_bilinear.npz"


def _rand_frames(fmt, w, h, n, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    out = []
    for t in range(n):
        Y = ((xx + 2 * yy + 3 * t) % 256 + rng.integers(-3, 4, size=(h, w))).clip(0, 255).astype(np.uint8)
        ch, cw = h // 2, w // 2
        cy, cx = np.mgrid[0:ch, 0:cw]
        U = ((cx + t) % 256 + rng.integers(-3, 4, size=(ch, cw))).clip(0, 255).astype(np.uint8)
        V = rng.integers(0, 256, size=(ch, cw), dtype=np.uint8) if t % 2 else ((cy + 2 * t) % 256).astype(np.uint8)
        if fmt == _lib.PIX_YUV420P:
            out.append(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]))
        else:
            a, b = (U, V) if fmt == _lib.PIX_NV12 else (V, U)
            out.append(np.concatenate([Y.ravel(), np.stack([a, b], axis=-1).ravel()]))
    return np.stack(out)


@pytest.mark.parametrize("sf,sw,sh,df,dw,dh", [
    (_lib.PIX_NV12, 192, 108, _lib.PIX_RGB24, 128, 72),        # cfg4 shape at 1/10 scale
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_RGB24, 1280, 720),    # cfg4 full size
    (_lib.PIX_NV12, 160, 120, _lib.PIX_RGB24, 64, 48),         # 2.5x down
    (_lib.PIX_NV12, 64, 48, _lib.PIX_RGB24, 96, 72),           # 1.5x up (yuv2rgb_2 path)
    (_lib.PIX_NV21, 96, 64, _lib.PIX_RGB24_REV, 64, 48),       # NV21 -> BGR24
    (_lib.PIX_NV12, 128, 72, _lib.PIX_RGB24, 128, 72),         # same size (yuv2rgb_1 path)
    (_lib.PIX_YUV420P, 192, 108, _lib.PIX_RGB24, 128, 72),     # planar source
    (_lib.PIX_YUV420P, 192, 108, _lib.PIX_YUV420P, 128, 72),   # MSSizeConv: I420 -> I420 down
    (_lib.PIX_YUV420P, 64, 48, _lib.PIX_YUV420P, 160, 120),    # MSSizeConv: I420 -> I420 up
    (_lib.PIX_NV12, 320, 240, _lib.PIX_YUV420P, 320, 240),     # MSPixConv-like: NV12 -> I420 same size through the scaler
    (_lib.PIX_NV12, 208, 112, _lib.PIX_RGB24, 150, 90),        # ragged tiles (150 % 128, 90 % 16 != 0)
    (_lib.PIX_NV12, 192, 108, _lib.PIX_YUV420P, 128, 72),      # NV12 -> I420 down: CbCr plane read in place (pair map)
    (_lib.PIX_NV21, 192, 108, _lib.PIX_YUV420P, 128, 72),      # ... and with Cr first
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_YUV420P, 1280, 720),  # cfg4's reference-shaped two-step as one pass, full size
    (_lib.PIX_NV21, 64, 48, _lib.PIX_YUV420P, 160, 120),       # 2.5x up from an interleaved chroma plane
    (_lib.PIX_NV12, 208, 112, _lib.PIX_YUV420P, 150, 90),      # ragged tiles, odd chroma width (75)
    (_lib.PIX_NV12, 1280, 720, _lib.PIX_YUV420P, 640, 360),    # 2:1 down
])
def test_scaler_bit_exact_vs_oracle(ctx, sf, sw, sh, df, dw, dh):
    L = O.oracle()
    n = 3 if sw < 1000 else 2
    frames = _rand_frames(sf, sw, sh, n, seed=sw * 7 + dh)
    sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
    got = sc.process(frames)
    o = L.orc_scaler_new(sw, sh, sf, dw, dh, df)
    assert L.orc_scaler_src_bytes(o) == sc.src_bytes and L.orc_scaler_dst_bytes(o) == sc.dst_bytes
    for i in range(n):
        exp = np.zeros(sc.dst_bytes + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        assert np.array_equal(got[i], exp[:-64]), (i, np.abs(got[i].astype(int) - exp[:-64].astype(int)).max())
    L.orc_scaler_free(o)
    sc.close()


@pytest.mark.parametrize("sf,sw,sh,df,dw,dh,tile", [
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_RGB24, 640, 360, "<64,2>"),       # 3:1 preview: 8 horizontal taps, 6 / 4 vertical
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_YUV420P, 960, 540, "<64,1>"),     # 2:1: 4-tap filters, windows too wide for 128-column tiles
    (_lib.PIX_YUV420P, 1280, 720, _lib.PIX_YUV420P, 320, 180, "<32,2>"),   # 4:1 thumbnail (MSSizeConv), planar chroma boxes
    (_lib.PIX_NV21, 1920, 1080, _lib.PIX_RGB24_REV, 160, 90, "<16,0>"),    # 12:1: 24 taps each way, Cr first, BGR
    (_lib.PIX_YUV420P, 640, 480, _lib.PIX_RGB24, 128, 96, "<32,0>"),       # 5:1: 12 taps (no register specialisation)
    (_lib.PIX_NV12, 320, 240, _lib.PIX_RGB24, 100, 74, "<64,2>"),          # ragged tiles, rows that are not 16-byte multiples
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_YUV420P, 480, 270, "<32,2>"),     # 4:1, interleaved chroma, last tile row of 14
    (_lib.PIX_YUV420P, 1920, 1088, _lib.PIX_YUV420P, 704, 400, "<64,2>"),  # 2.7:1 planar
])
def test_scaler_down_tiles_bit_exact(ctx, sf, sw, sh, df, dw, dh, tile):
    """down-scales by 2x and more (scale_down_kernel: narrow tiles, both passes through shared memory) == oracle == the
    tile-free kernel they used to run (path 5)"""
    L = O.oracle()
    n = 2
    frames = _rand_frames(sf, sw, sh, n, seed=sw + dh)
    frames[1] = np.random.default_rng(dw).integers(0, 256, size=frames.shape[1], dtype=np.uint8)  # full-range noise
    sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
    assert sc.path == 6, "geometry should select the down-scale tile kernel"
    got = sc.process(frames)
    o = L.orc_scaler_new(sw, sh, sf, dw, dh, df)
    for i in range(n):
        exp = np.zeros(sc.dst_bytes + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        bad = np.flatnonzero(got[i] != exp[:-64])
        assert bad.size == 0, (i, bad.size, bad[:8])
    L.orc_scaler_free(o)
    sc.set_path(5)
    assert sc.path == 5
    assert np.array_equal(sc.process(frames), got)
    sc.close()


@pytest.mark.parametrize("sf,sw,sh,df,dw,dh", [
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_RGB24, 1280, 720),     # cfg4: <4,2> taps, strips of 15 rows
    (_lib.PIX_NV12, 192, 108, _lib.PIX_RGB24, 128, 72),         # one tile column, strips of 9 rows
    (_lib.PIX_NV21, 384, 216, _lib.PIX_RGB24_REV, 256, 136),    # NV21 -> BGR24, ragged last tile row (136 = 2*64 + 8)
    (_lib.PIX_NV12, 256, 144, _lib.PIX_RGB24, 384, 216),        # 1.5x up: <2,2> taps (yuv2rgb_2 rounding)
    (_lib.PIX_NV21, 256, 144, _lib.PIX_RGB24, 256, 144),        # same size: <1,2> taps (yuv2rgb_1 rounding)
    (_lib.PIX_NV12, 640, 368, _lib.PIX_RGB24_REV, 512, 288),    # 1.25x down
])
def test_scaler_strip_kernel_bit_exact(ctx, sf, sw, sh, df, dw, dh):
    """the register-window strip kernel (default) == oracle == its streaming variant (4) == the tile kernels (2, 1)"""
    L = O.oracle()
    n = 2
    frames = _rand_frames(sf, sw, sh, n, seed=sw + dh)
    frames[1] = np.random.default_rng(dw).integers(0, 256, size=frames.shape[1], dtype=np.uint8)  # full-range noise
    sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
    assert sc.path == 3, "geometry should select the strip kernel"
    got = sc.process(frames)
    o = L.orc_scaler_new(sw, sh, sf, dw, dh, df)
    for i in range(n):
        exp = np.zeros(sc.dst_bytes + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        bad = np.flatnonzero(got[i] != exp[:-64])
        assert bad.size == 0, (i, bad.size, bad[:8] // 3 % dw, bad[:8] // 3 // dw)
    L.orc_scaler_free(o)
    for path, kind in ((4, 4), (1, 2), (2, 1)):
        sc.set_path(path)
        assert sc.path == kind
        assert np.array_equal(sc.process(frames), got)
    sc.close()


@pytest.mark.parametrize("sf,sw,sh,df,dw,dh,sched", [
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_RGB24, 1280, 720, 0),      # cfg4: 3:2, 88 of 90 strips on the schedule
    (_lib.PIX_NV21, 384, 216, _lib.PIX_RGB24_REV, 256, 144, 0),     # 3:2, NV21 -> BGR24
    (_lib.PIX_NV12, 1280, 720, _lib.PIX_RGB24, 1280, 720, 1),       # 1:1 (MSPixConv): <1,2> taps
    (_lib.PIX_NV21, 256, 144, _lib.PIX_RGB24_REV, 256, 144, 1),
    (_lib.PIX_NV12, 640, 368, _lib.PIX_RGB24, 512, 288, -1),        # 5:4: no instantiated schedule -> general loop only
])
def test_scaler_static_schedule_bit_exact(ctx, monkeypatch, sf, sw, sh, df, dw, dh, sched):
    """the straight-line (static row schedule) instantiation of the strip kernel == oracle == the general loop"""
    L = O.oracle()
    n = 3
    frames = _rand_frames(sf, sw, sh, n, seed=sw + dh)
    frames[1] = np.random.default_rng(dw).integers(0, 256, size=frames.shape[1], dtype=np.uint8)  # full-range noise
    sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
    k, regular, strips = sc.schedule
    assert sc.path == 3 and k == sched, (sc.path, k, regular, strips)
    if sched >= 0:
        assert regular >= strips - 3 and regular * 4 >= strips * 3, (regular, strips)
    got = sc.process(frames)
    sc.close()
    o = L.orc_scaler_new(sw, sh, sf, dw, dh, df)
    for i in range(n):
        exp = np.zeros(got.shape[1] + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        bad = np.flatnonzero(got[i] != exp[:-64])
        assert bad.size == 0, (i, bad.size, bad[:8] // 3 % dw, bad[:8] // 3 // dw)
    L.orc_scaler_free(o)
    monkeypatch.setenv("MSB200_SCALER_NO_SCHED", "1")
    sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
    assert sc.schedule[0] == -1
    assert np.array_equal(sc.process(frames), got)
    sc.close()


@pytest.mark.parametrize("sf,sw,sh,df,dw,dh", [
    (_lib.PIX_NV12, 1280, 720, _lib.PIX_RGB24, 640, 360),          # 2:1 (source window per tile > one TMA box)
    (_lib.PIX_NV21, 640, 480, _lib.PIX_RGB24_REV, 160, 120),       # 4:1, NV21 -> BGR24
    (_lib.PIX_YUV420P, 1280, 720, _lib.PIX_YUV420P, 320, 180),     # MSSizeConv thumbnail, 4:1 planar
    (_lib.PIX_YUV420P, 352, 288, _lib.PIX_YUV420P, 176, 144),      # CIF -> QCIF
    (_lib.PIX_NV12, 130, 74, _lib.PIX_RGB24, 86, 50),              # widths the TMA maps cannot take (pitch % 16 != 0)
    (_lib.PIX_YUV420P, 100, 60, _lib.PIX_YUV420P, 150, 90),        # odd-pitch planar up-scale
    (_lib.PIX_NV12, 1920, 1080, _lib.PIX_YUV420P, 640, 360),       # 3:1 NV12 -> I420
])
def test_scaler_direct_kernel_bit_exact(ctx, sf, sw, sh, df, dw, dh):
    """the tile-free direct kernel (row pitches that are not multiples of 16; >= 2x down-scales when asked for with
    set_path(5)): same swscale arithmetic, bit-exact vs the oracle"""
    L = O.oracle()
    n = 2
    frames = _rand_frames(sf, sw, sh, n, seed=sw + dh)
    frames[1] = np.random.default_rng(dw).integers(0, 256, size=frames.shape[1], dtype=np.uint8)
    sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
    if sc.path == 6:  # a >= 2x down-scale with TMA-able pitches: the down-scale tile kernel by default, this one on request
        sc.set_path(5)
    assert sc.path == 5
    got = sc.process(frames)
    sc.close()
    o = L.orc_scaler_new(sw, sh, sf, dw, dh, df)
    assert o
    for i in range(n):
        exp = np.zeros(got.shape[1] + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        bad = np.flatnonzero(got[i] != exp[:-64])
        assert bad.size == 0, (i, bad.size, bad[:8])
    L.orc_scaler_free(o)


def test_scaler_full_size_cfg4_matches_real_libswscale_digest(ctx):
    """NV12 1080p -> RGB24 720p: SHA-256 of the GPU output == SHA-256 of the real libswscale 9.1.100 output."""
    import hashlib

    from test_oracle_video import cfg4_frame

    g = np.load(GOLD)
    sc = F.Scaler(ctx, 1920, 1080, _lib.PIX_NV12, 1280, 720, _lib.PIX_RGB24)
    out = sc.process(cfg4_frame()[None, :])
    sc.close()
    assert hashlib.sha256(out[0].tobytes()).digest() == g["cfg4_sha256"].tobytes()


def test_scaler_golden_frames_from_real_libswscale(ctx):
    """the committed small golden frames (real library, default build): RGB24 cases bit-exact."""
    g = np.load(GOLD)
    av2ms = {0: 0, 1: 1, 2: 2, 3: 3, 15: 5, 23: 100, 24: 101, 26: 7, 28: 11}
    checked = 0
    k = 0
    while f"case{k}_src" in g:
        sf, sw, sh, df, dw, dh = [int(v) for v in g[f"case{k}_meta"]]
        k += 1
        sc = F.Scaler(ctx, sw, sh, av2ms[sf], dw, dh, av2ms[df])
        out = sc.process(g[f"case{k - 1}_src"][None, :])
        sc.close()
        assert np.array_equal(out[0], g[f"case{k - 1}_dst_bitexact"])
        if df == 2:
            assert np.array_equal(out[0], g[f"case{k - 1}_dst"])
        checked += 1
    assert checked == k and checked >= 16  # every golden case, the odd-geometry ones through the direct kernel


@pytest.mark.parametrize("fmt,w,h", [(_lib.PIX_RGB24, 96, 64), (_lib.PIX_RGB24_REV, 64, 48), (_lib.PIX_RGB24, 132, 70),
                                     (_lib.PIX_RGB24, 1280, 720), (_lib.PIX_RGB24_REV, 1920, 1080),
                                     (_lib.PIX_RGBA32, 96, 64), (_lib.PIX_RGBA32_REV, 132, 70), (_lib.PIX_RGBA32, 1280, 720),
                                     (_lib.PIX_RGB565, 96, 64), (_lib.PIX_RGB565, 132, 70), (_lib.PIX_RGB565, 1920, 1080)])
def test_pixconv_rgb24_to_i420_bit_exact(ctx, fmt, w, h):
    """MSPixConv's MS_RGB24 / MS_RGB24_REV inputs: GPU == oracle (itself bit-exact vs the real libswscale's C paths,
    tests/test_oracle_video.py); full-range noise, saturated primaries and a flat frame among the inputs."""
    L = O.oracle()
    rng = np.random.default_rng(w * 3 + h)
    bpp = 4 if fmt in (_lib.PIX_RGBA32, _lib.PIX_RGBA32_REV) else (2 if fmt == _lib.PIX_RGB565 else 3)
    frames = rng.integers(0, 256, size=(4, w * h * bpp), dtype=np.uint8)
    prim = {3: [255, 0, 0, 0, 255, 0, 0, 0, 255, 255, 255, 255], 4: [255, 0, 0, 7, 0, 255, 0, 9, 0, 0, 255, 1, 255, 255, 255, 0],
            2: [0x00, 0xF8, 0xE0, 0x07, 0x1F, 0x00, 0xFF, 0xFF]}[bpp]  # RGB565 little endian: red, green, blue, white
    frames[1] = np.tile(np.array(prim, np.uint8), w * h // 4)
    frames[2] = 0
    sc = F.Scaler(ctx, w, h, fmt, w, h, _lib.PIX_YUV420P)
    got = {0: sc.process(frames)}
    # ... and as a plain SWS_BILINEAR call returns them on x86 (the library's SIMD vertical scaler on the chroma rows, folded
    # taps on the top row): the oracle's x86 mode is pinned against the live plain-flag library (test_oracle_video_live.py)
    _lib.check(ctx.lib.msb200_scaler_set_x86_vertical(sc.h, 1))
    got[1] = sc.process(frames)
    sc.close()
    o = L.orc_scaler_new(w, h, fmt, w, h, _lib.PIX_YUV420P)
    assert o
    for mode in (0, 1):
        L.orc_scaler_set_x86_vertical(o, mode)
        for i in range(frames.shape[0]):
            exp = np.zeros(w * h * 3 // 2 + 64, np.uint8)
            L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
            assert np.array_equal(got[mode][i], exp[:-64]), (mode, i, int(np.abs(got[mode][i].astype(int) - exp[:-64].astype(int)).max()))
    L.orc_scaler_free(o)
    d = np.abs(got[0].astype(np.int16) - got[1].astype(np.int16))
    assert d.max() <= 1 and np.array_equal(got[0][:, :w * h], got[1][:, :w * h])  # luma untouched, chroma within one
    if fmt != _lib.PIX_RGB24_REV:
        assert d.max() == 1  # the two roundings really differ (BGR24's special converter has no vertical filter)


@pytest.mark.parametrize("fmt,w,h", [(_lib.PIX_YUYV, 96, 64), (_lib.PIX_UYVY, 64, 48), (_lib.PIX_YUY2, 1280, 720),
                                     (_lib.PIX_YUYV, 104, 32), (_lib.PIX_UYVY, 72, 16)])  # w % 16 == 8: the truncating tail group
def test_pixconv_packed422_to_i420_bit_exact(ctx, fmt, w, h):
    """MSPixConv's packed inputs: GPU == oracle (itself bit-exact vs real libswscale, tests/test_oracle_video.py)."""
    L = O.oracle()
    rng = np.random.default_rng(w + h)
    frames = rng.integers(0, 256, size=(3, w * h * 2), dtype=np.uint8)
    sc = F.Scaler(ctx, w, h, fmt, w, h, _lib.PIX_YUV420P)
    got = sc.process(frames)
    o = L.orc_scaler_new(w, h, fmt, w, h, _lib.PIX_YUV420P)
    for i in range(3):
        exp = np.zeros(w * h * 3 // 2 + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        assert np.array_equal(got[i], exp[:-64]), i
    L.orc_scaler_free(o)
    sc.close()
    g = np.load(GOLD)
    for k in range(20):
        if f"case{k}_meta" not in g:
            break
        sf, sw, sh, df, dw, dh = [int(v) for v in g[f"case{k}_meta"]]
        if (sf == 1 and fmt == _lib.PIX_YUYV) or (sf == 15 and fmt == _lib.PIX_UYVY):
            sc2 = F.Scaler(ctx, sw, sh, fmt, dw, dh, _lib.PIX_YUV420P)
            assert np.array_equal(sc2.process(g[f"case{k}_src"][None, :])[0], g[f"case{k}_dst"])  # the real library's output
            sc2.close()


def test_scaler_host_batches_flow_through_the_chunk_pipeline(ctx):
    """msb200_scaler_process with a batch large enough to be cut into >= 3 chunks (upload / kernels / download overlapped
    on three streams, ragged last chunk): every frame equals the frame converted alone"""
    sw, sh, dw, dh = 1920, 1080, 1280, 720
    n = 17  # 16 MB chunks of 5 frames: 5 + 5 + 5 + 2
    base = _rand_frames(_lib.PIX_NV12, sw, sh, 3, seed=99)
    rng = np.random.default_rng(5)
    frames = np.stack([np.roll(base[i % 3], int(rng.integers(0, 4096))) for i in range(n)])
    sc = F.Scaler(ctx, sw, sh, _lib.PIX_NV12, dw, dh, _lib.PIX_RGB24)
    src_pin = ctx.pinned(frames.shape, np.uint8)
    src_pin[...] = frames
    dst_pin = ctx.pinned((n, sc.dst_bytes), np.uint8)
    got = sc.process(src_pin, dst_pin)
    for i in range(n):
        alone = sc.process(frames[i:i + 1])
        assert np.array_equal(got[i], alone[0]), i
    sc.close()


@pytest.mark.parametrize("sw,sh,dw,dh", [
    (1920, 1080, 1280, 720),   # MSSizeConv's cfg4 step: <4> luma taps, <4> chroma taps (540 -> 360 rows)
    (1280, 720, 1920, 1080),   # up-scale: <2> taps
    (640, 480, 640, 480),      # same size: <1> tap (yuv2plane1)
    (704, 576, 200, 120),      # not reachable by the tiles (3.5x): direct kernel instead
    (384, 288, 264, 200),      # ragged last tile column (264 = 2 * 128 + 8) and tile row
    (192, 108, 128, 72),
])
def test_sizeconv_i420_plane_strips_bit_exact(ctx, sw, sh, dw, dh):
    """MSSizeConv (I420 -> I420 bilinear, sizeconv.c:159-169): the plane-strip kernel == oracle == the tile kernels"""
    L = O.oracle()
    n = 3
    frames = _rand_frames(_lib.PIX_YUV420P, sw, sh, n, seed=sw + dh)
    frames[1] = np.random.default_rng(dh).integers(0, 256, size=frames.shape[1], dtype=np.uint8)
    sc = F.Scaler(ctx, sw, sh, _lib.PIX_YUV420P, dw, dh, _lib.PIX_YUV420P)
    got = sc.process(frames)
    o = L.orc_scaler_new(sw, sh, _lib.PIX_YUV420P, dw, dh, _lib.PIX_YUV420P)
    for i in range(n):
        exp = np.zeros(got.shape[1] + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        bad = np.flatnonzero(got[i] != exp[:-64])
        assert bad.size == 0, (i, bad.size, bad[:8])
    L.orc_scaler_free(o)
    if sc.path != 5:
        sc.set_path(2)  # the generic tile kernels (scale_plane_kernel)
        assert np.array_equal(sc.process(frames), got)
    sc.close()


@pytest.mark.parametrize("sf,sw,sh,dw,dh", [
    (_lib.PIX_NV12, 1920, 1080, 1280, 720),   # the reference-shaped two-step of cfg4 in one call
    (_lib.PIX_NV21, 640, 480, 320, 240 + 8),  # NV21, 2:1-ish (chroma filter of 4 taps)
    (_lib.PIX_NV12, 640, 352, 640, 352),      # same size: plain de-interleave + yuv2plane1
    (_lib.PIX_NV12, 320, 240, 480, 360),      # up-scale
])
def test_nv12_to_i420_with_scaling_bit_exact(ctx, sf, sw, sh, dw, dh):
    """NV12 / NV21 -> I420 with bilinear scaling (exact de-interleave pre-pass + the plane-strip kernels) == oracle"""
    L = O.oracle()
    n = 2
    frames = _rand_frames(sf, sw, sh, n, seed=sw + dh)
    frames[1] = np.random.default_rng(dh).integers(0, 256, size=frames.shape[1], dtype=np.uint8)
    sc = F.Scaler(ctx, sw, sh, sf, dw, dh, _lib.PIX_YUV420P)
    got = sc.process(frames)
    got2 = sc.process(frames[::-1].copy())  # a second batch through the cached tensor maps / scratch buffer
    sc.close()
    o = L.orc_scaler_new(sw, sh, sf, dw, dh, _lib.PIX_YUV420P)
    assert o
    for i in range(n):
        exp = np.zeros(got.shape[1] + 64, np.uint8)
        L.orc_scaler_process(o, ptr(np.ascontiguousarray(frames[i])), ptr(exp))
        assert np.array_equal(got[i], exp[:-64]), i
        assert np.array_equal(got2[n - 1 - i], exp[:-64]), i
    L.orc_scaler_free(o)
