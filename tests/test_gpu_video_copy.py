"""GPU: msb200_yuv_copy_strided (csrc/video_copy.cu) against the reference's own eight patterns
(tester/mediastreamer2_framework_tester.c:393-500) and, on random geometries, against the UNMODIFIED
ms_yuv_buf_copy_with_pix_strides — every byte of the destination buffer, inside and outside the region."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
import yuv_copy_cases as Y
from mediastreamer2_b200 import _lib

pytestmark = pytest.mark.gpu


def gpu_copy(ctx, n_frames, src, sl, sroi, dst, dl, droi, frame_bytes):
    def lay(l):
        return _lib.YuvLayout((C.c_size_t * 3)(*l[0]), (C.c_int32 * 3)(*l[1]), (C.c_int32 * 3)(*l[2]), frame_bytes)

    a, b = lay(sl), lay(dl)
    _lib.check(ctx.lib.msb200_yuv_copy_strided(ctx.h, n_frames, O.ptr(src), C.byref(a), _lib.Rect(*sroi), O.ptr(dst), C.byref(b),
                                               _lib.Rect(*droi)))


@pytest.mark.parametrize("src_semi,dst_semi,sliding", Y.CASES)
def test_the_references_eight_patterns(ctx, src_semi, dst_semi, sliding):
    bw, bh, roi1, roi2, src, expected = Y.case_buffers(Y.VGA, src_semi, dst_semi, sliding)
    n = 3  # a batch: the same picture three times, frames back to back
    srcs = np.tile(src, n)
    dst = np.zeros_like(srcs)
    gpu_copy(ctx, n, srcs, Y.layout(bw, bh, src_semi), roi1, dst, Y.layout(bw, bh, dst_semi), roi2, src.size)
    for k in range(n):
        assert np.array_equal(dst[k * src.size:(k + 1) * src.size], expected), k


def test_random_regions_equal_the_unmodified_reference_function(ctx):
    R = O.ref()
    rng = np.random.default_rng(42)
    for case in range(40):
        bw, bh = int(rng.integers(8, 60)) * 2, int(rng.integers(8, 40)) * 2
        src_semi, dst_semi = bool(rng.integers(2)), bool(rng.integers(2))
        w, h = int(rng.integers(1, bw // 2)) * 2, int(rng.integers(1, bh // 2)) * 2
        sx, sy = int(rng.integers(0, (bw - w) // 2 + 1)) * 2, int(rng.integers(0, (bh - h) // 2 + 1)) * 2
        same = case % 4 == 0  # equal rectangles: planar -> planar then takes plane_copy's one-memcpy shortcut
        dx, dy = (sx, sy) if same else (int(rng.integers(0, (bw - w) // 2 + 1)) * 2, int(rng.integers(0, (bh - h) // 2 + 1)) * 2)
        sroi, droi = (sx, sy, w, h), (dx, dy, w, h)
        size = bw * bh * 3 // 2
        src = rng.integers(0, 256, size).astype(np.uint8)
        before = rng.integers(0, 256, size).astype(np.uint8)
        exp, got = before.copy(), before.copy()
        sl, dl = Y.layout(bw, bh, src_semi), Y.layout(bw, bh, dst_semi)
        # the shortcut copies row_stride * h bytes from the region's first byte: keep it inside the buffer as a caller must
        if same and not src_semi and not dst_semi and (sy * bw + sx + bw * h > bw * bh):
            continue
        Y.reference_copy(R, src, sl, sroi, exp, dl, droi)
        gpu_copy(ctx, 1, src, sl, sroi, got, dl, droi, size)
        assert np.array_equal(got, exp), (case, bw, bh, src_semi, dst_semi, sroi, droi)


@pytest.mark.parametrize("sw,sh", [(96, 64), (256, 192), (320, 144)])  # 1.5:1 (plane strips), 4:1 and 5:1 x 3:1 (down-scale tiles)
@pytest.mark.parametrize("src_fmt", [_lib.PIX_YUV420P, _lib.PIX_NV12])
def test_mosaic_canvas_equals_oracle_tiles_placed_by_hand(ctx, src_fmt, sw, sh):
    """msb200_scaler_set_canvas: N scaled participants land in their rectangles of one canvas in the scaler's own launches;
    every tile equals the oracle's scaled frame, every byte outside the tiles keeps the background"""
    L = O.oracle()
    lib = ctx.lib
    tw, th = 64, 48
    cw, ch = 144, 100
    tiles = [(0, 0), (72, 0), (0, 50), (80, 52)]
    n_canvas = 3
    n = len(tiles) * n_canvas
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, size=(n, sw * sh * 3 // 2)).astype(np.uint8)
    h = C.c_void_p()
    _lib.check(lib.msb200_scaler_create(ctx.h, sw, sh, src_fmt, tw, th, _lib.PIX_YUV420P, C.byref(h)))
    rects = (_lib.Rect * len(tiles))(*[_lib.Rect(x, y, tw, th) for x, y in tiles])
    _lib.check(lib.msb200_scaler_set_canvas(h, cw, ch, len(tiles), rects))
    cbytes = lib.msb200_scaler_canvas_bytes(h)
    assert cbytes == cw * ch * 3 // 2
    d_src, d_dst = ctx.dev_alloc(src.nbytes), ctx.dev_alloc(cbytes * n_canvas)
    background = np.full(cbytes * n_canvas, 0x5A, np.uint8)
    ctx.h2d(d_src, src)
    ctx.h2d(d_dst, background)
    _lib.check(lib.msb200_scaler_process_dev(h, n, d_src, d_dst))
    got = np.empty_like(background)
    ctx.d2h(got, d_dst)
    # a rectangle that would need unaligned 32-bit stores is refused, not mis-written
    bad = (_lib.Rect * 1)(_lib.Rect(4, 0, tw, th))
    assert lib.msb200_scaler_set_canvas(h, cw, ch, 1, bad) == _lib.EINVAL
    lib.msb200_scaler_destroy(h)
    ctx.dev_free(d_src)
    ctx.dev_free(d_dst)
    o = L.orc_scaler_new(sw, sh, src_fmt, tw, th, _lib.PIX_YUV420P)
    exp = background.copy().reshape(n_canvas, cbytes)
    for k in range(n):
        tile = np.zeros(tw * th * 3 // 2, np.uint8)
        L.orc_scaler_process(o, O.ptr(src[k]), O.ptr(tile))
        c, (x, y) = k // len(tiles), tiles[k % len(tiles)]
        Y_ = exp[c][:cw * ch].reshape(ch, cw)
        U_ = exp[c][cw * ch:cw * ch + cw * ch // 4].reshape(ch // 2, cw // 2)
        V_ = exp[c][cw * ch + cw * ch // 4:].reshape(ch // 2, cw // 2)
        Y_[y:y + th, x:x + tw] = tile[:tw * th].reshape(th, tw)
        U_[y // 2:y // 2 + th // 2, x // 2:x // 2 + tw // 2] = tile[tw * th:tw * th + tw * th // 4].reshape(th // 2, tw // 2)
        V_[y // 2:y // 2 + th // 2, x // 2:x // 2 + tw // 2] = tile[tw * th + tw * th // 4:].reshape(th // 2, tw // 2)
    L.orc_scaler_free(o)
    assert np.array_equal(got.reshape(n_canvas, cbytes), exp)


@pytest.mark.parametrize("sw,sh,dw,dh,src_fmt", [(1920, 1080, 1280, 720, _lib.PIX_YUV420P), (640, 480, 352, 288, _lib.PIX_YUV420P),
                                                 (320, 240, 480, 360, _lib.PIX_NV12), (1280, 720, 320, 180, _lib.PIX_YUV420P),
                                                 (176, 144, 352, 288, _lib.PIX_NV21), (650, 366, 322, 182, _lib.PIX_YUV420P),
                                                 (1920, 1080, 960, 540, _lib.PIX_NV12), (640, 480, 128, 96, _lib.PIX_NV21)])
def test_planar_scaler_x86_vertical_rounding_equals_oracle_x86_mode(ctx, sw, sh, dw, dh, src_fmt):
    """msb200_scaler_set_x86_vertical(1): MSSizeConv's output as a plain SWS_BILINEAR call returns it on x86 — the oracle's
    x86 mode is pinned bit-exact against the live library (tests/test_oracle_video_live.py); the strip kernel (TMA-able
    geometries), the down-scale tile kernel (2:1, 4:1, 5:1) and the tile-free kernel (odd pitch) must all equal it, and differ from the C rounding"""
    L = O.oracle()
    lib = ctx.lib
    rng = np.random.default_rng(sw + dh)
    n = 2
    sbytes = sw * sh * 3 // 2
    src = rng.integers(0, 256, size=(n, sbytes)).astype(np.uint8)
    h = C.c_void_p()
    _lib.check(lib.msb200_scaler_create(ctx.h, sw, sh, src_fmt, dw, dh, _lib.PIX_YUV420P, C.byref(h)))
    dbytes = lib.msb200_scaler_dst_frame_bytes(h)
    outs = {}
    for mode in (0, 1):
        _lib.check(lib.msb200_scaler_set_x86_vertical(h, mode))
        got = np.zeros((n, dbytes), np.uint8)
        _lib.check(lib.msb200_scaler_process(h, n, O.ptr(src), O.ptr(got)))
        outs[mode] = got
    lib.msb200_scaler_destroy(h)
    o = L.orc_scaler_new(sw, sh, src_fmt, dw, dh, _lib.PIX_YUV420P)
    for mode in (0, 1):
        L.orc_scaler_set_x86_vertical(o, mode)
        for k in range(n):
            exp = np.zeros(dbytes, np.uint8)
            L.orc_scaler_process(o, O.ptr(src[k]), O.ptr(exp))
            assert np.array_equal(outs[mode][k], exp), (mode, k)
    L.orc_scaler_free(o)
    if dh != sh:
        d = np.abs(outs[0].astype(np.int16) - outs[1].astype(np.int16))
        assert d.max() == 1  # the two roundings really differ, by one at most
