"""The reference's test material for ms_yuv_buf_copy_with_pix_strides (tester/mediastreamer2_framework_tester.c:67-217,
393-500), restated as numpy: VGA picture inside a buffer padded by 16 columns and 16 rows, pixel value = plane code in the
top two bits | running index mod 32; planar or semi-planar on either side; optionally "sliding" the picture by the padding.
Shared by the CPU test (pins this restatement against the UNMODIFIED reference function) and the GPU test."""
from __future__ import annotations

import ctypes as C

import numpy as np

VGA = (640, 480)
PAD = 16


class MSRect(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int)]


def _plane(code: int, bw: int, bh: int, roi) -> np.ndarray:
    x, y, w, h = roi
    p = np.zeros((bh, bw), np.uint8)
    idx = np.arange(w * h, dtype=np.int64).reshape(h, w)
    p[y:y + h, x:x + w] = (code << 6) | (idx % 32)
    return p


def generate_picture(bw: int, bh: int, roi, semi_planar: bool) -> np.ndarray:
    """generate_picture(): one buffer of bw*bh*3/2 bytes"""
    x, y, w, h = roi
    croi = (x // 2, y // 2, w // 2, h // 2)
    yp = _plane(1, bw, bh, roi)
    up, vp = _plane(2, bw // 2, bh // 2, croi), _plane(3, bw // 2, bh // 2, croi)
    if not semi_planar:
        return np.concatenate([yp.reshape(-1), up.reshape(-1), vp.reshape(-1)])
    inter = np.stack([up, vp], axis=-1)
    return np.concatenate([yp.reshape(-1), inter.reshape(-1)])


def layout(bw: int, bh: int, semi_planar: bool):
    """(plane offsets, row strides, pixel strides) as the reference's test sets them up (:412-454)"""
    n = bw * bh
    if not semi_planar:
        return (0, n, n + n // 4), (bw, bw // 2, bw // 2), (1, 1, 1)
    return (0, n, n + 1), (bw, bw, bw), (1, 2, 2)


CASES = [(s, d, slide) for slide in (False, True) for s in (False, True) for d in (False, True)]  # the reference's eight


def case_buffers(size, src_semi: bool, dst_semi: bool, sliding: bool):
    w, h = size
    bw, bh = w + PAD, h + PAD
    roi1 = (0, 0, w, h)
    roi2 = (PAD, PAD, w, h) if sliding else roi1
    src = generate_picture(bw, bh, roi1, src_semi)
    expected = generate_picture(bw, bh, roi2, dst_semi)
    return bw, bh, roi1, roi2, src, expected


def reference_copy(R, src: np.ndarray, sl, sroi, dst: np.ndarray, dl, droi) -> None:
    """the UNMODIFIED ms_yuv_buf_copy_with_pix_strides (src/voip/msvideo.c:245-270) from oracle/_ref, in place on dst"""
    f = R.ms_yuv_buf_copy_with_pix_strides
    f.restype = None
    f.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), MSRect, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                  C.POINTER(C.c_int), MSRect]
    sp = (C.c_void_p * 3)(*[src.ctypes.data + o for o in sl[0]])
    dp = (C.c_void_p * 3)(*[dst.ctypes.data + o for o in dl[0]])
    f(sp, (C.c_int * 3)(*sl[1]), (C.c_int * 3)(*sl[2]), MSRect(*sroi), dp, (C.c_int * 3)(*dl[1]), (C.c_int * 3)(*dl[2]), MSRect(*droi))
