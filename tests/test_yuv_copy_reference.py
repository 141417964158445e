"""CPU: the reference's eight ms_yuv_buf_copy_with_pix_strides patterns run through the UNMODIFIED function (oracle/_ref):
pins tests/yuv_copy_cases.py, the material the GPU kernel is checked with (tests/test_gpu_video_copy.py)."""
import numpy as np
import pytest

import _oracle as O
import yuv_copy_cases as Y


@pytest.mark.parametrize("src_semi,dst_semi,sliding", Y.CASES)
def test_reference_function_on_its_own_patterns(src_semi, dst_semi, sliding):
    R = O.ref()
    bw, bh, roi1, roi2, src, expected = Y.case_buffers(Y.VGA, src_semi, dst_semi, sliding)
    dst = np.zeros_like(src)
    Y.reference_copy(R, src, Y.layout(bw, bh, src_semi), roi1, dst, Y.layout(bw, bh, dst_semi), roi2)
    assert np.array_equal(dst, expected)  # check_picture(): every pixel of the region carries its plane code and index
