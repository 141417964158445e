"""G.711 on the GPU (the decode / encode stubs of BASELINE cfg5 made real): bit-exact against
oracle/oracle_g711.c, which tests/test_oracle_vs_reference.py pins exhaustively against the unmodified g711.c."""
import numpy as np
import pytest

import _oracle as O
from _oracle import ptr
from mediastreamer2_b200 import filters as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = F.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("law", [F.G711_ALAW, F.G711_ULAW])
def test_g711_every_value_bit_exact(ctx, law):
    L = O.oracle()
    pcm = np.arange(-32768, 32768, dtype=np.int16)
    exp_c = np.zeros(pcm.size, np.uint8)
    L.orc_g711_encode(law, ptr(pcm), ptr(exp_c), pcm.size)
    assert np.array_equal(F.g711_encode(ctx, law, pcm), exp_c)
    codes = np.arange(256, dtype=np.uint8)
    exp_p = np.zeros(256, np.int16)
    L.orc_g711_decode(law, ptr(codes), ptr(exp_p), 256)
    assert np.array_equal(F.g711_decode(ctx, law, codes), exp_p)


@pytest.mark.parametrize("law", [F.G711_ALAW, F.G711_ULAW])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 80, 4096 * 80 + 7, 4096 * 160])
def test_g711_batches_ragged_and_full(ctx, law, n):
    """4096 streams x 10 / 20 ms of 8 kHz payload as one flat batch, plus ragged sizes (vector body + scalar tail)"""
    L = O.oracle()
    rng = np.random.default_rng(n + law)
    pcm = rng.integers(-32768, 32768, n).astype(np.int16)
    code = F.g711_encode(ctx, law, pcm)
    exp_c = np.zeros(n, np.uint8)
    L.orc_g711_encode(law, ptr(pcm), ptr(exp_c), n)
    assert np.array_equal(code, exp_c)
    dec = F.g711_decode(ctx, law, code)
    exp_p = np.zeros(n, np.int16)
    L.orc_g711_decode(law, ptr(exp_c), ptr(exp_p), n)
    assert np.array_equal(dec, exp_p)
    # size-independent property: decode(encode(.)) is idempotent under a second round trip
    assert np.array_equal(F.g711_encode(ctx, law, dec), code)


def test_g711_unaligned_device_buffers_take_the_scalar_path(ctx):
    L = O.oracle()
    n = 1000
    rng = np.random.default_rng(3)
    code = rng.integers(0, 256, n + 16).astype(np.uint8)
    d_code, d_pcm = ctx.dev_alloc(n + 64), ctx.dev_alloc(2 * n + 64)
    ctx.h2d(d_code, code)
    F.g711_decode_dev(ctx, F.G711_ULAW, d_code + 1, d_pcm + 2, n)  # misaligned on purpose
    got = np.zeros(n + 1, np.int16)
    ctx.d2h(got, d_pcm)
    exp = np.zeros(n, np.int16)
    L.orc_g711_decode(1, ptr(np.ascontiguousarray(code[1:n + 1])), ptr(exp), n)
    assert np.array_equal(got[1:], exp)
