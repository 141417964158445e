"""MSGenericPLC on the GPU (the concealment half): a bank of streams with different loss patterns,
bit-exact against oracle/oracle_plc.c, which tests/test_oracle_vs_reference.py pins against the unmodified reference
filter (mixed-radix float kiss_fft included)."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
from _oracle import ptr
from mediastreamer2_b200 import filters as F
from test_oracle_vs_reference import plc_signal

pytestmark = pytest.mark.gpu

LOSS = [
    set(range(10, 14)) | {20} | set(range(30, 52)) | {60, 61},  # burst, single, 220 ms hole, pair
    {5},
    set(),                                                      # never loses anything: pure 5 ms delay
    set(range(3, 70)),                                          # hole to the end: fade, silence
    {8, 10, 12, 14, 16, 18},                                    # every other block
    set(range(25, 36)) | set(range(40, 44)),                    # 110 ms: the generated signal is stretched a second time
    set(range(0, 6)) | {30},                                    # starts with a hole (nothing to conceal before a block)
    set(range(15, 19)) | set(range(50, 58)),
]


def _drive(ctx, rate, ticks, loss, cn_at=None):
    """runs the oracle (filter level: its concealer clock decides) and the bank side by side; returns #concealed blocks"""
    L = O.oracle()
    n, streams = rate // 100, len(loss)
    cn_at = cn_at or {}
    x = np.stack([plc_signal(rate, ticks * n, seed=10 + s) for s in range(streams)])
    bank = F.GenericPLC(ctx, streams, rate, n)
    orc = [L.orc_plc_create(rate) for _ in range(streams)]
    assert bank.history_samples == L.orc_plc_history_len(orc[0])
    cng_running = [False] * streams
    concealed, kind = 0, C.c_int(0)
    for k in range(ticks):
        # received blocks
        io = np.zeros((streams, n), np.int16)
        mode = np.zeros(streams, np.uint8)
        want = {}
        for s in range(streams):
            if k in cn_at.get(s, ()):
                L.orc_plc_filter_set_cn(orc[s])
            if k not in loss[s]:
                b = x[s, k * n:(k + 1) * n].copy()
                io[s] = b
                mode[s] = F.GenericPLC.PACKET | (F.GenericPLC.AFTER_CNG if cng_running[s] else 0)
                cng_running[s] = False
                L.orc_plc_filter_packet(orc[s], k * 10, ptr(b), n, 1)
                want[s] = b
        if mode.any():
            got = bank.process(io, mode)
            for s, b in want.items():
                assert np.array_equal(got[s], b), f"tick {k} stream {s}: received block"
        # end of tick: conceal where the oracle's concealer clock says a block is missing
        mode[:] = 0
        want = {}
        for s in range(streams):
            o = np.zeros(n, np.int16)
            m = L.orc_plc_filter_tick(orc[s], k * 10, 10, 1, ptr(o), C.byref(kind))
            if m and kind.value == 1:
                mode[s] = F.GenericPLC.CONCEAL
                want[s] = o
            elif m:
                cng_running[s] = True  # comfort noise: the host emits flagged silence, the bank is not involved
        if mode.any():
            got = bank.process(np.zeros((streams, n), np.int16), mode)
            for s, o in want.items():
                assert np.array_equal(got[s], o), f"tick {k} stream {s}: concealed block"
                concealed += 1
    for c in orc:
        L.orc_plc_destroy(c)
    bank.close()
    return concealed


@pytest.mark.parametrize("rate", [8000, 16000, 32000, 48000])
def test_plc_bank_bit_exact_vs_oracle(ctx, rate):
    """8 streams, 70 ticks: every received (delayed, cross-faded) and every concealed block equals the oracle's; transform
    sizes 400/800 ... 2400/4800 exercise the radix 2, 3, 4 and 5 butterflies"""
    assert _drive(ctx, rate, 70, LOSS) > 100


def test_plc_bank_counters_wrap_like_the_reference(ctx):
    """a 2.2 s hole at 48 kHz: plc_samples_used wraps at 65536 (genericplc.h:46) and stale signal reappears"""
    assert _drive(ctx, 48000, 260, [set(range(20, 240)), set(range(100, 250))]) > 300


def test_plc_bank_after_comfort_noise(ctx):
    """MS_GENERIC_PLC_SET_CN before a hole: no concealment during it, and the first block afterwards fades in from zero"""
    loss = [set(range(12, 20)) | set(range(30, 34)), set(range(12, 20))]
    assert _drive(ctx, 16000, 50, loss, cn_at={0: (12,)}) > 8


def test_plc_rejects_rates_with_large_prime_factors(ctx):
    with pytest.raises(Exception):
        F.GenericPLC(ctx, 4, 44100)  # 2200-point transform: 11 is not a supported radix


def test_plc_reset_stream(ctx):
    rate, n = 16000, 160
    bank = F.GenericPLC(ctx, 2, rate, n)
    x = np.stack([plc_signal(rate, 10 * n, seed=s) for s in range(2)])
    for k in range(10):
        bank.process(x[:, k * n:(k + 1) * n], [1, 1])
    bank.reset_stream(1)
    out = bank.process(np.zeros((2, n), np.int16), [2, 2])
    assert out[0].any() and not out[1].any()  # stream 1 conceals from an empty history
    bank.close()
