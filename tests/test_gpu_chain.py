"""GPU parity for the resident cfg2 chain (2x MSResample -> MSSpeexEC -> MSVolume [-> MSAudioMixer]) vs the CPU
composition of the oracle pieces, tick by tick, through the host-buffer C-ABI entry point."""
import numpy as np
import pytest

from chain_reference import OracleChain, smoke_chain
from mediastreamer2_b200 import filters as F
from synth import cfg2_stream

pytestmark = pytest.mark.gpu


def test_chain_smoke(ctx):
    assert smoke_chain(ctx) <= 2


@pytest.mark.parametrize("in_rate,rate,pins", [(16000, 48000, 0), (8000, 16000, 0), (16000, 48000, 4)])
def test_chain_matches_oracle_composition(ctx, in_rate, rate, pins):
    n, ticks = 8, 40
    ti = in_rate // 100
    data = [cfg2_stream(100 + s, ti * ticks, in_rate) for s in range(n)]
    ref = np.stack([d[0] for d in data]).reshape(n, ticks, ti)
    mic = np.stack([d[1] for d in data]).reshape(n, ticks, ti)
    ch = F.AudioChain(ctx, n, in_rate, rate, 250, 0.8, pins)
    oc = OracleChain(n, in_rate, rate, 250, 0.8, pins)
    sizes = []
    for t in range(ticks):
        out, k = ch.tick(np.ascontiguousarray(ref[:, t]), np.ascontiguousarray(mic[:, t]))
        exp = oc.tick_all(ref[:, t], mic[:, t])
        assert exp.shape[1] == k, (t, exp.shape, k)
        sizes.append(k)
        if k:
            d = np.abs(out[:, :k].astype(np.int32) - exp.astype(np.int32))
            # conference outputs sum (pins-1) streams: per-stream 2 LSB tolerance accumulates
            assert d.max() <= (2 if not pins else 2 * (pins - 1)), (t, d.max())
    if not pins:
        F_ = oc.F
        assert set(sizes) <= {0, F_, 2 * F_}
        assert abs(sum(sizes) - ticks * (rate // 100)) < F_  # nothing lost, less than one frame still buffered
    ch.close()
    oc.close()


@pytest.mark.parametrize("pins", [0, 4])
def test_pipelined_submit_wait_equals_synchronous_tick(ctx, pins):
    """msb200_chain_submit / _wait (copies overlapped with the neighbouring ticks' kernels, two ticks in flight) gives
    bit-identical samples and block sizes to the synchronous msb200_chain_tick"""
    n, ticks, in_rate, rate = 16, 30, 16000, 48000
    ti = in_rate // 100
    data = [cfg2_stream(300 + s, ti * ticks, in_rate) for s in range(n)]
    ref = np.stack([d[0] for d in data]).reshape(n, ticks, ti)
    mic = np.stack([d[1] for d in data]).reshape(n, ticks, ti)
    a = F.AudioChain(ctx, n, in_rate, rate, 250, 0.8, pins)
    b = F.AudioChain(ctx, n, in_rate, rate, 250, 0.8, pins)
    ref_pin = ctx.pinned((ticks, n, ti), np.int16)
    mic_pin = ctx.pinned((ticks, n, ti), np.int16)
    ref_pin[...] = ref.transpose(1, 0, 2)
    mic_pin[...] = mic.transpose(1, 0, 2)
    outs = [ctx.pinned((n, b.max_out), np.int16) for _ in range(2)]
    sync = [a.tick(np.ascontiguousarray(ref[:, t]), np.ascontiguousarray(mic[:, t])) for t in range(ticks)]
    sync = [(o[:, :k].copy(), k) for o, k in sync]
    got, pending = [], []
    for t in range(ticks):
        if len(pending) == 2:
            b.wait()
            tt, k = pending.pop(0)
            got.append((outs[tt & 1][:, :k].copy(), k))
        k = b.submit(ref_pin[t], mic_pin[t], outs[t & 1])
        pending.append((t, k))
    with pytest.raises(Exception):
        # a third tick in flight is refused (the oldest one must be collected first)
        if len(pending) == 2:
            b.submit(ref_pin[0], mic_pin[0], outs[0])
        else:
            raise RuntimeError("pipeline not full")
    while pending:
        b.wait()
        tt, k = pending.pop(0)
        got.append((outs[tt & 1][:, :k].copy(), k))
    assert [k for _, k in got] == [k for _, k in sync]
    for t, ((o1, k1), (o2, _)) in enumerate(zip(got, sync)):
        assert np.array_equal(o1, o2), t
    a.close()
    b.close()


def test_chain_large_bank_crosses_the_4_gib_state_boundary(ctx):
    """16384 streams: the echo cancellers' device state (4.7 GB) spans the 32-bit offset boundary. Eight distinct inputs
    tiled over the bank: every stream must produce exactly what its twin in the first eight does (which the tests above
    check against the oracle), tick after tick — any 32-bit addressing slip in a kernel or an init pass breaks this."""
    n, base, ticks, in_rate, rate = 16384, 8, 24, 16000, 48000
    ti = in_rate // 100
    data = [cfg2_stream(700 + s, ti * ticks, in_rate) for s in range(base)]
    ref = np.stack([d[0] for d in data]).reshape(base, ticks, ti)
    mic = np.stack([d[1] for d in data]).reshape(base, ticks, ti)
    idx = np.arange(n) % base
    ch = F.AudioChain(ctx, n, in_rate, rate, 250, 0.8, 0)
    produced = 0
    for t in range(ticks):
        out, k = ch.tick(np.ascontiguousarray(ref[idx, t]), np.ascontiguousarray(mic[idx, t]))
        if k:
            produced += k
            got = out[:, :k].reshape(n // base, base, k)
            assert np.array_equal(got, np.broadcast_to(got[0], got.shape)), f"tick {t}"
            assert got[0].any()
    assert produced >= (ticks - 2) * (rate // 100)
    ch.close()


@pytest.mark.parametrize("in_rate,rate,n", [(16000, 48000, 600), (8000, 16000, 64)])
def test_overlap_mode_of_tick_dev_equals_the_serial_order(ctx, in_rate, rate, n):
    """msb200_chain_set_overlap: the resamplers of tick T+1 and the volume / hand-out of tick T-1 on side streams beside
    the canceller of tick T. Same samples as the one-stream order, tick by tick, with a grid that fills the chip (the side
    streams really run beside the canceller) and a rate pair whose rings are as tight as the overlap's run-ahead allows;
    switching the mode on and off in mid-stream included."""
    ticks = 36
    ti = in_rate // 100
    base = [cfg2_stream(700 + s, ti * ticks, in_rate) for s in range(8)]
    rng = np.random.default_rng(2)
    pick, gain = rng.integers(0, 8, n), rng.uniform(0.4, 1.0, (n, 1))
    ref = (np.stack([base[p][0] for p in pick]) * gain).astype(np.int16).reshape(n, ticks, ti)
    mic = (np.stack([base[p][1] for p in pick]) * gain).astype(np.int16).reshape(n, ticks, ti)
    d_ref, d_mic = ctx.dev_alloc(ticks * n * ti * 2), ctx.dev_alloc(ticks * n * ti * 2)
    ctx.h2d(d_ref, np.ascontiguousarray(ref.transpose(1, 0, 2)))
    ctx.h2d(d_mic, np.ascontiguousarray(mic.transpose(1, 0, 2)))
    ctx.sync()
    outs = []
    for overlap in (False, True):
        ch = F.AudioChain(ctx, n, in_rate, rate, 250, 0.8, 0)
        d_out = ctx.dev_alloc(ticks * n * ch.max_out * 2)  # one output block per tick: nothing is overwritten in flight
        got, sizes = np.zeros((ticks, n, ch.max_out), np.int16), []
        for t in range(ticks):
            ch.set_overlap(overlap and not 20 <= t < 24)  # a stretch of serial ticks in the middle of the overlapped run
            sizes.append(ch.tick_dev(d_ref + t * n * ti * 2, d_mic + t * n * ti * 2, d_out + t * n * ch.max_out * 2))
        ch.join()
        ctx.d2h(got, d_out)
        ctx.sync()
        outs.append((got, sizes))
        ch.close()
        ctx.dev_free(d_out)
    ctx.dev_free(d_ref)
    ctx.dev_free(d_mic)
    assert outs[0][1] == outs[1][1] and sum(outs[0][1]) > 0
    for t, k in enumerate(outs[0][1]):
        assert np.array_equal(outs[0][0][t, :, :k], outs[1][0][t, :, :k]), t
    assert np.abs(outs[0][0]).max() > 0
