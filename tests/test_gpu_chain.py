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
