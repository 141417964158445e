"""MSAudioFlowControl on the GPU (the flow-control half): bit-exact against oracle/oracle_audio.c,
which tests/test_oracle_vs_reference.py pins against the unmodified reference filter."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
from _oracle import OrcFlowCtl, ptr
from mediastreamer2_b200 import filters as F
from test_oracle_vs_reference import _flowctl_signal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [160, 480, 333])
def test_flowcontrol_bank_bit_exact_vs_oracle(ctx, n):
    """8 streams with different strategies, targets armed at different ticks, silent frames, flat runs (ties), the
    too-many-samples whole-frame drop; every block's samples, length and controller state equal the oracle's"""
    L = O.oracle()
    streams, ticks, rate = 8, 40, 100 * n
    rng = np.random.default_rng(n)
    x = np.stack([_flowctl_signal(np.random.default_rng(100 + s), n, ticks) for s in range(streams)])
    x[5] = (x[5].astype(np.int32) * 5).clip(-32768, 32767)  # loud: three-sample measures above 32768 occur
    cfg = [(1, 30, 200), (1, 8, 100), (0, 40, 300), (1, 120, 150), (1, 15, 400), (1, 25, 250), (0, 10, 50), (1, 60, 120)]
    arm_at = [2, 0, 5, 3, 1, 7, 4, 2]
    fc = F.FlowControl(ctx, streams, n)
    orc = [OrcFlowCtl() for _ in range(streams)]
    for s in range(streams):
        L.orc_flowctl_init(C.byref(orc[s]))
        orc[s].strategy = cfg[s][0]
        fc.set_config(s, cfg[s][0], 0.02)
    dropped_any = 0
    for t in range(ticks):
        for s in range(streams):
            if t == arm_at[s] or (t == arm_at[s] + 25 and s % 2 == 0):  # some streams are armed a second time
                tgt, tot = cfg[s][1] * rate // 1000, cfg[s][2] * rate // 1000
                fc.set_target(s, tgt, tot)
                L.orc_flowctl_set_target(C.byref(orc[s]), tgt, tot)
        blk = np.ascontiguousarray(x[:, t * n:(t + 1) * n])
        got, got_n = fc.process(blk)
        for s in range(streams):
            exp = blk[s].copy()
            k = L.orc_flowctl_process(C.byref(orc[s]), ptr(exp), n)
            assert got_n[s] == k, (t, s, got_n[s], k)
            assert np.array_equal(got[s, :k], exp[:k]), (t, s)
            dropped_any += n - k
            st = fc.state(s)
            assert (st.target_samples, st.total_samples, st.current_pos, st.current_dropped) == \
                   (orc[s].target_samples, orc[s].total_samples, orc[s].current_pos, orc[s].current_dropped), (t, s)
    assert dropped_any > 0
    fc.close()
