"""oracle/oracle_plc.c against the committed outputs of the UNMODIFIED reference MSGenericPLC (tests/golden/plc_reference.npz,
made by tests/golden/make_plc_golden.py from oracle/_ref): the pin that travels with the repository."""
from pathlib import Path

import numpy as np
import pytest

from golden.make_plc_golden import CASES
from test_oracle_vs_reference import plc_oracle_run, plc_schedule

GOLD = Path(__file__).resolve().parent / "golden" / "plc_reference.npz"


@pytest.mark.parametrize("name", sorted(CASES))
def test_plc_oracle_equals_reference_golden(name):
    rate, ticks, lost, block_ms, cn_at, _seed = CASES[name]
    g = np.load(GOLD)
    out, blocks = plc_oracle_run(rate, ticks, plc_schedule(rate, ticks, set(lost), block_ms), g[f"{name}_in"], cn_at)
    assert [b for _, b in blocks] == list(g[f"{name}_sizes"])
    assert np.array_equal(out, g[f"{name}_out"])
    assert np.abs(out.astype(int)).max() > 1000
