"""oracle/oracle_g711.c against the committed exhaustive tables of the UNMODIFIED reference g711.c
(tests/golden/g711_reference.npz, made by tests/golden/make_g711_golden.py): the pin that travels with the repository."""
from pathlib import Path

import numpy as np
import pytest

import _oracle as O
from _oracle import ptr

GOLD = Path(__file__).resolve().parent / "golden" / "g711_reference.npz"


@pytest.mark.parametrize("law,name", [(0, "alaw"), (1, "ulaw")])
def test_g711_oracle_equals_reference_tables(law, name):
    L, g = O.oracle(), np.load(GOLD)
    pcm = np.arange(-32768, 32768, dtype=np.int16)
    code = np.zeros(pcm.size, np.uint8)
    L.orc_g711_encode(law, ptr(pcm), ptr(code), pcm.size)
    assert np.array_equal(code, g[f"{name}_enc"])
    codes = np.arange(256, dtype=np.uint8)
    lin = np.zeros(256, np.int16)
    L.orc_g711_decode(law, ptr(codes), ptr(lin), 256)
    assert np.array_equal(lin, g[f"{name}_dec"])
