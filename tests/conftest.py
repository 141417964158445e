import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One msb200 context for the whole GPU session. Fails loudly (no skip) if the library cannot run."""
    from mediastreamer2_b200.filters import Context

    c = Context(0)
    yield c
    c.close()
