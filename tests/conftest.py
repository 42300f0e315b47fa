import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def gpu():
    """Initialises the CUDA library; fails (does not skip) when it cannot: a GPU test
    that silently falls back to the CPU would void the parity claim."""
    from algoplonk_b200 import _lib
    _lib.init()
    return _lib
