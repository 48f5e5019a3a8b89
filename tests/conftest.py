import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libb200seed.so / liboracle.so when they are missing (prebuilt files travel to
    the GPU box; nvcc is also available there)."""
    from traccc_b200 import build as b
    b.build()
    from oracle import oracle
    oracle.build()
    yield
