import os
import sys

import pytest

# pinned oracle configuration: one OpenMP thread, so the oracle's reductions sum in a fixed order
os.environ.setdefault("OMP_NUM_THREADS", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One libs3d_cuda context for the whole GPU session.  Raises (never skips, never falls back)
    when the library or the GPU is missing: a GPU test that cannot reach the CUDA path must fail."""
    from struct3d_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
