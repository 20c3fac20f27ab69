import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "parallel-in-time-ode-filters_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """The built CUDA library (compiled here with nvcc; loading it does not need a GPU)."""
    import __graft_entry__ as g

    g.build()
    from pof import _native

    return _native
