import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    import orc

    orc.build()
    yield


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes front-end); builds the CUDA library in-tree if needed."""
    import __graft_entry__ as ge

    ge.build()
    return ge.load_package()
