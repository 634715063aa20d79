import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library. Built on demand in a fresh checkout (the driver builds it through
    __graft_entry__.build() before the tests)."""
    from envutil_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build_library()
    return capi.load()


@pytest.fixture(scope="session")
def engine(lib):
    from envutil_b200.engine import Engine
    eng = Engine(0)  # raises without a GPU: there is no CPU path
    yield eng
    eng.close()
