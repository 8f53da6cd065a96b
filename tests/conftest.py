import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu through gpurun)")


def pytest_sessionstart(session):
    """The built libraries are git-ignored: in a fresh checkout build them first (nvcc cross-compiles
    sm_100a without a GPU; the oracle is plain gcc).  A no-op when they are already there."""
    lib = os.path.join(ROOT, "gpuacceleratedtracking_b200", "libgat.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def gat():
    import gpuacceleratedtracking_b200 as g
    return g


@pytest.fixture(scope="session")
def engine(gat):
    eng = gat.Engine(0)
    yield eng
    eng.close()
