import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu through gpurun)")


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
