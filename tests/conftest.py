import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def coracle():
    """The C oracle (oracle/zkp_oracle.c), built on demand.  Test-side checker only."""
    import coracle as c
    c.build()
    c.lib()
    return c


@pytest.fixture(scope="session")
def pyref():
    import pyref as o
    return o


@pytest.fixture(scope="session")
def engine():
    """The product: CUDA engine through the C ABI.  Fails (does not skip) when CUDA is missing."""
    import zkvm_pairings_b200 as z
    from zkvm_pairings_b200 import build
    build.build()
    eng = z.PairingEngine([0])
    yield eng
    eng.close()
