import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than a few seconds")


@pytest.fixture(scope="session")
def oracle():
    """The C++ oracle (oracle/c), built on demand.  Checker only."""
    from oracle import c_oracle
    c_oracle.build()
    c_oracle.lib()
    return c_oracle


@pytest.fixture(scope="session")
def gpu():
    """masp_b200.prover bound to cuda:0 through the C ABI (no fallback)."""
    import masp_b200.prover as pv
    pv.init(0)
    return pv


@pytest.fixture(scope="session")
def emu():
    """The device code compiled for the host (tests/emu): exercises launch
    orchestration and index arithmetic on CPU.  Never a product path; it is
    bound to a private copy of the ctypes layer so the package's own library
    handle is untouched."""
    from util import load_emu
    return load_emu()
