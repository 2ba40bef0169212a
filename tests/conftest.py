import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # pytest-timeout registers this itself when it is installed; harmless (and silent) otherwise
    config.addinivalue_line("markers", "timeout(seconds): fail instead of hanging (pytest-timeout)")


@pytest.fixture(scope="session")
def bfsm():
    import bfsm_b200
    return bfsm_b200


@pytest.fixture(scope="session")
def port_oracle():
    from oracle import oracle as O
    return O.PortOracle()
