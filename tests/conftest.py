import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: about a minute of CPU oracle time (still part of -m gpu; deselect with -m 'gpu and not slow')")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib

    oracle_lib.build()
    return oracle_lib
