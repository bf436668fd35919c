import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine_lib():
    """The CUDA engine library; GPU tests fail loudly (never fall back) when it is missing."""
    from skirt9_b200 import abi
    return abi.load_engine_library()
