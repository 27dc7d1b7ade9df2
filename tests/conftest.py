import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The product library; building it needs nvcc but no GPU."""
    from bmfr_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def limits():
    from bmfr_b200 import synth
    return synth.limits()
