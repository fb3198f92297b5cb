import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def capi():
    from parallelfdtd_b200 import capi as c
    c.lib()   # raises loudly when libpfdtd_b200.so is missing: there is no fallback
    return c


@pytest.fixture(scope="session")
def gpu(capi):
    n = capi.device_count()
    if n < 1:
        pytest.skip("no CUDA device in this container (GPU tests run through gpurun)")
    return n
