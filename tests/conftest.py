import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def cuda_api():
    """The product library through its C ABI. Fails loudly when it is missing: no fallback."""
    from vcfx_b200 import api
    api.load()
    if api.device_count() <= 0:
        pytest.fail("no CUDA device visible to libvcfx_cuda (gpu-marked test run without a GPU)")
    return api
