import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def po():
    """the CPU oracle (test infrastructure)"""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_lib():
    """librmcrt_b200.so + a visible CUDA device; fails loudly (never skips to a fallback) on a GPU run"""
    import torch
    from uintah_b200 import capi
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return capi.load_library()
