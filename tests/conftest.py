import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from fastvideotagging_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the hot path has no CPU fallback")
    from fastvideotagging_b200 import _lib
    _lib.check(lib.fvt_device_check(0))
    return torch.device("cuda:0")
