import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand with nvcc; no GPU needed to build or load it)."""
    from magma_b200 import _lib, build
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def gpu_queue(lib):
    import torch
    assert torch.cuda.is_available(), "gpu-marked test run without a CUDA device"
    from magma_b200 import Queue, magma_init
    assert magma_init() == 0
    torch.cuda.set_device(0)
    q = Queue.from_torch(0)
    yield q
