import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from ctgcn_b200 import _lib
    return _lib


@pytest.fixture(scope="session")
def cuda_device(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    assert lib.lib.ctgcn_device_check() == 0, lib.last_error()
    return torch.device("cuda:0")
