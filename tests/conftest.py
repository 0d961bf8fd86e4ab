import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ss-gnn_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The CUDA path: fails loudly (no fallback) if the library is not built or no GPU is present."""
    import torch
    from ugs_b200 import _lib
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    return _lib.lib()
