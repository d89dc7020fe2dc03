import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The CPU oracle runs on oneDNN: on boxes whose CPU quota is far below the visible core count (GPU box:
    # 128 visible, 16 usable) the default thread count makes it ~25x slower.
    import torch
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    import migan_b200
    from migan_b200 import _abi

    migan_b200.build.build()
    return _abi.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    return torch.device("cuda:0")
