import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _cap_threads():
    """torch's CPU ops (seeded weight generation, the fp32 oracle) are several times SLOWER with 128 intra-op threads
    than with 16 on the 128-way GPU host (measured, tools/cpu_threads.py) — cap them for the whole test session."""
    import torch

    torch.set_num_threads(min(16, os.cpu_count() or 1))


def pytest_configure(config):
    _cap_threads()
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ming_univision_b200 import _lib

    _lib.require_device()
    return torch.device("cuda:0")
