import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def avc():
    """The built C-ABI library through its ctypes binding (built on demand, in-tree)."""
    import __graft_entry__ as entry

    entry.build()
    import audio_visual_llm_b200 as pkg

    return pkg


@pytest.fixture(scope="session")
def cuda_dev(avc):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    avc._lib.require_device(0)
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
