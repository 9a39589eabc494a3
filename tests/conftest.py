import ctypes
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "collab-splats_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hostmath():
    """ctypes handle of the test-only host build of csrc/rade_math.cuh."""
    import __graft_entry__ as ge
    return ctypes.CDLL(str(ge.build_hostmath()))


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from radegs_b200 import backend
    backend.load()  # fails loudly if the extension was not built
    return torch.device("cuda:0")
