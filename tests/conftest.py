"""pytest configuration: registers the ``gpu`` marker and makes the repo root importable.

``python -m pytest tests -m "not gpu"`` runs on a CPU-only box (oracle vs golden vectors, host
logic, C-ABI symbol checks, gloo world_size-2 plumbing); ``-m gpu`` needs a B200 and runs the
parity tests proper through the C-ABI.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class _G:
        def __getitem__(self, name):
            return np.load(GOLDEN / f"{name}.npz")

    return _G()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
