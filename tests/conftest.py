import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref_avx():
    """The reference's own CPU library (oracle/_ref), one instance per session (file-static state)."""
    import oracle_lib
    return oracle_lib.ref_avx()


@pytest.fixture(scope="session")
def b200():
    """The product library.  GPU tests only; fails loudly when it is not built."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nbody6ppgpu_b200 import load
    lib = load()
    lib.devinit(0)
    return lib
