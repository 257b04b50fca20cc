import os
import sys
from pathlib import Path

import pytest

# the reference AVX library asserts omp threads <= 32 (reg.avx.cpp:7,103); set before libgomp loads
os.environ["OMP_NUM_THREADS"] = str(max(1, min(len(os.sched_getaffinity(0)), 32)))

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
