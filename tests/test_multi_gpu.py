"""Multi-GPU exchange step on real GPUs (needs >= 2 devices; skipped on a 1-GPU box).
Each case runs in a fresh process because device discovery happens once per process (like the reference)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
WORKER = str(ROOT / "tests" / "multi_gpu_worker.py")


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("G", [2, 4, 8])
def test_inprocess_multi_gpu(G):
    if _ngpu() < G:
        pytest.skip(f"needs {G} GPUs")
    r = subprocess.run([sys.executable, WORKER, "inproc", str(G)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("G", [2, 4, 8])
def test_nccl_one_process_per_gpu(G):
    if _ngpu() < G:
        pytest.skip(f"needs {G} GPUs")
    port = 29600 + G
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={G}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, "nccl"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
