"""examples/fpoly0_sweep.c: a COMPILED caller of the reference ABI (what the Fortran side does: send, regf per block of
1024 with the overflow retry loop, self removal, gpupot) linked against the reference's AVX library and against this
repo's library without a line of difference."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "examples" / "fpoly0_sweep.c"


def build(tmp_path, which):
    exe = tmp_path / f"fpoly0_{which}"
    if which == "avx":
        libdir, lib, extra = ROOT / "oracle" / "_ref", "gpunb_ref_avx", ["-DNO_DEVINIT"]
    else:
        libdir, lib, extra = ROOT / "nbody6ppgpu_b200", "gpunb_b200", []
    if not (libdir / f"lib{lib}.so").exists():
        pytest.skip(f"lib{lib}.so not built")
    cmd = ["gcc", "-O2", *extra, str(SRC), "-o", str(exe), f"-L{libdir}", f"-l{lib}", f"-Wl,-rpath,{libdir}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run(exe, *args):
    env = dict(os.environ, OMP_NUM_THREADS=os.environ.get("OMP_NUM_THREADS", "4"), GPU_LIST="0")
    r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("FPOLY0 ")][-1].split()
    return {line[k]: line[k + 1] for k in range(1, len(line), 2)}


def test_compiled_caller_links_both_libraries_and_runs_on_the_reference(tmp_path):
    build(tmp_path, "b200")                       # link check only: no GPU here
    avx = build(tmp_path, "avx")
    out = run(avx, 2048, 1, 48, 400)
    assert abs(float(out["etot"]) + 0.25) < 0.03            # Plummer sphere in N-body units
    assert 0.6 * 48 < float(out["mean_nnb"]) < 1.6 * 48
    assert int(out["retries"]) == 0
    tight = run(avx, 2048, 1, 30, 90)                        # lmax 90 -> nnbmax 40, NNBOPT 30: dense rows overflow, RS shrinks
    assert int(tight["retries"]) > 0 and float(tight["mean_nnb"]) < 40.0


@pytest.mark.gpu
def test_compiled_caller_same_results_behind_both_libraries(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    a = run(build(tmp_path, "avx"), 20000, 3, 64, 400)
    b = run(build(tmp_path, "b200"), 20000, 3, 64, 400)
    print("avx ", a); print("b200", b)
    assert a["retries"] == b["retries"]
    assert abs(float(a["mean_nnb"]) - float(b["mean_nnb"])) * 20000 <= 2          # band flips only
    for k, tol in (("fsum", 1e-6), ("psum", 1e-6), ("jsum", 3e-5), ("etot", 1e-6)):
        assert abs(float(a[k]) - float(b[k])) <= tol * abs(float(a[k])), (k, a[k], b[k])
    # overflow path: identical retry count and final lists behind both libraries
    a = run(build(tmp_path, "avx"), 6000, 5, 36, 100)
    b = run(build(tmp_path, "b200"), 6000, 5, 36, 100)
    assert int(a["retries"]) > 0 and a["retries"] == b["retries"]
    assert abs(float(a["mean_nnb"]) - float(b["mean_nnb"])) * 6000 <= 2
