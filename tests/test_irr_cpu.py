"""Irregular-force interface.  CPU: the fp64 statement nbody6ppgpu_b200/irr.py:firr_f64 is pinned
against the reference's own AVX library (oracle/_ref/libirr_ref_avx.so, compiled from src/Main/irr.avx.cpp), and the CUDA
library exports the reference's six symbols.  GPU: the CUDA library against the fp64 statement (validated on B200 in
session r2a: forces to 2e-14, nearest-neighbour addresses identical)."""
import ctypes
import os
import re
from pathlib import Path

import numpy as np
import pytest

from nbody6ppgpu_b200 import irr
from nbody6ppgpu_b200 import snapshots as S

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "libirr_ref_avx.so"


def make_case(n=3000, seed=2, nact=500):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    m, x, v = S.plummer(n, seed, "kroupa")
    a2 = 0.5 * rng.normal(size=(n, 3)); j6 = rng.normal(size=(n, 3)) / 6.0
    t0 = rng.integers(0, 16, size=n) * 2.0 ** -8
    tree = cKDTree(x)
    lists = []
    for i in range(n):
        k = int(rng.integers(1, 60))
        lists.append(np.sort(tree.query(x[i], k=k + 1)[1][1:]) + 1)          # 1-based, ascending, self excluded
    addr = np.sort(rng.choice(n, nact, replace=False)).astype(np.int32) + 1
    return m, x, v, a2, j6, t0, lists, addr


def run(lib, case, ti=0.07, lmax=128):
    m, x, v, a2, j6, t0, lists, addr = case
    n = m.shape[0]
    lib.open(n, lmax, 0)
    try:
        for i in range(n):
            lib.set_jp(i + 1, x[i], v[i], a2[i], j6[i], m[i], t0[i])
            lib.set_list(i + 1, irr.pad_list(lists[i]))
        # a particle and a list set twice before the next force call: the later values win
        lib.set_jp(1, x[0], v[0], a2[0], j6[0], m[0], t0[0]); lib.set_list(1, irr.pad_list(lists[0]))
        out = lib.firr_vec(ti, addr)
        lib.profile(0)
    finally:
        lib.close(0)
    return out


def test_fp64_statement_is_pinned_against_the_reference_library():
    if not REF.exists():
        pytest.skip("oracle/_ref/libirr_ref_avx.so not built")
    case = make_case()
    acc, jrk, nn = run(irr.IrrLib(REF), case)
    a64, j64, n64 = irr.firr_f64(0.07, case[7], case[6], case[1], case[2], case[3], case[4], case[0], case[5])
    rel = lambda a, b: float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)))
    assert rel(acc, a64) < 5e-6 and rel(jrk, j64) < 2e-5          # the reference predicts and sums in FP32
    assert np.array_equal(nn, n64)


def test_cuda_library_exports_the_reference_symbols():
    so = irr.lib_path()
    assert so.exists(), "libirr_b200.so not built: run __graft_entry__.build()"
    hdr = (ROOT / "include" / "irr_b200.h").read_text()
    names = set(re.findall(r"\b(irr_simd_[a-z_]+_|irr_b200_[a-z0-9_]+)\s*\(", hdr))
    assert {"irr_simd_open_", "irr_simd_close_", "irr_simd_profile_", "irr_simd_set_jp_", "irr_simd_set_list_",
            "irr_simd_firr_vec_"} <= names
    lib = ctypes.CDLL(str(so))
    for nme in sorted(names):
        assert hasattr(lib, nme), nme


@pytest.mark.gpu
def test_cuda_library_against_the_fp64_statement():
    case = make_case()
    acc, jrk, nn = run(irr.IrrLib(irr.lib_path()), case)
    a64, j64, n64 = irr.firr_f64(0.07, case[7], case[6], case[1], case[2], case[3], case[4], case[0], case[5])
    rel = lambda a, b: float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)))
    assert rel(acc, a64) < 1e-11 and rel(jrk, j64) < 1e-10
    assert np.array_equal(nn, n64)
