"""The committed golden fixtures (tests/golden/*.npz: outputs of the reference's own library, oracle/make_golden.py)
against the library under test.  CPU: the reference library built here must still reproduce them (the fixtures are
current).  GPU: the CUDA path against the same vectors -- list rows may differ only where the oracle sees a pair inside
the stated ulp band, sums agree to the REFERENCE's FP32 accumulation accuracy (the tight 1e-6 bar is against the fp64
oracle, tests/test_regf_gpu.py)."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from nbody6ppgpu_b200 import snapshots as S

ROOT = Path(__file__).resolve().parent.parent
BAND_ULP = 4.0


def check_against_golden(lib, oracle, acc_tol, pot_tol):
    files = sorted((ROOT / "tests" / "golden").glob("regf_*.npz"))
    assert files, "golden fixtures missing"
    for f in files:
        g = np.load(f)
        n, m_flag, lmax, nnbmax = int(g["n"]), int(g["m_flag"]), int(g["lmax"]), int(g["nnbmax"])
        m, x, v = S.plummer(n, int(g["seed"]), str(g["imf"]))
        h2, dtr = S.radii(x, m, float(g["rs0"]), 0.125, m_flag)
        sel = g["isel"]
        lib.open(n + 10, 0)
        lib.send(m, x, v)
        acc, jrk, pot, lst = lib.regf(h2[sel], dtr[sel], x[sel], v[sel], lmax, nnbmax, m_flag)
        lib.close()
        band = oracle.regf_f64(m, x, v, h2[sel], dtr[sel], x[sel], v[sel], lmax, nnbmax, m_flag, BAND_ULP)[4]
        differ = oracle_lib.list_rows_equal(lst, g["list"])
        outside = [i for i in differ if band[i] > BAND_ULP]
        assert not outside, (f.name, outside[:5])
        ok = (g["list"][:, 0] >= 0) & (lst[:, 0] >= 0)
        ok[differ] = False                                  # a band flip moves one pair between the two sums
        assert ok.any(), f.name
        assert oracle_lib.relerr(acc[ok], g["acc"][ok]) < acc_tol, f.name
        assert oracle_lib.relerr(pot[ok], g["pot"][ok]) < pot_tol, f.name
    for f in sorted((ROOT / "tests" / "golden").glob("pot_*.npz")):
        g = np.load(f)
        m, x, v = S.plummer(int(g["n"]), int(g["seed"]), str(g["imf"]))
        phi = lib.gpupot(int(g["istart"]), int(g["ni"]), m, x)
        assert np.max(np.abs(phi - g["pot"]) / g["pot"]) < 5e-6, f.name


def test_reference_library_reproduces_the_golden_fixtures(ref_avx, oracle):
    if ref_avx is None:
        pytest.skip("oracle/_ref not built")
    check_against_golden(ref_avx, oracle, 1e-12, 1e-12)


@pytest.mark.gpu
def test_cuda_path_against_golden_fixtures(b200, oracle):
    check_against_golden(b200, oracle, 5e-5, 5e-5)
