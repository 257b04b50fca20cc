"""oracle/regcor_oracle.c (the CPU restatement of util_gpu.F:102-111 + regcor_gpu.F:263-459 that the GPU test checks the
CUDA path against).

THE PIN: the reference's own Fortran text, executed statement by statement by oracle/f77_interp.py (this image has no
Fortran compiler).  tests/golden/regcor_f77_*.npz are its outputs on seeded rows (oracle/make_regcor_golden.py; the source
is read under /root/reference, nothing is copied) -- they travel to machines without the reference; where the reference is
present the interpreter also runs live on larger cases.  Beside the pin, two further independent statements of the same
Fortran: a hand-made GO TO transcription and plain set differences / vectorised sums (tests/regcor_cases.py)."""
import sys
from pathlib import Path

import numpy as np
import pytest

import regcor_cases as RC

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import make_regcor_golden as MG  # noqa: E402
import regcor_fortran as RF  # noqa: E402

GOLDEN = sorted((ROOT / "tests" / "golden").glob("regcor_f77_*.npz"))


@pytest.fixture(scope="module")
def case():
    return RC.make_case(overflow_rows=(5,), empty_old_rows=(0, 17))


def run_oracle(oracle, c, step=True, rows=None):
    s = slice(None) if rows is None else rows
    return oracle.regcor(c["index_i"][s], c["ifirst"], c["n"], c["ntot"], c["new"][s], c["old"][s], c["m"], c["x"], c["v"],
                         c["rs2"][s], c["step"] if step else None, c["smin"], c["nnbmax"], c["freg"][s], c["fdr"][s])


def test_golden_fixtures_are_present():
    assert len(GOLDEN) == 5, "tests/golden/regcor_f77_*.npz missing: run oracle/make_regcor_golden.py where /root/reference exists"


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_oracle_against_the_golden_vectors_of_the_interpreted_fortran(oracle, path):
    """Every integer the reference's text leaves behind (NNB, NLIST, NBLOSS, NBGAIN, JJLIST, NBSMIN) equal, every fp64 result
    (FREG, FDR, DFIRR, DFD) bit for bit."""
    c, g = MG.load_case(path)
    out = run_oracle(oracle, c)
    rows = [r for r in range(c["index_i"].shape[0]) if g["f77_valid"][r]]
    assert len(rows) >= 96
    retained = RC.compare_rows(out, c, rows, MG.golden_walk(g))
    assert retained == out["nbsmin"] == int(g["f77_nbsmin"].sum())
    assert str(g["source"]).startswith("src/Main/util_gpu.F:102-111 + src/Main/regcor_gpu.F:263-459")


@pytest.mark.skipif(not RF.available(), reason="the reference sources are not on this machine (the golden vectors cover it)")
def test_interpreter_live_on_the_reference_text(oracle, case):
    """Where /root/reference exists: the fixtures are what the interpreter produces from the text as it lies there today
    (fingerprint of the interpreted statements), and the oracle equals the interpreted Fortran on the large module case, on
    the rows without a STEP array, and on a fresh random case no fixture holds."""
    _, g = MG.load_case(GOLDEN[0])
    assert str(g["source_sha256"]) == RF.source_fingerprint()
    c = case
    out = run_oracle(oracle, c)
    rows = [r for r in range(c["index_i"].shape[0]) if c["new"][r, 0] >= 0]
    assert RC.compare_rows(out, c, rows, RF.interpreted_walk) == out["nbsmin"] > 0
    out0 = run_oracle(oracle, c, step=False)
    RC.compare_rows(out0, c, rows[:64], lambda cc, r: RF.interpreted_walk(cc, r, use_step=False))
    c2 = RC.make_random_case(seed=1234)
    out2 = run_oracle(oracle, c2)
    assert RC.compare_rows(out2, c2, range(c2["index_i"].shape[0]), RF.interpreted_walk) == out2["nbsmin"] > 30
    # and the hand transcription says the same as the text it transcribes
    for r in rows[:40]:
        a, b = RF.interpreted_walk(c, r), RC.fortran_walk(c, r)
        assert all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in a), r


def test_oracle_matches_the_literal_fortran_walk(oracle, case):
    c = case
    out = run_oracle(oracle, c)
    rows = [r for r in range(c["index_i"].shape[0]) if c["new"][r, 0] >= 0]
    retained = RC.compare_rows(out, c, rows, RC.fortran_walk)
    assert retained > 0, "the case must exercise the retention branch (regcor_gpu.F:338-420)"
    assert out["nbsmin"] == retained
    assert out["nbloss"].sum() > 100 and out["nbgain"].sum() > 100
    # overflow rows pass through untouched
    assert out["nlist"][5, 0] == c["new"][5, 0] and out["nbloss"][5] == 0 and out["nbgain"][5] == 0
    assert np.array_equal(out["freg"][5], c["freg"][5])
    # NNB0 = 0: everything gained, filed from JJLIST(1) (regcor_gpu.F:271-283)
    for r in (0, 17):
        assert out["nbloss"][r] == 0 and out["nbgain"][r] == out["nlist"][r, 0]
        assert list(out["jjlist"][r, :out["nbgain"][r]]) == list(out["nlist"][r, 1:1 + out["nlist"][r, 0]])


def test_oracle_without_steps_is_the_set_difference(oracle, case):
    c = case
    out = run_oracle(oracle, c, step=False)
    assert out["nbsmin"] == 0
    for r in range(c["index_i"].shape[0]):
        if c["new"][r, 0] < 0 or c["old"][r, 0] == 0:
            continue
        s = RC.sets_and_sums(c, r)
        nl = out["nlist"][r]
        assert list(nl[1:1 + nl[0]]) == s["members"]
        assert list(out["jjlist"][r, :out["nbloss"][r]]) == s["lost"]
        nnb0 = c["old"][r, 0]
        assert list(out["jjlist"][r, nnb0:nnb0 + out["nbgain"][r]]) == s["gained"]
        scale = np.abs(s["dfirr"]).max() + 1e-300
        assert np.allclose(out["dfirr"][r], s["dfirr"], rtol=0, atol=1e-11 * max(scale, 1.0))
        assert np.allclose(out["dfd"][r], s["dfd"], rtol=0, atol=1e-10 * max(np.abs(s["dfd"]).max(), 1.0))
        assert np.array_equal(out["freg"][r], c["freg"][r]) and np.array_equal(out["fdr"][r], c["fdr"][r])


def test_rows_are_independent(oracle, case):
    c = case
    full = run_oracle(oracle, c)
    part = run_oracle(oracle, c, rows=slice(40, 90))
    for k in ("nlist", "nbloss", "nbgain", "jjlist", "freg", "fdr", "dfirr", "dfd"):
        assert np.array_equal(full[k][40:90], part[k]), k


def test_cm_body_rows_are_never_retained(oracle):
    # I > N (a c.m. body) skips the retention loop at its first test (regcor_gpu.F:342)
    c = RC.make_case(n_tot=1200, ni=1200, n_cm=300, seed=9, lmax=96, nnb_mean=20.0)
    out = run_oracle(oracle, c)
    cm = c["index_i"] > c["n"]
    assert cm.any()
    RC.compare_rows(out, c, list(np.nonzero(cm)[0][:60]) + list(np.nonzero(~cm)[0][:60]), RC.fortran_walk)
    assert np.array_equal(out["freg"][cm], c["freg"][cm])


@pytest.mark.parametrize("smin_mode,nnbmax", [("case", None), ("all_small", None), ("case", 20)])
def test_edge_rows(oracle, smin_mode, nnbmax):
    """Hand-built rows (only self, both empty, identical, no self, longest rows) and the two ends of the retention step:
    every lost member has a small step (all of them inside 2 RS are put back), NNB > NNBMAX at entry (none is)."""
    c = RC.make_edge_case(seed=21, ni=120, n_tot=1500, lmax=96, nnb_mean=24.0)
    if smin_mode == "all_small":
        c["smin"] = 10.0
    if nnbmax is not None:
        c["nnbmax"] = nnbmax
    out = run_oracle(oracle, c)
    retained = RC.compare_rows(out, c, range(120), RC.fortran_walk)
    assert out["nbsmin"] == retained
    assert out["nbgain"][1] == 0 and out["nlist"][1, 0] + out["nbloss"][1] == c["old"][1, 0]      # row 1: every old member is lost or put back
    assert out["nlist"][2, 0] == 0 and out["nbloss"][2] == 0 and out["nbgain"][2] == 0
    assert out["nbloss"][3] == 0 and out["nbgain"][3] == 0
    assert out["nbloss"][6] >= 40 and out["nbgain"][6] >= 39
    if smin_mode == "all_small":
        assert retained > 50


def test_random_small_lists_against_the_literal_walk(oracle):
    """400 random rows over a universe of 60 particles (short lists, many empty or nearly empty ones, every second row with a
    c.m. particle or NNB near NNBMAX): the corners of the walk and of the retention loop that a physical snapshot rarely
    visits.  Oracle == literal Fortran transcription, integers and bits."""
    c = RC.make_random_case()
    ni = c["index_i"].shape[0]
    out = run_oracle(oracle, c)
    retained = RC.compare_rows(out, c, range(ni), RC.fortran_walk)
    assert out["nbsmin"] == retained and retained > 30
    assert (out["nlist"][:, 0] == 0).sum() > 3 and (c["old"][:, 0] == 0).sum() > 10
