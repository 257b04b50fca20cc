"""The reference's fp64 statement REGINT (regint.f:28-79) executed by oracle/f77_interp.py -- tests/golden/regint_f77.npz,
oracle/make_regint_golden.py -- against the fp64 statement of oracle/regf_oracle.c that the 1e-6 force / jerk / potential
bar is measured against: same membership => the same sums to rounding (the two evaluate
m r^-3 with differently ordered but mathematically identical factors), and the fp32 GPU predicate picks the same members
as REGINT's fp64 '<=' test away from the RS boundary."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import make_regint_golden as MR  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "regint_f77.npz"


def abi_rows(g, mf):
    """REGINT rows (1-based J, self skipped) -> gpunb_regf_ rows (0-based j ascending, self included)."""
    lst, isel = g["f77_list_m%d" % mf], g["isel"]
    rows = np.zeros((len(isel), int(g["lmax"]) + 2), dtype=np.int32)
    for r, i in enumerate(isel):
        mem = np.sort(np.append(lst[r, 1:1 + lst[r, 0]] - 1, i))
        rows[r, 0] = mem.size; rows[r, 1:1 + mem.size] = mem
    return rows


def relerr(a, b):
    return (np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)).max()


@pytest.mark.parametrize("mf", [0, 1])
def test_fp64_oracle_equals_the_interpreted_regint(oracle, mf):
    g = np.load(GOLDEN)
    m, x, v, isel = g["m"], g["x"], g["v"], g["isel"]
    rows = abi_rows(g, mf)
    acc, jrk, pot = oracle.regf_f64_given_list(m, x, v, x[isel], v[isel], rows)[:3]
    assert relerr(acc, g["f77_freg_m%d" % mf]) < 1e-13
    assert relerr(jrk, g["f77_fdr_m%d" % mf]) < 1e-12
    # REGINT's POT includes the neighbours (regint.f:78), the GPU contract excludes them: add their part back
    full = pot.copy()
    for r, i in enumerate(isel):
        nb = rows[r, 1:1 + rows[r, 0]]
        nb = nb[nb != i]
        full[r] += (m[nb] / np.linalg.norm(x[nb] - x[i], axis=1)).sum()
    assert np.abs(full / g["f77_pot_m%d" % mf] - 1.0).max() < 1e-13
    # membership: the fp32 predicate of the GPU contract against REGINT's fp64 one -- identical on this seeded case (a pair
    # would have to sit within an fp32 ulp of the boundary to differ)
    h2 = g["rs"][isel] ** 2 / (float(g["bodym"]) if mf else 1.0)
    lst = oracle.regf_f64(m, x, v, h2, g["dtr"][isel], x[isel], v[isel], rows.shape[1], rows.shape[1] - 2, mf)[3]
    for r in range(len(isel)):
        assert list(lst[r, :lst[r, 0] + 1]) == list(rows[r, :rows[r, 0] + 1]), r
    assert rows[:, 0].max() >= 12


@pytest.mark.skipif(not Path(MR.REFERENCE, MR.SPEC[0]).is_file(), reason="the reference sources are not on this machine")
def test_interpreter_live_on_the_reference_text():
    g = np.load(GOLDEN)
    sel = g["isel"][:3]
    freg, fdr, pot, lst = MR.interpreted_regint(g["m"], g["x"], g["v"], g["rs"], g["dtr"], sel, 1, float(g["bodym"]))
    assert np.array_equal(freg, g["f77_freg_m1"][:3]) and np.array_equal(fdr, g["f77_fdr_m1"][:3])
    assert np.array_equal(pot, g["f77_pot_m1"][:3]) and np.array_equal(lst, g["f77_list_m1"][:3])
