#!/usr/bin/env python
"""Generate tests/golden/regint_f77.npz: the reference's fp64 regular-force statement REGINT (src/Main/regint.f lines 28-79:
neighbour-sphere limit, the J loop with force, derivative, velocity criterion, list and potential) executed by
oracle/f77_interp.py for a few i-particles of a seeded Plummer model, m_flag 0 and 1.  Test infrastructure: the pin of the
fp64 statement in oracle/regf_oracle.c that the 1e-6 force / jerk / potential bar is measured against (SURVEY 8a row a15,
8c).  The source is read where it lies (/root/reference); nothing is copied.

Run in the build container:   python oracle/make_regint_golden.py
Differences REGINT has from the GPU contract: membership in fp64 with '<=' (GPU: fp32 with '<'), POT includes the
neighbours, self is skipped, the list is 1-based without self.  The fixture stores REGINT's results as they are.
"""
import hashlib
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
from f77_interp import FArray, Machine, farray_numpy, read_statements  # noqa: E402

REFERENCE = os.environ.get("NBODY6_REFERENCE", "/root/reference")
SPEC = ("src/Main/regint.f", 28, 79)
OUT = ROOT / "tests" / "golden" / "regint_f77.npz"
LMAX = 128


def statements():
    return read_statements(os.path.join(REFERENCE, SPEC[0]), SPEC[1], SPEC[2])


def interpreted_regint(m, x, v, rs, dtr, isel, m_flag, bodym):
    """FREG, FDR, POT, LISTGP rows ([count, members 1-based]) of particles isel (0-based) against all n (IFIRST = 1, N = n)."""
    n, ni = m.shape[0], len(isel)
    freg, fdr, pot = np.full((ni, 3), np.nan), np.full((ni, 3), np.nan), np.zeros(n)
    listgp = np.zeros((ni, LMAX), dtype=np.int64)
    dv, dp = np.zeros(3), np.zeros(3)
    arrays = {"X": farray_numpy(x), "XDOT": farray_numpy(v), "BODY": farray_numpy(m), "RS": farray_numpy(rs),
              "DTR": farray_numpy(dtr), "FREG": farray_numpy(freg), "FDR": farray_numpy(fdr), "POT": farray_numpy(pot),
              "DV": farray_numpy(dv), "DP": farray_numpy(dp),
              "LISTGP": FArray(lambda ix: int(listgp[ix[1] - 1, ix[0] - 1]), lambda ix, val: listgp.__setitem__((ix[1] - 1, ix[0] - 1), val))}
    st = statements()
    for ii, i in enumerate(isel):
        env = Machine(st, {"I": int(i) + 1, "II": ii + 1, "IFIRST": 1, "N": n, "M_FLAG": int(m_flag), "BODYM": float(bodym),
                           "LMAX": LMAX}, arrays).run()
        listgp[ii, 0] = env["NNB"] - 1                  # regint.f:96
    return freg, fdr, pot[np.asarray(isel)], listgp


def make_case(n=800, seed=17, ni=40):
    from nbody6ppgpu_b200 import snapshots as S
    m, x, v = S.plummer(n, seed, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 30.0))
    rs = np.sqrt(h2)
    rng = np.random.default_rng(seed)
    isel = np.sort(rng.choice(n, ni, replace=False))
    return m, x, v, rs, dtr, isel, float(m.mean())


def main():
    if not os.path.isfile(os.path.join(REFERENCE, SPEC[0])):
        raise SystemExit("needs the reference sources under %s" % REFERENCE)
    m, x, v, rs, dtr, isel, bodym = make_case()
    out = {}
    for mf in (0, 1):
        freg, fdr, pot, lst = interpreted_regint(m, x, v, rs, dtr, isel, mf, bodym)
        out.update({"f77_freg_m%d" % mf: freg, "f77_fdr_m%d" % mf: fdr, "f77_pot_m%d" % mf: pot, "f77_list_m%d" % mf: lst.astype(np.int32)})
        print("m_flag %d: neighbours per row min %d mean %.1f max %d" % (mf, lst[:, 0].min(), lst[:, 0].mean(), lst[:, 0].max()))
    fp = hashlib.sha256("\n".join("%s|%s|%d" % s for s in statements()).encode()).hexdigest()
    np.savez_compressed(OUT, m=m, x=x, v=v, rs=rs, dtr=dtr, isel=isel.astype(np.int32), bodym=np.array(bodym), lmax=np.array(LMAX),
                        source=np.array("%s:%d-%d" % SPEC), source_sha256=np.array(fp), **out)
    print("regint: %d i x %d j, two m_flags -> %s (%d KB)" % (len(isel), m.shape[0], OUT.name, OUT.stat().st_size // 1024))


if __name__ == "__main__":
    main()
