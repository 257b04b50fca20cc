#!/usr/bin/env python
"""Generate tests/golden/regcor_f77_*.npz: inputs and outputs of the reference's OWN Fortran text for the neighbour-list
bookkeeping (util_gpu.F:102-111 + regcor_gpu.F:263-459), executed statement by statement by oracle/f77_interp.py on seeded
rows.  Test infrastructure only.  This is the pin of oracle/regcor_oracle.c: the image has no Fortran compiler, so the
reference source is interpreted where it lies (/root/reference, nothing copied); the vectors travel to machines without it.

Run in the build container:   python oracle/make_regcor_golden.py
Every fixture holds the complete inputs of its rows (snapshot, lists, steps, forces) and, per row, what the Fortran left in
NNB / NLIST / NBLOSS / NBGAIN / JJLIST / FREG / FDR / DFIRR / DFD / NBSMIN, plus a sha256 of the interpreted statements.
Before writing, every row is also compared with oracle/regcor_oracle.c (integers equal, fp64 bit for bit).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "oracle"))
import oracle_lib  # noqa: E402
import regcor_cases as RC  # noqa: E402
import regcor_fortran as RF  # noqa: E402

OUT = ROOT / "tests" / "golden"
INPUT_KEYS = ("m", "x", "v", "index_i", "new", "old", "rs2", "step", "freg", "fdr")
SCALAR_KEYS = ("ifirst", "n", "ntot", "lmax", "nnbmax", "smin")


def cases():
    yield "plummer", RC.make_case(n_tot=700, ni=96, lmax=96, nnb_mean=24.0, seed=31, empty_old_rows=(0, 17))
    e = RC.make_edge_case(seed=21, ni=120, n_tot=900, lmax=96, nnb_mean=24.0)
    e["smin"] = 10.0                                   # every lost member has a small step: the retention loop at full length
    yield "edge_all_small_steps", e
    e = RC.make_edge_case(seed=22, ni=120, n_tot=900, lmax=96, nnb_mean=24.0)
    e["nnbmax"] = 20                                   # NNB > NNBMAX at entry: nobody is retained (regcor_gpu.F:342)
    yield "edge_nnbmax", e
    yield "random_short_rows", RC.make_random_case()
    yield "cm_bodies", RC.make_case(n_tot=500, ni=160, n_cm=150, seed=9, lmax=96, nnb_mean=20.0)


def walk_all(c):
    ni, lmax = c["index_i"].shape[0], int(c["lmax"])
    out = dict(valid=np.zeros(ni, np.int8), nnb=np.zeros(ni, np.int32), members=np.zeros((ni, lmax), np.int32),
               nnb0=np.zeros(ni, np.int32), nbloss=np.zeros(ni, np.int32), nbgain=np.zeros(ni, np.int32),
               lost=np.zeros((ni, lmax), np.int32), gained=np.zeros((ni, lmax), np.int32), nbsmin=np.zeros(ni, np.int32),
               freg=np.zeros((ni, 3)), fdr=np.zeros((ni, 3)), dfirr=np.zeros((ni, 3)), dfd=np.zeros((ni, 3)))
    for r in range(ni):
        if c["new"][r, 0] < 0:
            continue                                   # overflow rows never reach the bookkeeping
        w = RF.interpreted_walk(c, r)
        out["valid"][r] = 1
        for k in ("nnb", "nnb0", "nbloss", "nbgain", "nbsmin"):
            out[k][r] = w[k]
        out["members"][r, :w["nnb"]] = w["members"]
        out["lost"][r, :w["nbloss"]] = w["lost"]
        out["gained"][r, :w["nbgain"]] = w["gained"]
        for k in ("freg", "fdr", "dfirr", "dfd"):
            out[k][r] = w[k]
    return out


def golden_walk(g):
    """walk(c, r) over a loaded fixture, in the shape tests/regcor_cases.compare_rows expects."""
    def walk(c, r):
        assert g["f77_valid"][r]
        nnb, nbloss, nbgain = int(g["f77_nnb"][r]), int(g["f77_nbloss"][r]), int(g["f77_nbgain"][r])
        return dict(nnb=nnb, members=list(g["f77_members"][r, :nnb]), nnb0=int(g["f77_nnb0"][r]), nbloss=nbloss, nbgain=nbgain,
                    lost=list(g["f77_lost"][r, :nbloss]), gained=list(g["f77_gained"][r, :nbgain]),
                    freg=g["f77_freg"][r], fdr=g["f77_fdr"][r], dfirr=g["f77_dfirr"][r], dfd=g["f77_dfd"][r],
                    nbsmin=int(g["f77_nbsmin"][r]))
    return walk


def load_case(path):
    """(case dict in the shape of regcor_cases.make_case, fixture) from a golden file."""
    g = np.load(path)
    c = {k: g[k] for k in INPUT_KEYS}
    for k in SCALAR_KEYS:
        c[k] = float(g[k]) if k == "smin" else int(g[k])
    return c, g


def main():
    if not RF.available():
        raise SystemExit("needs the reference sources under %s" % RF.REFERENCE)
    OUT.mkdir(parents=True, exist_ok=True)
    oracle = oracle_lib.Oracle()
    fp = RF.source_fingerprint()
    for name, c in cases():
        out = walk_all(c)
        # the oracle against the interpreted Fortran before anything is written
        o = oracle.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c["old"], c["m"], c["x"], c["v"], c["rs2"],
                          c["step"], c["smin"], c["nnbmax"], c["freg"], c["fdr"])
        g = {("f77_" + k): v for k, v in out.items()}
        rows = [r for r in range(c["index_i"].shape[0]) if out["valid"][r]]
        retained = RC.compare_rows(o, c, rows, golden_walk(g))
        assert retained == o["nbsmin"]
        path = OUT / ("regcor_f77_%s.npz" % name)
        np.savez_compressed(path, source_sha256=np.array(fp), source=np.array("%s:%d-%d + %s:%d-%d" % (RF.UTIL_GPU + RF.REGCOR)),
                            **{k: np.asarray(c[k]) for k in INPUT_KEYS + SCALAR_KEYS}, **g)
        print("%-24s rows %4d  lost %5d  gained %5d  retained %4d  empty new %3d  empty old %3d  -> %s (%d KB)" % (
            name, len(rows), int(out["nbloss"].sum()), int(out["nbgain"].sum()), retained, int((out["nnb"][rows] == 0).sum()),
            int((out["nnb0"][rows] == 0).sum()), path.name, path.stat().st_size // 1024))


if __name__ == "__main__":
    main()
