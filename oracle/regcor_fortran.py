"""regcor_fortran.py -- the reference's OWN Fortran text for the list bookkeeping, executed through oracle/f77_interp.py.

TEST INFRASTRUCTURE.  ``interpreted_walk(c, r)`` runs, for row r of a case of tests/regcor_cases.py,

  * /root/reference/src/Main/util_gpu.F    lines 102-111  (0-based GPU row -> NLIST: + IFIRST, self dropped), then
  * /root/reference/src/Main/regcor_gpu.F  lines 263-459  (DFIRR / DFD = 0, NBLOSS / NBGAIN walk, retention of small-step
                                                            neighbours with the FREG / FDR correction, the force swap)

statement by statement, with the reference's variable names bound to the case's arrays, and returns the same dict as the
hand transcription ``regcor_cases.fortran_walk``.  The source is read where it lies; nothing of it is stored here.  What
the text needs from the lines above it is bound explicitly: NNB = NLIST(1), NNB0 = LIST(1,I) (regcor_gpu.F:24-25),
RS2 = RS(I)**2 (:39), XI / XIDOT = the predicted X / XDOT of particle I (intgrt.F passes them), NBSMIN = 0.
"""
from __future__ import annotations

import os

import numpy as np

from f77_interp import FArray, Machine, farray_numpy, read_statements

REFERENCE = os.environ.get("NBODY6_REFERENCE", "/root/reference")
UTIL_GPU = ("src/Main/util_gpu.F", 102, 111)
REGCOR = ("src/Main/regcor_gpu.F", 263, 459)


def available():
    return all(os.path.isfile(os.path.join(REFERENCE, f)) for f, _, _ in (UTIL_GPU, REGCOR))


_cache = {}


def _statements(spec):
    if spec not in _cache:
        f, a, b = spec
        _cache[spec] = read_statements(os.path.join(REFERENCE, f), a, b)
    return _cache[spec]


def source_fingerprint():
    """sha256 of the two line ranges as interpreted (labels + statements): stored with the golden vectors."""
    import hashlib
    h = hashlib.sha256()
    for spec in (UTIL_GPU, REGCOR):
        for lab, txt, no in _statements(spec):
            h.update(("%s|%s|%d\n" % (lab, txt, no)).encode())
    return h.hexdigest()


def interpreted_walk(c, r, use_step=True):
    ifirst, n, ntot, lmax = int(c["ifirst"]), int(c["n"]), int(c["ntot"]), int(c["lmax"])
    I = int(c["index_i"][r])
    x, v, m = c["x"], c["v"], c["m"]
    step = c["step"] if use_step else np.full(x.shape[0], 1.0e30)       # no STEP array: nobody has a small step
    # ---- util_gpu.F:102-111 on LISTGP(:,II), II = 1 --------------------------------------------------------------------
    listgp = np.zeros(lmax + 4, dtype=np.int64)
    listgp[:lmax] = c["new"][r]
    nnb_gpu = int(listgp[0])
    if nnb_gpu < 0:
        raise ValueError("overflow rows never reach the bookkeeping (util_gpu.F:71-97)")
    arr = {"LISTGP": FArray(lambda ix: int(listgp[ix[0] - 1]), lambda ix, val: listgp.__setitem__(ix[0] - 1, val))}
    Machine(_statements(UTIL_GPU), {"NNB": nnb_gpu, "II": 1, "IDI": I, "IFIRST": ifirst}, arr).run()
    # ---- regcor_gpu.F:263-459 -----------------------------------------------------------------------------------------
    nlist = listgp                                                       # NLIST is LISTGP(1,II) passed down (util_gpu.F / intgrt.F)
    old = np.zeros(lmax + 4, dtype=np.int64)
    old[:lmax] = c["old"][r]
    jjlist = np.zeros(2 * lmax + 4, dtype=np.int64)
    freg, fdr = c["freg"][r].astype(np.float64).copy(), c["fdr"][r].astype(np.float64).copy()
    dfirr, dfd, dv = np.full(3, np.nan), np.full(3, np.nan), np.zeros(3)
    xi, xidot = x[I - ifirst].copy(), v[I - ifirst].copy()
    fr = np.zeros(3)

    def list_get(ix):
        assert ix[1] == I
        return int(old[ix[0] - 1])

    def list_set(ix, val):
        assert ix[1] == I
        old[ix[0] - 1] = val

    arrays = {
        "NLIST": farray_numpy(nlist, 1, integer=True), "JJLIST": farray_numpy(jjlist, 1, integer=True),
        "LIST": FArray(list_get, list_set),
        "FREG": farray_numpy(freg), "FDR": farray_numpy(fdr), "DFIRR": farray_numpy(dfirr), "DFD": farray_numpy(dfd),
        "DV": farray_numpy(dv), "XI": farray_numpy(xi), "XIDOT": farray_numpy(xidot),
        "FR": FArray(lambda ix: float(fr[ix[0] - 1])),
        "X": farray_numpy(x, ifirst), "XDOT": farray_numpy(v, ifirst),
        "BODY": farray_numpy(m, ifirst), "STEP": farray_numpy(step, ifirst),
    }
    scalars = {"I": I, "NNB": int(nlist[0]), "NNB0": int(old[0]), "NTOT": ntot, "N": n, "IFIRST": ifirst,
               "NNBMAX": int(c["nnbmax"]), "RS2": float(c["rs2"][r]), "SMIN": float(c["smin"]), "NBSMIN": 0}
    nnb0 = int(old[0])
    env = Machine(_statements(REGCOR), scalars, arrays).run()
    nnb, nbloss, nbgain = int(env["NNB"]), int(env["NBLOSS"]), int(env["NBGAIN"])
    return dict(nnb=nnb, members=[int(q) for q in nlist[1:1 + nnb]], nnb0=nnb0, nbloss=nbloss, nbgain=nbgain,
                lost=[int(q) for q in jjlist[:nbloss]], gained=[int(q) for q in jjlist[nnb0:nnb0 + nbgain]],
                freg=freg, fdr=fdr, dfirr=dfirr, dfd=dfd, nbsmin=int(env["NBSMIN"]))
