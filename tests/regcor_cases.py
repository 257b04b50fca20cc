"""Seeded inputs for the neighbour-list bookkeeping tests (gpunb_b200_regcor_ / oracle/regcor_oracle.c) and two independent
statements of what the Fortran computes:

* ``fortran_walk``  -- a literal transcription of regcor_gpu.F:267-470 (1-based arrays, the GO TO structure kept as a label
                       state machine, sentinel NTOT+1 and all), deliberately NOT sharing structure with the C restatement;
* ``sets_and_sums`` -- lost / gained as Python set differences and the pair terms in vectorised numpy.
"""
from __future__ import annotations

import numpy as np


def make_case(n_tot=3000, ni=257, ifirst=7, n_cm=40, lmax=128, nnb_mean=30.0, seed=3, overflow_rows=(), empty_old_rows=(),
              drift=0.02, smin_frac=0.15):
    """Snapshot of NTOT - IFIRST + 1 = n_tot particles (the last n_cm are c.m. bodies, numbers > N); rows of ni particles:
    new rows in gpunb_regf_ format (0-based j, self included) from neighbour spheres at the current positions, old lists
    (Fortran numbers, self excluded) from slightly different spheres at drifted positions."""
    from scipy.spatial import cKDTree
    from nbody6ppgpu_b200 import snapshots as S
    rng = np.random.default_rng(seed)
    m, x, v = S.plummer(n_tot, seed, "kroupa")
    ntot = ifirst + n_tot - 1
    n = ntot - n_cm
    rows_j = np.sort(rng.choice(n_tot, ni, replace=False))                   # snapshot index of each row's particle
    index_i = (rows_j + ifirst).astype(np.int32)
    tree = cKDTree(x)
    d, _ = tree.query(x[rows_j], k=int(nnb_mean) + 1)
    rs = d[:, -1] * rng.uniform(0.8, 1.2, ni)
    x_old = x + drift * rs.mean() * rng.normal(size=x.shape)
    tree_old = cKDTree(x_old)
    new = np.zeros((ni, lmax), dtype=np.int32)
    old = np.zeros((ni, lmax), dtype=np.int32)
    for r in range(ni):
        nb = np.sort(np.asarray(tree.query_ball_point(x[rows_j[r]], rs[r]), dtype=np.int64))[: lmax - 12]
        if rows_j[r] not in nb:
            nb = np.sort(np.append(nb, rows_j[r]))
        new[r, 0] = nb.size; new[r, 1:1 + nb.size] = nb                      # 0-based j, self included
        ob = np.sort(np.asarray(tree_old.query_ball_point(x_old[rows_j[r]], rs[r] * rng.uniform(0.9, 1.1)), dtype=np.int64))
        ob = ob[ob != rows_j[r]][: lmax - 12] + ifirst
        if r in empty_old_rows:
            ob = ob[:0]
        old[r, 0] = ob.size; old[r, 1:1 + ob.size] = ob
    for r in overflow_rows:
        new[r, 0] = -(lmax + 5)
    step = 2.0 ** -rng.integers(3, 12, size=n_tot).astype(np.float64)
    smin = float(np.quantile(step, smin_frac))
    step[rng.random(n_tot) < 0.02] = smin                                      # STEP == SMIN: retained but never sets JMIN
    freg = rng.normal(size=(ni, 3)); fdr = rng.normal(size=(ni, 3))
    return dict(m=m, x=x, v=v, index_i=index_i, ifirst=ifirst, n=n, ntot=ntot, lmax=lmax, nnbmax=lmax - 50 if lmax > 100 else lmax - 8,
                new=new, old=old, rs2=rs ** 2, step=step, smin=smin, freg=freg, fdr=fdr)


def make_edge_case(**kw):
    """make_case plus hand-built rows: only self in the new row (everything lost), both lists empty, identical lists, a new
    row without self, the longest lists lmax allows."""
    c = make_case(**kw)
    ifirst, lmax = c["ifirst"], c["lmax"]
    me = lambda r: int(c["index_i"][r]) - ifirst
    new, old = c["new"], c["old"]
    new[1, :] = 0; new[1, 0] = 1; new[1, 1] = me(1)                          # only self: NNB = 0, every old member is lost
    new[2, :] = 0; old[2, :] = 0                                            # nothing at all
    keep = [int(q) - ifirst for q in old[3, 1:1 + old[3, 0]]]
    row = sorted(set(keep + [me(3)]))
    new[3, :] = 0; new[3, 0] = len(row); new[3, 1:1 + len(row)] = row        # the old members + self: no change
    row = [int(q) for q in new[4, 1:1 + new[4, 0]] if int(q) != me(4)]
    new[4, :] = 0; new[4, 0] = len(row); new[4, 1:1 + len(row)] = row        # a row without self (the caller passed i not in j)
    cap = lmax - 3
    allj = [j for j in range(c["m"].shape[0]) if j != me(6)]
    o = allj[:cap]
    old[6, :] = 0; old[6, 0] = len(o); old[6, 1:1 + len(o)] = np.asarray(o) + ifirst
    nw = sorted(allj[40:40 + cap - 1] + [me(6)])
    new[6, :] = 0; new[6, 0] = len(nw); new[6, 1:1 + len(nw)] = nw           # the longest rows lmax allows, shifted by 40 members
    return c


def make_random_case(seed=77):
    """400 random rows over a universe of 60 particles: short lists, many empty or nearly empty ones, c.m. particles, NNB near
    NNBMAX -- the corners of the walk and of the retention loop that a physical snapshot rarely visits."""
    rng = np.random.default_rng(seed)
    n_tot, ifirst, lmax = 60, 4, 24
    m = rng.uniform(0.5, 1.5, n_tot); x = rng.normal(size=(n_tot, 3)); v = rng.normal(size=(n_tot, 3))
    ni = 400
    rows_j = rng.integers(0, n_tot, ni)
    new = np.zeros((ni, lmax), dtype=np.int32); old = np.zeros((ni, lmax), dtype=np.int32)
    for r in range(ni):
        k_new, k_old = int(rng.integers(0, 13)), int(rng.integers(0, 13))
        others = np.setdiff1d(np.arange(n_tot), [rows_j[r]])
        nb = np.sort(rng.choice(others, k_new, replace=False))
        if rng.random() < 0.8:
            nb = np.sort(np.append(nb, rows_j[r]))                   # self included, as gpunb_regf_ returns it
        new[r, 0] = nb.size; new[r, 1:1 + nb.size] = nb
        ob = np.sort(rng.choice(others, k_old, replace=False)) + ifirst
        old[r, 0] = ob.size; old[r, 1:1 + ob.size] = ob
    step = 2.0 ** -rng.integers(1, 6, size=n_tot).astype(np.float64)
    c = dict(m=m, x=x, v=v, index_i=(rows_j + ifirst).astype(np.int32), ifirst=ifirst, n=ifirst + n_tot - 1 - 8, ntot=ifirst + n_tot - 1,
             lmax=lmax, nnbmax=9, new=new, old=old, rs2=rng.uniform(0.2, 6.0, ni), step=step, smin=0.125,
             freg=rng.normal(size=(ni, 3)), fdr=rng.normal(size=(ni, 3)))
    return c


def _pair(xi, vi, xj, vj, mj):
    a = xj - xi; dv = vj - vi
    rij2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2]
    dr2i = 1.0 / rij2
    dr3i = mj * dr2i * np.sqrt(dr2i)
    drdv = a[0] * dv[0] + a[1] * dv[1] + a[2] * dv[2]
    drdp = 3.0 * drdv * dr2i
    return a * dr3i, (dv - a * drdp) * dr3i


def fortran_walk(c, r, use_step=True):
    """Row r the way regcor_gpu.F states it (labels 50-70 as a state machine).  Returns dict like ForceLib.regcor for one row."""
    I, ifirst, N, NTOT = int(c["index_i"][r]), c["ifirst"], c["n"], c["ntot"]
    X = lambda J: c["x"][J - ifirst]
    XDOT = lambda J: c["v"][J - ifirst]
    BODY = lambda J: c["m"][J - ifirst]
    STEP = (lambda J: c["step"][J - ifirst]) if use_step else (lambda J: np.inf)
    SMIN, NNBMAX = c["smin"], c["nnbmax"]
    XI, XIDOT, RS2 = X(I), XDOT(I), c["rs2"][r]
    lmax = c["lmax"]
    # util_gpu.F:102-111
    row = c["new"][r]
    NLIST = [0] * (lmax + 3)                       # 1-based: NLIST[1] = count
    L1 = 1
    for LL in range(2, row[0] + 2):
        ITEMP = int(row[LL - 1]) + ifirst
        if ITEMP != I:
            L1 += 1
            NLIST[L1] = ITEMP
    NLIST[1] = L1 - 1
    NNB = NLIST[1]
    LIST = [0] + [int(q) for q in c["old"][r]] + [0, 0]      # 1-based: LIST[1] = NNB0
    NNB0 = LIST[1]
    JJLIST = [0] * (2 * lmax + 2)
    FREG = c["freg"][r].copy(); FDR = c["fdr"][r].copy()
    DFIRR = np.zeros(3); DFD = np.zeros(3)
    NBLOSS = NBGAIN = 0
    NBSMIN = 0
    label = 50
    if NNB0 == 0:
        NBGAIN = NNB
        for L in range(1, NNB + 1):
            JJLIST[L] = NLIST[L + 1]
        label = 70
    if label != 70:
        JMIN = 0
        L = 2; LG = 2
        NLIST[NNB + 2] = NTOT + 1
        NLIST[1] = LIST[NNB0 + 1]
        label = 56
        while True:
            if label == 56:
                if LIST[L] == NLIST[LG]:
                    label = 58; continue
                if LIST[L] >= NLIST[LG]:
                    NBGAIN += 1
                    JJLIST[NNB0 + NBGAIN] = NLIST[LG]
                    L -= 1
                else:
                    NBLOSS += 1
                    J = LIST[L]
                    JJLIST[NBLOSS] = J
                    if STEP(J) < SMIN:
                        JMIN = J
                    LG -= 1
                label = 58
            if label == 58:
                if L <= NNB0:
                    L += 1; LG += 1
                    label = 56; continue
                elif LG <= NNB:
                    LG += 1
                    LIST[L] = NTOT + 1
                    label = 56; continue
                break
        if JMIN != 0:
            K = 1
            label = 60
            while True:
                if label == 60:
                    if NNB > NNBMAX or I > N:
                        break
                    J = JJLIST[K]
                    if STEP(J) > SMIN or J < ifirst or J > N:
                        label = 68
                    else:
                        RIJ2 = (XI[0] - X(J)[0]) ** 2 + (XI[1] - X(J)[1]) ** 2 + (XI[2] - X(J)[2]) ** 2
                        if RIJ2 > 4.0 * RS2:
                            label = 68
                        else:
                            L = NNB + 1
                            while True:                        # 62
                                if NLIST[L] < J:
                                    break
                                NLIST[L + 1] = NLIST[L]
                                L -= 1
                                if not L > 1:
                                    break
                            NLIST[L + 1] = J                    # 64
                            NNB += 1
                            NBLOSS -= 1
                            LIST[NNB0 + 1] = NLIST[1]
                            NBSMIN += 1
                            f, fd = _pair(XI, XIDOT, X(J), XDOT(J), BODY(J))
                            FREG = FREG - f; FDR = FDR - fd
                            if K > NBLOSS:
                                break
                            for L in range(K, NBLOSS + 1):
                                JJLIST[L] = JJLIST[L + 1]
                            K -= 1
                            label = 68
                if label == 68:
                    K += 1
                    if K <= NBLOSS:
                        label = 60; continue
                    break
    for L in range(1, NBLOSS + 1):
        J = JJLIST[L]
        f, fd = _pair(XI, XIDOT, X(J), XDOT(J), BODY(J))
        DFIRR = DFIRR - f; DFD = DFD - fd
    for L in range(1, NBGAIN + 1):
        J = JJLIST[NNB0 + L]
        f, fd = _pair(XI, XIDOT, X(J), XDOT(J), BODY(J))
        DFIRR = DFIRR + f; DFD = DFD + fd
    return dict(nnb=NNB, members=NLIST[2:NNB + 2], nbloss=NBLOSS, nbgain=NBGAIN, lost=JJLIST[1:NBLOSS + 1],
                gained=JJLIST[NNB0 + 1:NNB0 + NBGAIN + 1], freg=FREG, fdr=FDR, dfirr=DFIRR, dfd=DFD, nbsmin=NBSMIN, nnb0=NNB0)


def sets_and_sums(c, r):
    """No retention: lost / gained as set differences, DFIRR / DFD as vectorised sums (order-free, so only ~1e-13 exact)."""
    I, ifirst = int(c["index_i"][r]), c["ifirst"]
    row = c["new"][r]
    new = [int(q) + ifirst for q in row[1:1 + row[0]] if int(q) + ifirst != I]
    old = [int(q) for q in c["old"][r][1:1 + c["old"][r][0]]]
    lost = sorted(set(old) - set(new)); gained = sorted(set(new) - set(old))
    xi, vi = c["x"][I - ifirst], c["v"][I - ifirst]

    def total(js):
        if not js:
            return np.zeros(3), np.zeros(3)
        j = np.asarray(js) - ifirst
        a = c["x"][j] - xi; dv = c["v"][j] - vi
        r2 = (a * a).sum(1)
        w = c["m"][j] / (r2 * np.sqrt(r2))
        return (a * w[:, None]).sum(0), ((dv - a * (3.0 * (a * dv).sum(1) / r2)[:, None]) * w[:, None]).sum(0)
    fl, dl = total(lost); fg, dg = total(gained)
    return dict(members=new, lost=lost, gained=gained, dfirr=fg - fl, dfd=dg - dl)


def compare_rows(out, c, rows, walk):
    """out: dict of ForceLib.regcor / Oracle.regcor; walk(c, r) -> fortran_walk-like dict.  Integer results must be equal,
    fp64 results bit for bit.  Returns the total number of retained members seen."""
    lmax = c["lmax"]
    total = 0
    for r in rows:
        w = walk(c, r)
        nl = out["nlist"][r]
        assert nl[0] == w["nnb"] and list(nl[1:1 + nl[0]]) == list(w["members"]), ("NLIST", r)
        assert out["nbloss"][r] == w["nbloss"] and out["nbgain"][r] == w["nbgain"], ("counts", r)
        jj = out["jjlist"][r]
        assert list(jj[:w["nbloss"]]) == list(w["lost"]), ("lost", r)
        assert list(jj[w["nnb0"]:w["nnb0"] + w["nbgain"]]) == list(w["gained"]), ("gained", r)
        for k in ("freg", "fdr", "dfirr", "dfd"):
            assert np.array_equal(out[k][r], w[k]), (k, r, out[k][r], w[k])
        total += w["nbsmin"]
    return total
