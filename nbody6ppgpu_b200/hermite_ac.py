"""Minimal Ahmad-Cohen Hermite driver around the regular-force C-ABI.

No Fortran compiler exists in this image, so the second half of BASELINE.json's metric ("wall s per N-body time unit")
and the north-star's energy-drift comparison cannot be produced with nbody6++ itself.  This driver restates the part of
the integrator that sits directly on the hot path, with block time steps, so that one N-body time unit can be integrated
with ANY library exporting the reference ABI (this repo's libgpunb_b200.so or the reference's own libraries), on
identical snapshots, and the energy errors laid side by side:

  * regular block   -- predict all particles, ``gpunb_send_`` of the predicted snapshot, ``gpunb_regf_`` on the block's
                       i-particles in chunks of 1024 with h2 = RS^2 and dtr = STEPR (intgrt.F:912-974,
                       util_gpu.F:33-60), overflow -> shrink RS and retry the chunk (util_gpu.F:71-97), self removed and
                       indices shifted by the caller (util_gpu.F:102-111);
  * irregular force -- fp64 sum over the returned neighbour list at the predicted positions (what nbint.f does);
  * corrector       -- 4th-order Hermite on the total force; the regular force is extrapolated linearly between regular
                       steps (intgrt.F:284-293), the regular polynomial of a regular step is taken over the OLD
                       neighbour list at both ends, i.e. corrected for the gained / lost neighbours
                       (regcor_gpu.F:510-552);
  * neighbour radius-- RS scaled towards NNBOPT members after every regular step (regcor_gpu.F:623-760, simplified to the
                       volume rule with a stabilising factor);
  * energy          -- kinetic energy + potential from ``gpupot_`` (energy.F / gpupot.gpu.cu).

Optional device paths (this repo's libraries only; the reference libraries run the host statements of the same steps):

  * ``irr_lib``     -- an ``irr.IrrLib`` (libirr_b200.so, or the reference's AVX library) takes the irregular force
                       (``irr_simd_firr_vec_``: the library predicts the neighbours itself, so the host predicts only the
                       active particles); the driver feeds it ``set_jp`` after every corrector and ``set_list`` after every
                       regular step, batched (intgrt.F:199-207,545);
  * ``use_regcor``  -- ``gpunb_b200_regcor_`` does the list bookkeeping of a regular block on the device: NLIST, and the force
                       swap DFIRR / DFD that turns the irregular force over the old list into the one over the new list
                       (regcor_gpu.F:267-470) -- the host neither diffs lists nor sums the changed members.

``csrc/ac_driver.cpp`` (libac_driver.so, ``ac_native.py``) is the C++ twin of this file, statement by statement: behind the
same libraries the two give the same integration bit for bit (tests/test_hermite_ac.py).  This file is the readable
statement and the place to change the algorithm; the twin is what ``bench.py --time-unit`` times.

What it deliberately leaves out: KS / chain regularisation, stellar evolution, external tides, the full RS control logic,
retention of small-step neighbours (SMIN is not used).  Host arithmetic is numpy fp64; the driver is a measurement harness,
not a production integrator.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

MAXTHR = 1024          # i-particles per gpunb_regf_ call (util_gpu.F:7)


@dataclass
class ACStats:
    t: float = 0.0
    irr_steps: int = 0
    reg_steps: int = 0
    reg_blocks: int = 0
    regf_calls: int = 0
    overflow_retries: int = 0
    wall_regf: float = 0.0
    wall_send: float = 0.0
    wall_total: float = 0.0
    wall_irr: float = 0.0
    wall_regcor: float = 0.0
    block_steps: int = 0
    energies: list = field(default_factory=list)      # (t, E)


def _pow2_floor(dt, dtmax):
    """Largest power-of-two fraction of dtmax not above dt (block time steps)."""
    e = np.floor(np.log2(np.maximum(dt, 1e-300) / dtmax))
    return dtmax * np.exp2(np.minimum(e, 0.0))


class AhmadCohen:
    def __init__(self, lib, m, x, v, *, nnbopt=40, lmax=128, eta_i=0.02, eta_r=0.02, dtmax=0.125, dtmin=2.0 ** -22,
                 m_flag=0, rs0=None, device_predictor=False, irr_lib=None, use_regcor=False):
        self.lib = lib
        self.irr = irr_lib
        self.use_regcor = bool(use_regcor)
        if self.use_regcor and not getattr(lib, "is_b200", False):
            raise ValueError("use_regcor needs libgpunb_b200.so (gpunb_b200_regcor_)")
        # device_predictor: regular blocks call gpunb_b200_predict_send_ (state kept on the device, updated with the
        # particles advanced since the last regular block) instead of uploading the host-predicted snapshot
        self.device_predictor = bool(device_predictor)
        if self.device_predictor and not getattr(lib, "is_b200", False):
            raise ValueError("device_predictor needs libgpunb_b200.so (gpunb_b200_state_all_/_update_/_predict_send_)")
        self._dirty = None
        self.n = n = m.shape[0]
        self.m = np.ascontiguousarray(m, dtype=np.float64)
        self.x0 = np.array(x, dtype=np.float64)
        self.v0 = np.array(v, dtype=np.float64)
        self.nnbopt, self.lmax, self.nnbmax = nnbopt, lmax, min(n // 2, lmax - 50) if lmax > 100 else lmax - 8
        self.eta_i, self.eta_r, self.dtmax, self.dtmin, self.m_flag = eta_i, eta_r, dtmax, dtmin, m_flag
        self.bodym = float(self.m.mean())
        r2 = (self.x0 ** 2).sum(1)
        if rs0 is None:          # initial guess: volume of nnbopt members at the half-mass density
            rs0 = (nnbopt / (0.5 * n) * 0.8 ** 3) ** (1.0 / 3.0)
        self.rs = rs0 * np.sqrt(1.0 + r2)              # fpoly0.F:53-56: RS grows outwards
        self.t = 0.0
        self.t0 = np.zeros(n); self.t0r = np.zeros(n)
        self.dt = np.full(n, dtmin); self.dtr = np.full(n, dtmax)
        self.fi = np.zeros((n, 3)); self.fid = np.zeros((n, 3))
        self.fr = np.zeros((n, 3)); self.frd = np.zeros((n, 3))
        self.f = np.zeros((n, 3)); self.fd = np.zeros((n, 3))      # total force / derivative at t0 (prediction)
        self.nb = np.full((n, self.nnbmax + 1), -1, dtype=np.int64)  # neighbour lists, -1 padded
        self.nnb = np.zeros(n, dtype=np.int64)
        self.stats = ACStats()
        lib.open(n + 10, 0)
        if self.irr is not None:
            self.lstride = 1 + 8 * ((self.nnbmax + 8) // 8) + 8
            self.irr.open(n, self.lstride, 0)
        self._initial_forces()

    def close(self):
        self.lib.close()
        if self.irr is not None:
            self.irr.close(0)

    # ---- irregular-force library feeds (intgrt.F:199-207: set_jp after the corrector, set_list after a regular step) ----
    def _irr_push_particles(self, idx):
        self.irr.set_jp_batch(idx.astype(np.int32) + 1, self.x0[idx], self.v0[idx], 0.5 * self.f[idx],
                              self.fd[idx] * (1.0 / 6.0), self.m[idx], self.t0[idx])

    def _irr_push_lists(self, idx):
        rows = np.zeros((idx.size, self.lstride), dtype=np.int32)
        rows[:, 0] = self.nnb[idx]
        k = self.nb.shape[1]
        rows[:, 1:1 + k] = np.where(self.nb[idx] >= 0, self.nb[idx] + 1, 0)
        self.irr.set_list_batch(idx.astype(np.int32) + 1, rows)

    def _predict_idx(self, idx, t):
        s = (t - self.t0[idx])[:, None]
        f2, fd6 = 0.5 * self.f[idx], self.fd[idx] * (1.0 / 6.0)
        xp = ((fd6 * s + f2) * s + self.v0[idx]) * s + self.x0[idx]
        vp = (fd6 * (1.5 * s) + f2) * (2.0 * s) + self.v0[idx]
        return xp, vp

    # ---- force pieces ------------------------------------------------------------------------
    def _predict(self, t):
        """xbpredall.f:17-26, in the integrator's conventions (F2 = F/2, FD6 = FDOT/6), fp64 without fusion."""
        s = (t - self.t0)[:, None]
        f2, fd6 = 0.5 * self.f, self.fd * (1.0 / 6.0)
        xp = ((fd6 * s + f2) * s + self.v0) * s + self.x0
        vp = (fd6 * (1.5 * s) + f2) * (2.0 * s) + self.v0
        return xp, vp

    def _push_state(self, idx=None):
        """Device-resident predictor: full state, or the particles just advanced (gpunb_b200_state_all_/_update_)."""
        if idx is None:
            self.lib.state_all(self.m, self.x0, self.v0, 0.5 * self.f, self.fd * (1.0 / 6.0), self.t0)
        else:
            self.lib.state_update(idx, self.m[idx], self.x0[idx], self.v0[idx], 0.5 * self.f[idx],
                                  self.fd[idx] * (1.0 / 6.0), self.t0[idx])

    def _irregular(self, idx, xp, vp, lists, counts):
        """fp64 force and derivative on particles idx from their neighbour lists (rows of -1 padded indices)."""
        k = int(counts.max()) if idx.size else 0
        fi = np.zeros((idx.size, 3)); fd = np.zeros((idx.size, 3))
        if k == 0:
            return fi, fd
        nbr = lists[:, :k]
        ok = nbr >= 0
        j = np.where(ok, nbr, 0)
        dx = xp[j] - xp[idx][:, None, :]
        dv = vp[j] - vp[idx][:, None, :]
        r2 = (dx * dx).sum(2)
        r2 = np.where(ok, r2, 1.0)
        rinv2 = 1.0 / r2
        mr3 = np.where(ok, self.m[j], 0.0) * rinv2 * np.sqrt(rinv2)
        rv = 3.0 * (dx * dv).sum(2) * rinv2
        fi = (mr3[:, :, None] * dx).sum(1)
        fd = (mr3[:, :, None] * (dv - rv[:, :, None] * dx)).sum(1)
        return fi, fd

    def _regular(self, idx, xp, vp, t=0.0, xi=None, vi=None, raw=False):
        """gpunb_send_ + gpunb_regf_ over the block idx; returns (fr, frd, lists[-1 padded], counts).
        xp, vp: predicted snapshot of all particles (None with the device predictor: nothing is uploaded); xi, vi: the
        block's own predicted positions when xp is None.  raw: also return the untouched gpunb_regf_ rows."""
        st = self.stats
        t0 = time.perf_counter()
        if self.device_predictor and self._dirty is not None:
            dirty = np.nonzero(self._dirty)[0]
            if dirty.size:
                self._push_state(dirty)
                self._dirty[:] = False
            self.lib.predict_send(self.n, t)
        else:
            self.lib.send(self.m, xp, vp)
        st.wall_send += time.perf_counter() - t0
        if xi is None:
            xi, vi = xp[idx], vp[idx]
        raw_rows = []
        nreg = idx.size
        fr = np.zeros((nreg, 3)); frd = np.zeros((nreg, 3))
        lists = np.full((nreg, self.nnbmax + 1), -1, dtype=np.int64)
        counts = np.zeros(nreg, dtype=np.int64)
        for c0 in range(0, nreg, MAXTHR):
            sel = idx[c0:c0 + MAXTHR]
            while True:
                h2 = self.rs[sel] ** 2 / (self.bodym if self.m_flag else 1.0)
                t0 = time.perf_counter()
                acc, jrk, pot, lst = self.lib.regf(h2, self.dtr[sel], xi[c0:c0 + MAXTHR], vi[c0:c0 + MAXTHR], self.lmax,
                                                   self.nnbmax, self.m_flag)
                st.wall_regf += time.perf_counter() - t0
                st.regf_calls += 1
                over = lst[:, 0] < 0
                if not over.any():
                    break
                # util_gpu.F:83-90 (NB_FLAG = 1): RS towards NNBOPT members
                cnt = -lst[over, 0].astype(np.float64)
                scale = np.where(cnt > self.nnbopt, (self.nnbopt / cnt) ** 0.333, (self.nnbopt / self.nnbmax) ** 0.4)
                self.rs[sel[over]] *= scale
                st.overflow_retries += 1
            fr[c0:c0 + sel.size] = acc; frd[c0:c0 + sel.size] = jrk
            if raw:                            # the rows go to gpunb_b200_regcor_ as they are
                raw_rows.append(lst.copy())
                continue
            # util_gpu.F:102-111: drop self, keep ascending order (all rows at once)
            body = lst[:, 1:]
            keep = (np.arange(body.shape[1])[None, :] < lst[:, :1]) & (body != sel[:, None])
            pos = np.cumsum(keep, axis=1) - 1
            rr, cc = np.nonzero(keep)
            lists[c0 + rr, pos[rr, cc]] = body[rr, cc]
            counts[c0:c0 + sel.size] = keep.sum(1)
        if raw:
            return fr, frd, np.concatenate(raw_rows), None
        return fr, frd, lists, counts

    def _regcor(self, idx, rows):
        """gpunb_b200_regcor_ on the rows gpunb_regf_ just returned for the block idx: (lists[-1 padded], counts, dfirr, dfd)."""
        t0 = time.perf_counter()
        old = np.zeros((idx.size, self.lmax), dtype=np.int32)
        old[:, 0] = self.nnb[idx]
        k = min(self.nb.shape[1], self.lmax - 1)
        old[:, 1:1 + k] = np.where(self.nb[idx, :k] >= 0, self.nb[idx, :k] + 1, 0)
        z = np.zeros((idx.size, 3))
        out = self.lib.regcor(idx.astype(np.int32) + 1, 1, self.n, self.n, rows, old, self.rs[idx] ** 2, None, 0.0,
                              self.nnbmax, z, z)
        nl = out["nlist"]
        counts = nl[:, 0].astype(np.int64)
        lists = np.full((idx.size, self.nnbmax + 1), -1, dtype=np.int64)
        kk = min(lists.shape[1], self.lmax - 1)
        ar = np.arange(kk)[None, :]
        lists[:, :kk] = np.where(ar < counts[:, None], nl[:, 1:1 + kk].astype(np.int64) - 1, -1)
        self.stats.wall_regcor += time.perf_counter() - t0
        return lists, counts, out["dfirr"], out["dfd"]

    def _firr(self, idx, t):
        t0 = time.perf_counter()
        acc, jrk, _ = self.irr.firr_vec(t, idx.astype(np.int32) + 1)
        self.stats.wall_irr += time.perf_counter() - t0
        return acc, jrk

    def _initial_forces(self):
        idx = np.arange(self.n)
        fr, frd, lists, counts = self._regular(idx, self.x0, self.v0)
        self.nb[:, :lists.shape[1]] = lists; self.nnb = counts
        if self.irr is not None:               # records without F / FDOT yet: at t = t0 the prediction does not use them
            self._irr_push_particles(idx); self._irr_push_lists(idx)
            fi, fid = self._firr(idx, 0.0)
        else:
            fi, fid = self._irregular(idx, self.x0, self.v0, lists, counts)
        self.fr, self.frd, self.fi, self.fid = fr, frd, fi, fid
        self.f = fi + fr; self.fd = fid + frd
        # starting steps from F and FDOT only (no higher derivatives yet)
        fa = np.linalg.norm(self.f, axis=1); fda = np.linalg.norm(self.fd, axis=1) + 1e-300
        fia = np.linalg.norm(fi, axis=1) + 1e-300; fida = np.linalg.norm(fid, axis=1) + 1e-300
        fra = np.linalg.norm(fr, axis=1) + 1e-300; frda = np.linalg.norm(frd, axis=1) + 1e-300
        dt_i = 0.1 * self.eta_i ** 0.5 * np.minimum(fa / fda, fia / fida) if True else None
        dt_r = 0.1 * self.eta_r ** 0.5 * fra / frda
        self.dt = np.maximum(_pow2_floor(dt_i, self.dtmax), self.dtmin)
        self.dtr = np.maximum(_pow2_floor(np.maximum(dt_r, self.dt), self.dtmax), self.dt)
        self._adjust_rs(idx, counts)
        if self.irr is not None:
            self._irr_push_particles(idx)
        if self.device_predictor:
            self._push_state()
            self._dirty = np.zeros(self.n, dtype=bool)

    def _adjust_rs(self, idx, counts):
        """Volume rule towards NNBOPT members with a stabilising factor (regcor_gpu.F:623-760, simplified)."""
        ratio = np.clip((self.nnbopt / np.maximum(counts, 1.0)) ** (1.0 / 3.0), 0.9, 1.1)
        self.rs[idx] *= ratio

    @staticmethod
    def _hermite_coeffs(f0, fd0, f1, fd1, dt):
        dt = dt[:, None]
        a2 = (-6.0 * (f0 - f1) - dt * (4.0 * fd0 + 2.0 * fd1)) / dt ** 2
        a3 = (12.0 * (f0 - f1) + 6.0 * dt * (fd0 + fd1)) / dt ** 3
        return a2, a3

    @staticmethod
    def _aarseth(eta, f1, fd1, a2, a3, dt):
        a2n = a2 + dt[:, None] * a3
        n = np.linalg.norm
        num = n(f1, axis=1) * n(a2n, axis=1) + n(fd1, axis=1) ** 2
        den = n(fd1, axis=1) * n(a3, axis=1) + n(a2n, axis=1) ** 2 + 1e-300
        return np.sqrt(eta * num / den)

    # ---- one block step ----------------------------------------------------------------------
    def step(self):
        st = self.stats
        tn = float((self.t0 + self.dt).min())
        act = np.nonzero(self.t0 + self.dt == tn)[0]
        use_irr = self.irr is not None
        st.block_steps += 1
        if use_irr:                                # the library predicts the neighbours itself: the host predicts the actives only
            xa, va = self._predict_idx(act, tn)
            xp = vp = None
        else:
            xp, vp = self._predict(tn)
            xa, va = xp[act], vp[act]
        isreg = self.t0r[act] + self.dtr[act] <= tn
        reg = act[isreg]
        dti = tn - self.t0[act]

        fr_new = self.fr[act] + self.frd[act] * (tn - self.t0r[act])[:, None]        # intgrt.F:284-293
        frd_new = self.frd[act].copy()
        lists = self.nb[act]; counts = self.nnb[act]
        fi_act = None
        if use_irr:                                # irregular force of every active particle over the list it holds (old lists)
            fi_act, fid_act = self._firr(act, tn)
        if reg.size:
            st.reg_blocks += 1; st.reg_steps += reg.size
            # regular polynomial over the OLD list at both ends (list changes corrected, regcor_gpu.F:510-552)
            if use_irr:
                fi_old, fid_old = fi_act[isreg], fid_act[isreg]
                if not self.device_predictor:
                    xp, vp = self._predict(tn)     # the snapshot gpunb_send_ uploads
            else:
                fi_old, fid_old = self._irregular(reg, xp, vp, self.nb[reg], self.nnb[reg])
            if self.use_regcor:
                # the device diffs the lists and returns the force swap: F_irr(new list) = F_irr(old list) + DFIRR
                frn, frdn, rows, _ = self._regular(reg, xp, vp, tn, xi=xa[isreg], vi=va[isreg], raw=True)
                lnew, cnew, dfi, dfd = self._regcor(reg, rows)
                fin, fidn = fi_old + dfi, fid_old + dfd
            else:
                frn, frdn, lnew, cnew = self._regular(reg, xp, vp, tn, xi=xa[isreg], vi=va[isreg])
                if use_irr:
                    self.nb[reg] = lnew; self.nnb[reg] = cnew
                    self._irr_push_lists(reg)
                    fin, fidn = self._firr(reg, tn)
                else:
                    fin, fidn = self._irregular(reg, xp, vp, lnew, cnew)
            ftot, fdtot = fin + frn, fidn + frdn
            fr_oldlist, frd_oldlist = ftot - fi_old, fdtot - fid_old
            dtr = tn - self.t0r[reg]
            a2r, a3r = self._hermite_coeffs(self.fr[reg], self.frd[reg], fr_oldlist, frd_oldlist, dtr)
            dtr_new = self._aarseth(self.eta_r, fr_oldlist, frd_oldlist, a2r, a3r, dtr)
            fr_new[isreg] = frn; frd_new[isreg] = frdn
            lists = lists.copy(); counts = counts.copy()
            lists[isreg] = lnew; counts[isreg] = cnew
        if use_irr or self.use_regcor:
            if fi_act is None:
                fi_act, fid_act = self._irregular(act, xp, vp, self.nb[act], self.nnb[act])
            fi_new, fid_new = fi_act.copy(), fid_act.copy()
            if reg.size:
                fi_new[isreg] = fin; fid_new[isreg] = fidn
        else:
            fi_new, fid_new = self._irregular(act, xp, vp, lists, counts)
        f1, fd1 = fi_new + fr_new, fid_new + frd_new
        a2, a3 = self._hermite_coeffs(self.f[act], self.fd[act], f1, fd1, dti)
        d = dti[:, None]
        self.x0[act] = xa + d ** 4 / 24.0 * a2 + d ** 5 / 120.0 * a3
        self.v0[act] = va + d ** 3 / 6.0 * a2 + d ** 4 / 24.0 * a3
        self.t0[act] = tn
        self.fi[act], self.fid[act] = fi_new, fid_new
        self.f[act], self.fd[act] = f1, fd1
        st.irr_steps += act.size
        if self._dirty is not None:
            self._dirty[act] = True

        # new irregular steps: block quantised, at most doubled, doubled only on even block boundaries
        dt_new = self._aarseth(self.eta_i, f1, fd1, a2, a3, dti)
        old = self.dt[act]
        q = np.where(dt_new < old, np.maximum(_pow2_floor(dt_new, self.dtmax), self.dtmin), old)
        can_double = (dt_new >= 2.0 * old) & (np.mod(tn, 2.0 * old) == 0.0) & (2.0 * old <= self.dtmax)
        q = np.where(can_double, 2.0 * old, q)
        self.dt[act] = q
        if reg.size:
            self.fr[reg], self.frd[reg] = frn, frdn
            self.t0r[reg] = tn
            self.nb[reg] = lnew; self.nnb[reg] = cnew
            oldr = self.dtr[reg]
            qr = np.where(dtr_new < oldr, _pow2_floor(dtr_new, self.dtmax), oldr)
            dbl = (dtr_new >= 2.0 * oldr) & (np.mod(tn, 2.0 * oldr) == 0.0) & (2.0 * oldr <= self.dtmax)
            qr = np.where(dbl, 2.0 * oldr, qr)
            self.dtr[reg] = np.maximum(qr, self.dt[reg])
            self._adjust_rs(reg, cnew)
            if use_irr and self.use_regcor:            # the host-list branch above has pushed the new lists already
                self._irr_push_lists(reg)
        else:
            # the regular force of non-regular actives stays a linear extrapolation from t0r
            pass
        # an irregular step never exceeds the distance to the particle's next regular time
        nxt = self.t0r[act] + self.dtr[act] - tn
        self.dt[act] = np.where(nxt > 0, np.minimum(self.dt[act], _pow2_floor(nxt, self.dtmax)), self.dt[act])
        if use_irr:
            self._irr_push_particles(act)
        self.t = tn
        st.t = tn

    def energy(self):
        """E = T - U with U from gpupot_ (positive sum m/r per particle); all particles must be synchronised."""
        assert np.all(self.t0 == self.t), "energy() needs a synchronised system (t multiple of dtmax)"
        phi = self.lib.gpupot(1, self.n, self.m, self.x0)
        return float(0.5 * (self.m * (self.v0 ** 2).sum(1)).sum() - 0.5 * (self.m * phi).sum())

    def run(self, t_end, energy_every=None):
        st = self.stats
        w0 = time.perf_counter()
        if not st.energies:
            st.energies.append((self.t, self.energy()))
        next_e = self.t + energy_every if energy_every else None
        while self.t < t_end:
            self.step()
            if next_e is not None and self.t >= next_e and np.all(self.t0 == self.t):
                st.energies.append((self.t, self.energy()))
                next_e += energy_every
        if st.energies[-1][0] != self.t:
            st.energies.append((self.t, self.energy()))
        st.wall_total += time.perf_counter() - w0
        return st


def energy_drift(lib, n=1024, seed=5, t_end=1.0, imf="equal", **kw):
    """Integrate a Plummer sphere for t_end N-body time units behind `lib`; returns (relative energy error, stats)."""
    from . import snapshots as S
    m, x, v = S.plummer(n, seed, imf)
    ac = AhmadCohen(lib, m, x, v, **kw)
    try:
        st = ac.run(t_end)
    finally:
        ac.close()
    e0, e1 = st.energies[0][1], st.energies[-1][1]
    return (e1 - e0) / abs(e0), st
