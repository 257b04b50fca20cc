"""Test-side loader of the CHECKERS under oracle/ (never imported by the product package)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


def build():
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR)], check=True, capture_output=True)


def _d(a):
    return a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self):
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            build()
        self.lib = C.CDLL(str(so))
        self.lib.oracle_regf_f32.restype = None
        self.lib.oracle_regf_f64.restype = None
        self.lib.oracle_regf_f64_given_list.restype = None
        self.lib.oracle_pot_f64.restype = None

    @staticmethod
    def _prep(m, x, v, h2, dtr, xi, vi):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        return f(m), f(x), f(v), f(h2), f(dtr), f(xi), f(vi)

    def regf_f32(self, m, x, v, h2, dtr, xi, vi, lmax, nnbmax, m_flag=0):
        m, x, v, h2, dtr, xi, vi = self._prep(m, x, v, h2, dtr, xi, vi)
        ni, nj = h2.shape[0], m.shape[0]
        acc = np.zeros((ni, 3)); jrk = np.zeros((ni, 3)); pot = np.zeros(ni)
        lst = np.zeros((ni, lmax), dtype=np.int32)
        self.lib.oracle_regf_f32(C.c_int(ni), C.c_int(nj), _d(m), _d(x), _d(v), _d(h2), _d(dtr), _d(xi), _d(vi),
                                 _d(acc), _d(jrk), _d(pot), C.c_int(lmax), C.c_int(nnbmax),
                                 lst.ctypes.data_as(_ip), C.c_int(m_flag))
        return acc, jrk, pot, lst

    def regf_f64(self, m, x, v, h2, dtr, xi, vi, lmax, nnbmax, m_flag=0, band_k=4.0):
        """fp64 force with the reference FP32 membership; also per-row min ulp distance to the RS boundary."""
        m, x, v, h2, dtr, xi, vi = self._prep(m, x, v, h2, dtr, xi, vi)
        ni, nj = h2.shape[0], m.shape[0]
        acc = np.zeros((ni, 3)); jrk = np.zeros((ni, 3)); pot = np.zeros(ni)
        lst = np.zeros((ni, lmax), dtype=np.int32)
        band = np.zeros(ni, dtype=np.float32)
        self.scale = np.zeros((ni, 2))
        nband = C.c_longlong(0)
        self.lib.oracle_regf_f64(C.c_int(ni), C.c_int(nj), _d(m), _d(x), _d(v), _d(h2), _d(dtr), _d(xi), _d(vi),
                                 _d(acc), _d(jrk), _d(pot), C.c_int(lmax), C.c_int(nnbmax),
                                 lst.ctypes.data_as(_ip), C.c_int(m_flag),
                                 band.ctypes.data_as(_fp), C.c_float(band_k), C.byref(nband), _d(self.scale))
        return acc, jrk, pot, lst, band, int(nband.value)

    def regf_f64_given_list(self, m, x, v, xi, vi, lst):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        m, x, v, xi, vi = f(m), f(x), f(v), f(xi), f(vi)
        lst = np.ascontiguousarray(lst, dtype=np.int32)
        ni, nj, lmax = xi.shape[0], m.shape[0], lst.shape[1]
        acc = np.zeros((ni, 3)); jrk = np.zeros((ni, 3)); pot = np.zeros(ni)
        self.scale = np.zeros((ni, 2))
        self.lib.oracle_regf_f64_given_list(C.c_int(ni), C.c_int(nj), _d(m), _d(x), _d(v), _d(xi), _d(vi),
                                            _d(acc), _d(jrk), _d(pot), C.c_int(lmax), lst.ctypes.data_as(_ip), _d(self.scale))
        return acc, jrk, pot

    def regcor(self, index_i, ifirst, n, ntot, new_list, old_list, m, x, v, rs2, step, smin, nnbmax, freg, fdr,
               dfirr=None, dfd=None):
        """oracle/regcor_oracle.c: util_gpu.F:102-111 + regcor_gpu.F:267-470 for a batch of rows; same argument meaning
        and return dict as ForceLib.regcor (the snapshot is passed explicitly)."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        index_i = np.ascontiguousarray(index_i, dtype=np.int32); ni = index_i.shape[0]
        nl = np.array(new_list, dtype=np.int32, order="C"); lmax = nl.shape[1]
        ol = np.ascontiguousarray(old_list, dtype=np.int32)
        m, x, v, rs2 = f(m), f(x), f(v), f(rs2)
        st = None if step is None else f(step)
        fr = np.array(freg, dtype=np.float64, order="C"); fd = np.array(fdr, dtype=np.float64, order="C")
        di = np.zeros((ni, 3)) if dfirr is None else np.array(dfirr, dtype=np.float64, order="C")
        dd = np.zeros((ni, 3)) if dfd is None else np.array(dfd, dtype=np.float64, order="C")
        nbloss = np.zeros(ni, dtype=np.int32); nbgain = np.zeros(ni, dtype=np.int32)
        jj = np.zeros((ni, 2 * lmax), dtype=np.int32)
        nbsmin = C.c_int(0)
        ip = lambda a: a.ctypes.data_as(_ip)
        self.lib.oracle_regcor.restype = None
        import time as _t
        _t0 = _t.perf_counter()
        self.lib.oracle_regcor(C.c_int(ni), ip(index_i), C.c_int(ifirst), C.c_int(n), C.c_int(ntot), C.c_int(lmax), ip(nl),
                               ip(ol), _d(m), _d(x), _d(v), _d(rs2), None if st is None else _d(st), C.c_double(smin),
                               C.c_int(nnbmax), _d(fr), _d(fd), _d(di), _d(dd), ip(nbloss), ip(nbgain), ip(jj),
                               C.byref(nbsmin))
        self.last_regcor_s = _t.perf_counter() - _t0          # the C call alone
        return dict(nlist=nl, nbloss=nbloss, nbgain=nbgain, jjlist=jj, freg=fr, fdr=fd, dfirr=di, dfd=dd,
                    nbsmin=int(nbsmin.value))

    def pot_f64(self, istart, ni, m, x):
        m = np.ascontiguousarray(m, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
        pot = np.zeros(ni)
        self.lib.oracle_pot_f64(C.c_int(istart), C.c_int(ni), C.c_int(m.shape[0]), _d(m), _d(x), _d(pot))
        return pot


_REF = {}


def ref_avx():
    """oracle/_ref/libgpunb_ref_avx.so through the same ctypes mirror the product uses."""
    from nbody6ppgpu_b200.gpunb import ForceLib
    if "avx" not in _REF:
        so = ORACLE_DIR / "_ref" / "libgpunb_ref_avx.so"
        if not so.exists():
            build()
        _REF["avx"] = ForceLib(so)
    return _REF["avx"]


def ref_gpu():
    from nbody6ppgpu_b200.gpunb import ForceLib
    if "gpu" not in _REF:
        _REF["gpu"] = ForceLib(ORACLE_DIR / "_ref" / "libgpunb_ref_gpu.so")
    return _REF["gpu"]


def relerr(a, b):
    """max over rows of |a-b| / |b| (vector norm per particle)."""
    a = np.asarray(a); b = np.asarray(b)
    if a.ndim == 1:
        return float(np.max(np.abs(a - b) / np.abs(b)))
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)))


def relerr_scaled(a, b, scale):
    """max over rows of |a-b| / max(|b|, scale): `scale` is the quadrature sum of the row's pair terms
    (Oracle.scale), the natural magnitude of a random-walk sum.  Equals relerr() except for rows whose
    fp64 sum cancels below that magnitude, where no FP32 pair arithmetic can hold 1e-6 of |b|."""
    d = np.linalg.norm(np.asarray(a) - np.asarray(b), axis=1)
    return float(np.max(d / np.maximum(np.linalg.norm(b, axis=1), scale)))


def list_rows_equal(la, lb):
    """Rows equal in count and in entries 1..count (entries past count are unspecified)."""
    bad = []
    for i in range(la.shape[0]):
        na, nb = int(la[i, 0]), int(lb[i, 0])
        if na != nb or (na > 0 and not np.array_equal(la[i, 1:1 + na], lb[i, 1:1 + nb])):
            bad.append(i)
    return bad
