"""ctypes mirror of the irregular-force C-ABI and its fp64 statement.

``IrrLib(path)`` binds any shared library exporting the reference's six Fortran-callable symbols
(src/Main/irr.avx.cpp:565-603): ``irr_simd_open_ / close_ / profile_ / set_jp_ / set_list_ / firr_vec_`` --

* ``oracle/_ref/libirr_ref_avx.so``          the reference's own AVX library (tests only),
* ``nbody6ppgpu_b200/libirr_b200.so``         this repo's CUDA library (DRAFT: compiled and linked, not yet validated
                                              on a GPU -- the round's GPU budget was spent on the regular-force path).

Argument meaning follows the reference: particle addresses are 1-based; ``set_jp`` stores X0, X0DOT, F/2, FDOT/6, BODY,
T0 of one particle (irr.avx.cpp:437-447); ``set_list`` takes the NBODY6 list ``[nnb, j1, j2, ...]`` with 1-based
neighbour addresses (:449-494); ``firr_vec(ti, addr[ni])`` predicts every involved particle to time ``ti`` and returns the
force and its derivative over each particle's list, plus the address of its nearest neighbour (:496-563).

``firr_f64`` is the fp64 statement of the same sum (what nbint.f computes); the tests check every library against it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def lib_path() -> Path:
    return _HERE / "libirr_b200.so"


class IrrLib:
    def __init__(self, path):
        path = Path(path)
        if not path.exists():
            raise RuntimeError(f"{path} not found -- build it first")
        self.lib = L = C.CDLL(str(path))
        L.irr_simd_open_.argtypes = [_ip, _ip, _ip]
        L.irr_simd_close_.argtypes = [_ip]
        L.irr_simd_profile_.argtypes = [_ip]
        L.irr_simd_set_jp_.argtypes = [_ip, _dp, _dp, _dp, _dp, _dp, _dp]
        L.irr_simd_set_list_.argtypes = [_ip, _ip]
        L.irr_simd_firr_vec_.argtypes = [_dp, _ip, _ip, _dp, _dp, _ip]
        for f in (L.irr_simd_open_, L.irr_simd_close_, L.irr_simd_profile_, L.irr_simd_set_jp_, L.irr_simd_set_list_,
                  L.irr_simd_firr_vec_):
            f.restype = None

    def open(self, nmax: int, lmax: int, rank: int = 0):
        self.lib.irr_simd_open_(C.byref(C.c_int(nmax)), C.byref(C.c_int(lmax)), C.byref(C.c_int(rank)))

    def close(self, rank: int = 0):
        self.lib.irr_simd_close_(C.byref(C.c_int(rank)))

    def profile(self, rank: int = 0):
        self.lib.irr_simd_profile_(C.byref(C.c_int(rank)))

    def set_jp(self, addr: int, pos, vel, acc2, jrk6, mass: float, time: float):
        a = [np.ascontiguousarray(q, dtype=np.float64) for q in (pos, vel, acc2, jrk6)]
        self.lib.irr_simd_set_jp_(C.byref(C.c_int(addr)), *[q.ctypes.data_as(_dp) for q in a],
                                  C.byref(C.c_double(mass)), C.byref(C.c_double(time)))

    def set_list(self, addr: int, nblist):
        """nblist = [nnb, j1, ..., j_nnb] (1-based), padded by the caller to a multiple of 8 entries behind the count
        (the reference reads the list 8 at a time, irr.avx.cpp:464-487)."""
        nb = np.ascontiguousarray(nblist, dtype=np.int32)
        self.lib.irr_simd_set_list_(C.byref(C.c_int(addr)), nb.ctypes.data_as(_ip))

    def firr_vec(self, ti: float, addr):
        addr = np.ascontiguousarray(addr, dtype=np.int32)
        ni = addr.shape[0]
        acc = np.zeros((ni, 3)); jrk = np.zeros((ni, 3)); nnbid = np.zeros(ni, dtype=np.int32)
        self.lib.irr_simd_firr_vec_(C.byref(C.c_double(ti)), C.byref(C.c_int(ni)), addr.ctypes.data_as(_ip),
                                    acc.ctypes.data_as(_dp), jrk.ctypes.data_as(_dp), nnbid.ctypes.data_as(_ip))
        return acc, jrk, nnbid


def pad_list(neigh_1based) -> np.ndarray:
    """[nnb, j...] with room for the reference's 8-wide reads."""
    nnb = len(neigh_1based)
    out = np.zeros(1 + 8 * ((nnb + 7) // 8) + 8, dtype=np.int32)
    out[0] = nnb
    out[1:1 + nnb] = neigh_1based
    return out


def firr_f64(ti, addr, lists, x0, v0, a2, j6, m, t0):
    """fp64 statement: predict (irr.avx.cpp:143-157: pos = x0 + s (v0 + s (a2 + s j6)), vel = v0 + 2 s (a2 + 1.5 s j6),
    s = ti - t0), then force / derivative over each list (:319-353) and the nearest neighbour (:339-342).
    addr and list entries are 1-based; arrays are indexed by address - 1."""
    s = (ti - t0)[:, None]
    xp = x0 + s * (v0 + s * (a2 + s * j6))
    vp = v0 + 2.0 * s * (a2 + 1.5 * s * j6)
    ni = len(addr)
    acc = np.zeros((ni, 3)); jrk = np.zeros((ni, 3)); nnbid = np.zeros(ni, dtype=np.int32)
    for k, a in enumerate(addr):
        i = a - 1
        nb = np.asarray(lists[i], dtype=np.int64) - 1
        if nb.size == 0:
            continue
        dx = xp[nb] - xp[i]
        dv = vp[nb] - vp[i]
        r2 = (dx * dx).sum(1)
        rv = (dx * dv).sum(1)
        rinv2 = 1.0 / r2
        mr3 = m[nb] * rinv2 * np.sqrt(rinv2)
        acc[k] = (mr3[:, None] * dx).sum(0)
        jrk[k] = (mr3[:, None] * (dv - 3.0 * (rv * rinv2)[:, None] * dx)).sum(0)
        nnbid[k] = nb[np.argmin(r2)] + 1
    return acc, jrk, nnbid
