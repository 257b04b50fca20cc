"""ctypes mirror of the irregular-force C-ABI and its fp64 statement.

``IrrLib(path)`` binds any shared library exporting the reference's six Fortran-callable symbols
(src/Main/irr.avx.cpp:565-603): ``irr_simd_open_ / close_ / profile_ / set_jp_ / set_list_ / firr_vec_`` --

* ``oracle/_ref/libirr_ref_avx.so``          the reference's own AVX library (tests only),
* ``nbody6ppgpu_b200/libirr_b200.so``         this repo's CUDA library.

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
        self.is_b200 = hasattr(L, "irr_b200_version")
        if self.is_b200:
            L.irr_b200_set_jp_batch_.argtypes = [_ip, _ip] + [_dp] * 6
            L.irr_b200_set_list_batch_.argtypes = [_ip, _ip, _ip, _ip]
            L.irr_b200_counters.argtypes = [_dp]
            L.irr_b200_particle_records_.argtypes = [_ip]
            L.irr_b200_particle_records_.restype = C.c_void_p
            L.irr_b200_flush_.argtypes = []
            L.irr_b200_flush_.restype = None
            for f in (L.irr_b200_set_jp_batch_, L.irr_b200_set_list_batch_, L.irr_b200_counters):
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


    # batch forms (libirr_b200.so only; the reference library takes the same data one particle per call)
    def set_jp_batch(self, addr, pos, vel, acc2, jrk6, mass, time):
        addr = np.ascontiguousarray(addr, dtype=np.int32); n = addr.shape[0]
        a = [np.ascontiguousarray(q, dtype=np.float64) for q in (pos, vel, acc2, jrk6, mass, time)]
        if self.is_b200:
            self.lib.irr_b200_set_jp_batch_(C.byref(C.c_int(n)), addr.ctypes.data_as(_ip), *[q.ctypes.data_as(_dp) for q in a])
        else:
            for k in range(n):
                self.set_jp(int(addr[k]), a[0][k], a[1][k], a[2][k], a[3][k], float(a[4][k]), float(a[5][k]))

    def set_list_batch(self, addr, lists):
        """lists[k] = [nnb, j1, ...] (1-based), rows of equal stride with >= 8 spare entries behind the longest list."""
        addr = np.ascontiguousarray(addr, dtype=np.int32); lists = np.ascontiguousarray(lists, dtype=np.int32)
        if self.is_b200:
            self.lib.irr_b200_set_list_batch_(C.byref(C.c_int(addr.shape[0])), addr.ctypes.data_as(_ip),
                                              C.byref(C.c_int(lists.shape[1])), lists.ctypes.data_as(_ip))
        else:
            for k in range(addr.shape[0]):
                self.lib.irr_simd_set_list_(C.byref(C.c_int(int(addr[k]))), lists[k].ctypes.data_as(_ip))

    def particle_records(self):
        """(device address of the record of address 1, doubles per record) of the particle table on the device."""
        stride = C.c_int(0)
        p = self.lib.irr_b200_particle_records_(C.byref(stride))
        return int(p), int(stride.value)

    def flush(self):
        """Pending set_jp / set_list reach the device; the library's stream is drained."""
        self.lib.irr_b200_flush_()

    def set_timing(self, on: bool):
        self.lib.irr_b200_set_timing(int(on))

    def counters(self):
        out = np.zeros(3)
        self.lib.irr_b200_counters(out.ctypes.data_as(_dp))
        return dict(kernel_ms=float(out[0]), calls=float(out[1]), interactions=float(out[2]))


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
