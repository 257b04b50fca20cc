"""ctypes mirror of the regular-force C-ABI.

``ForceLib(path)`` binds ANY shared library that exports the reference's seven symbols
(gpunb_devinit_ is optional, reg.avx.cpp does not have it), so the same Python calls drive

* ``nbody6ppgpu_b200/libgpunb_b200.so``  -- this repo's CUDA library (the product),
* ``oracle/_ref/libgpunb_ref_avx.so``    -- the reference's CPU library (tests / CPU baseline only),
* ``oracle/_ref/libgpunb_ref_gpu.so``    -- the reference's CUDA library (tests / comparison only).

Argument meaning follows the reference (src/Main/gpunb.velocity.cu:904-939, gpupot.gpu.cu:116-127):
all scalars by reference, REAL*8 arrays, ``X(3,N)`` == C ``x[N][3]``, neighbour rows ``list[i*lmax]``
= count (negative on overflow) followed by 0-based ascending j.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_c_int_p = C.POINTER(C.c_int)
_c_dbl_p = C.POINTER(C.c_double)

N_COUNTERS = 28
CTR = dict(grav_ms=0, grav_launches=1, launches=2, h2d_bytes=3, d2h_bytes=4, interactions=5,
           merge_ms=6, pot_ms=7, near_tiles=8, all_tiles=9,
           tl_blocks=10, tl_isort_ms=11, tl_regf_ms=12, tl_merge_ms=13, tl_exch_ms=14,
           host_pack_ms=15, host_enqueue_ms=16, host_wait_ms=17, host_scatter_ms=18,
           sends=19, send_ms=20, send_stage_ms=21, send_tiles_ms=22, transposed_tiles=23, sends_order_kept=24, host_rendezvous_ms=25,
           regcor_ms=26, regcor_rows=27)


class LibraryMissing(RuntimeError):
    pass


def lib_path() -> Path:
    return _HERE / "libgpunb_b200.so"


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_c_dbl_p)


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class ForceLib:
    """One loaded regular-force library (file-static state, like the reference: one per process)."""

    def __init__(self, path: os.PathLike | str, mode: int = C.RTLD_LOCAL):
        path = Path(path)
        if not path.exists():
            raise LibraryMissing(f"{path} not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = C.CDLL(str(path), mode=mode)
        L = self.lib
        self.has_devinit = hasattr(L, "gpunb_devinit_")
        self.is_b200 = hasattr(L, "gpunb_b200_version")
        L.gpunb_open_.argtypes = [_c_int_p, _c_int_p]
        L.gpunb_close_.argtypes = []
        L.gpunb_send_.argtypes = [_c_int_p, _c_dbl_p, _c_dbl_p, _c_dbl_p]
        L.gpunb_regf_.argtypes = [_c_int_p] + [_c_dbl_p] * 7 + [_c_int_p, _c_int_p, _c_int_p, _c_int_p]
        L.gpunb_profile_.argtypes = [_c_int_p]
        L.gpupot_.argtypes = [_c_int_p] * 4 + [_c_dbl_p] * 3
        for f in (L.gpunb_open_, L.gpunb_close_, L.gpunb_send_, L.gpunb_regf_, L.gpunb_profile_, L.gpupot_):
            f.restype = None
        if self.has_devinit:
            L.gpunb_devinit_.argtypes = [_c_int_p]
            L.gpunb_devinit_.restype = None
        if self.is_b200:
            L.gpunb_b200_version.restype = C.c_int
            L.gpunb_b200_build_info.restype = C.c_char_p
            L.gpunb_b200_num_devices.restype = C.c_int
            L.gpunb_b200_resident_warps.restype = C.c_int
            L.gpunb_b200_get_counters.argtypes = [_c_dbl_p]
            L.gpunb_b200_set_radii.argtypes = [_c_int_p, _c_dbl_p, _c_dbl_p]
            L.gpunb_b200_sweep_resident.argtypes = [_c_int_p] * 6
            L.gpunb_b200_sweep_resident.restype = C.c_float
            L.gpunb_b200_fetch_last.argtypes = [_c_int_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_int_p, _c_int_p]
            L.gpunb_b200_fp32_microbench.argtypes = [C.c_int, C.c_int]
            L.gpunb_b200_fp32_microbench.restype = C.c_double
            L.gpunb_b200_nccl_unique_id.argtypes = [C.c_char_p]
            L.gpunb_b200_nccl_unique_id.restype = C.c_int
            L.gpunb_b200_nccl_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
            L.gpunb_b200_nccl_init.restype = C.c_int
            L.gpunb_b200_nccl_finalize.argtypes = []
            L.gpunb_b200_nccl_finalize.restype = None
            L.gpunb_b200_has_near_scalar_ab.restype = C.c_int
            L.gpunb_b200_set_near_exact.argtypes = [C.c_int]
            L.gpunb_b200_set_near_exact.restype = None
            L.gpunb_b200_pin_host_.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
            L.gpunb_b200_pin_host_.restype = C.c_int
            L.gpunb_b200_unpin_host_.argtypes = [C.c_void_p]
            L.gpunb_b200_unpin_host_.restype = None
            L.gpunb_b200_set_resort_every.argtypes = [C.c_int]
            L.gpunb_b200_set_resort_every.restype = None
            L.gpunb_b200_set_send_scatter.argtypes = [C.c_int]
            L.gpunb_b200_set_send_scatter.restype = None
            L.gpunb_b200_set_regf_oversub.argtypes = [C.c_int]
            L.gpunb_b200_set_regf_oversub.restype = None
            L.gpunb_b200_set_taper.argtypes = [C.c_int]
            L.gpunb_b200_set_taper.restype = None
            L.gpunb_b200_set_islice.argtypes = [C.c_int]
            L.gpunb_b200_set_islice.restype = None
            L.gpunb_b200_set_sub_pairs.argtypes = [C.c_double]
            L.gpunb_b200_set_sub_pairs.restype = None
            L.gpunb_b200_set_isort_pairs.argtypes = [C.c_double]
            L.gpunb_b200_set_isort_pairs.restype = None
            L.gpunb_b200_set_tuning.argtypes = [C.c_int, C.c_int]
            L.gpunb_b200_set_tuning.restype = None
            L.gpunb_b200_state_all_.argtypes = [_c_int_p] + [_c_dbl_p] * 6
            L.gpunb_b200_state_update_.argtypes = [_c_int_p, _c_int_p] + [_c_dbl_p] * 6
            L.gpunb_b200_predict_send_.argtypes = [_c_int_p, _c_dbl_p]
            L.gpunb_b200_get_predicted_.argtypes = [_c_int_p, _c_int_p, _c_dbl_p, _c_dbl_p]
            L.gpunb_b200_predict_send_records_.argtypes = [_c_int_p, _c_dbl_p, C.c_void_p, _c_int_p]
            L.gpunb_b200_predict_send_records_.restype = None
            for f in (L.gpunb_b200_state_all_, L.gpunb_b200_state_update_, L.gpunb_b200_predict_send_, L.gpunb_b200_get_predicted_):
                f.restype = None
            L.gpunb_b200_regcor_.argtypes = ([_c_int_p] * 8 + [_c_dbl_p, _c_dbl_p, _c_dbl_p, _c_int_p] + [_c_dbl_p] * 4
                                             + [_c_int_p] * 4)
            L.gpunb_b200_regcor_last_.argtypes = L.gpunb_b200_regcor_.argtypes
            L.gpunb_b200_regcor_last_.restype = None
            L.gpunb_b200_lists_put_.argtypes = [_c_int_p] * 4
            L.gpunb_b200_lists_get_.argtypes = [_c_int_p] * 4
            L.gpunb_b200_steps_all_.argtypes = [_c_int_p, _c_dbl_p]
            L.gpunb_b200_steps_update_.argtypes = [_c_int_p, _c_int_p, _c_dbl_p]
            for f in (L.gpunb_b200_regcor_, L.gpunb_b200_lists_put_, L.gpunb_b200_lists_get_, L.gpunb_b200_steps_all_,
                      L.gpunb_b200_steps_update_):
                f.restype = None
        self.nj = 0

    # ---- the reference interface -------------------------------------------------------------
    def devinit(self, irank: int = 0):
        if self.has_devinit:
            self.lib.gpunb_devinit_(C.byref(C.c_int(irank)))

    def open(self, nbmax: int, irank: int = 0):
        self.lib.gpunb_open_(C.byref(C.c_int(nbmax)), C.byref(C.c_int(irank)))

    def close(self):
        self.lib.gpunb_close_()

    def send(self, m, x, v):
        m = _f64(m)
        nj = m.shape[0]
        x = _f64(x, (nj, 3))
        v = _f64(v, (nj, 3))
        self.nj = nj
        self.lib.gpunb_send_(C.byref(C.c_int(nj)), _dp(m), _dp(x), _dp(v))

    def regf(self, h2, dtr, xi, vi, lmax: int, nnbmax: int, m_flag: int = 0, pad: int = 8):
        """Returns (acc[ni,3], jrk[ni,3], pot[ni], list[ni,lmax]).

        ``pad`` extra rows are allocated behind every array because the reference's AVX library
        reads/writes up to 3 elements past ni (reg.avx.cpp:204-314); rows >= ni are never returned.
        """
        h2 = _f64(h2)
        ni = h2.shape[0]
        n = ni + pad

        def padded(a, shape):
            out = np.zeros((n,) + shape, dtype=np.float64)
            out[:ni] = _f64(a, (ni,) + shape)
            return out

        h2p, dtrp = padded(h2, ()), padded(dtr, ())
        xip, vip = padded(xi, (3,)), padded(vi, (3,))
        acc = np.zeros((n, 3)); jrk = np.zeros((n, 3)); pot = np.zeros(n)
        lst = np.zeros((n, lmax), dtype=np.int32)
        self.lib.gpunb_regf_(C.byref(C.c_int(ni)), _dp(h2p), _dp(dtrp), _dp(xip), _dp(vip),
                             _dp(acc), _dp(jrk), _dp(pot),
                             C.byref(C.c_int(lmax)), C.byref(C.c_int(nnbmax)),
                             lst.ctypes.data_as(_c_int_p), C.byref(C.c_int(m_flag)))
        return acc[:ni], jrk[:ni], pot[:ni], lst[:ni]

    def caller_arrays(self, nimax: int, lmax: int, pad: int = 8):
        """Caller-owned output arrays, allocated ONCE like the Fortran caller's static GPUACC/GPUJRK/GPUPHI/LISTGP
        (util_gpu.F:14, fpoly0.F:11): (acc[n,3], jrk[n,3], pot[n], list[n,lmax]) with n = nimax + pad."""
        n = nimax + pad
        return (np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), np.zeros((n, lmax), dtype=np.int32))

    def regf_into(self, out, h2, dtr, xi, vi, lmax: int, nnbmax: int, m_flag: int = 0):
        """gpunb_regf_ on caller-owned arrays (see caller_arrays); inputs must be C-contiguous float64 and, for the
        reference AVX library, followed by >= 3 readable rows (slices of larger arrays are).  Returns views [:ni]."""
        acc, jrk, pot, lst = out
        ni = h2.shape[0]
        if lst.shape[1] != lmax or acc.shape[0] < ni:
            raise ValueError("caller arrays do not match ni/lmax")
        for a in (h2, dtr, xi, vi):
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("regf_into needs C-contiguous float64 inputs")
        self.lib.gpunb_regf_(C.byref(C.c_int(ni)), _dp(h2), _dp(dtr), _dp(xi), _dp(vi), _dp(acc), _dp(jrk), _dp(pot),
                             C.byref(C.c_int(lmax)), C.byref(C.c_int(nnbmax)),
                             lst.ctypes.data_as(_c_int_p), C.byref(C.c_int(m_flag)))
        return acc[:ni], jrk[:ni], pot[:ni], lst[:ni]

    def block_caller(self, h2, dtr, x, v, nimax: int, lmax: int, nnbmax: int, m_flag: int = 0, pad: int = 8):
        """What the Fortran caller is (fpoly0.F:72-125, util_gpu.F:33-60): static input/output arrays and by-reference
        scalars set up ONCE, then one bare ``gpunb_regf_`` call per i-block [i0, i0+ni) of the arrays -- no per-call
        Python marshalling beyond four address additions.  Returns ``call(i0, ni) -> (acc, jrk, pot, list)`` (views
        of the caller-owned output arrays, rows [:ni])."""
        for a in (h2, dtr, x, v):
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("block_caller needs C-contiguous float64 arrays")
        n = h2.shape[0]
        acc, jrk, pot, lst = self.caller_arrays(nimax, lmax, pad)
        c_ni, c_lmax, c_nnbmax, c_mflag = C.c_int(0), C.c_int(lmax), C.c_int(nnbmax), C.c_int(m_flag)
        fn = C.CFUNCTYPE(None, *([C.c_void_p] * 12))(("gpunb_regf_", self.lib))
        a_h2, a_dtr, a_x, a_v = h2.ctypes.data, dtr.ctypes.data, x.ctypes.data, v.ctypes.data
        fixed = (acc.ctypes.data, jrk.ctypes.data, pot.ctypes.data, C.addressof(c_lmax), C.addressof(c_nnbmax),
                 lst.ctypes.data, C.addressof(c_mflag))
        p_ni = C.addressof(c_ni)
        keep = (h2, dtr, x, v, acc, jrk, pot, lst, c_ni, c_lmax, c_nnbmax, c_mflag)

        def call(i0: int, ni: int, _keep=keep):
            if ni > nimax or i0 < 0 or i0 + ni > n:
                raise ValueError("i-block outside the arrays")
            c_ni.value = ni
            fn(p_ni, a_h2 + 8 * i0, a_dtr + 8 * i0, a_x + 24 * i0, a_v + 24 * i0, *fixed)
            return acc[:ni], jrk[:ni], pot[:ni], lst[:ni]
        call.outputs = (acc, jrk, pot, lst)            # the caller-owned result arrays (e.g. for pin_host)
        return call

    def profile(self, irank: int = 0):
        self.lib.gpunb_profile_(C.byref(C.c_int(irank)))

    def gpupot(self, istart: int, ni: int, m, x, irank: int = 0, pad: int = 8):
        """pot[ii] = sum_{j, r>0} m_j / r_ij for i = istart-1+ii (istart is 1-based)."""
        m = _f64(m)
        n = m.shape[0]
        # pot.avx.cpp:90-137 reads ptcl[i..i+7] and writes pot[i..i+7] past ni: pad both.
        mp = np.zeros(n + pad); mp[:n] = m
        xp = np.zeros((n + pad, 3)); xp[:n] = _f64(x, (n, 3))
        # Quirk of the reference: the CUDA library (and the MPI caller, energy_mpi.F:100-101, which passes
        # phidbl(istart+ifirst-1)) index pot[] RELATIVE to istart (gpupot.gpu.cu:102-105); the AVX twin
        # writes pot[istart-1+ii] (pot.avx.cpp:88-137: pot_reduce(..., pot+i) with absolute i).  This
        # library follows the CUDA/caller convention; the wrapper hides the AVX twin's offset.
        avx_abs = (not self.is_b200) and (not self.has_devinit)
        pot = np.zeros((n if avx_abs else ni) + pad)
        self.lib.gpupot_(C.byref(C.c_int(irank)), C.byref(C.c_int(istart)), C.byref(C.c_int(ni)),
                         C.byref(C.c_int(n)), _dp(mp), _dp(xp), _dp(pot))
        return pot[istart - 1:istart - 1 + ni].copy() if avx_abs else pot[:ni]

    # ---- gpunb_b200 extensions ---------------------------------------------------------------
    def _need_b200(self):
        if not self.is_b200:
            raise RuntimeError(f"{self.path.name} is not libgpunb_b200.so")

    def version(self) -> int:
        self._need_b200()
        return int(self.lib.gpunb_b200_version())

    def build_info(self) -> str:
        self._need_b200()
        return self.lib.gpunb_b200_build_info().decode()

    def num_devices(self) -> int:
        self._need_b200()
        return int(self.lib.gpunb_b200_num_devices())

    def resident_warps(self) -> int:
        self._need_b200()
        return int(self.lib.gpunb_b200_resident_warps())

    def counters(self) -> dict:
        self._need_b200()
        buf = np.zeros(N_COUNTERS)
        self.lib.gpunb_b200_get_counters(_dp(buf))
        return {k: float(buf[i]) for k, i in CTR.items()}

    def reset_counters(self):
        self._need_b200()
        self.lib.gpunb_b200_reset_counters()

    def set_radii(self, h2, dtr):
        self._need_b200()
        h2 = _f64(h2); dtr = _f64(dtr)
        self.lib.gpunb_b200_set_radii(C.byref(C.c_int(h2.shape[0])), _dp(h2), _dp(dtr))

    def sweep_resident(self, i0: int, ni: int, block: int, lmax: int, nnbmax: int, m_flag: int = 0) -> float:
        """Device-resident regular-force sweep over i = j[i0:i0+ni]; returns device milliseconds."""
        self._need_b200()
        a = [C.c_int(v) for v in (i0, ni, block, lmax, nnbmax, m_flag)]
        return float(self.lib.gpunb_b200_sweep_resident(*[C.byref(v) for v in a]))

    def fetch_last(self, lmax: int, nimax: int = 2048):
        self._need_b200()
        acc = np.zeros((nimax, 3)); jrk = np.zeros((nimax, 3)); pot = np.zeros(nimax)
        lst = np.zeros((nimax, lmax), dtype=np.int32)
        n = C.c_int(0); lm = C.c_int(lmax)
        self.lib.gpunb_b200_fetch_last(C.byref(n), _dp(acc), _dp(jrk), _dp(pot), C.byref(lm),
                                       lst.ctypes.data_as(_c_int_p))
        if lm.value != lmax:
            raise ValueError("lmax differs from the one used by the sweep")
        k = n.value
        return acc[:k], jrk[:k], pot[:k], lst[:k]

    def nccl_unique_id(self) -> bytes:
        """128-byte ncclUniqueId (call on rank 0, broadcast to the other ranks)."""
        self._need_b200()
        buf = C.create_string_buffer(128)
        rc = self.lib.gpunb_b200_nccl_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
        return buf.raw

    def nccl_init(self, rank: int, nranks: int, uid: bytes):
        """Join one-GPU-per-process ranks into one j-sharded library (after devinit, before open)."""
        self._need_b200()
        assert len(uid) == 128
        rc = self.lib.gpunb_b200_nccl_init(rank, nranks, uid)
        if rc != 0:
            raise RuntimeError(f"gpunb_b200_nccl_init failed ({rc})")

    def nccl_finalize(self):
        self._need_b200()
        self.lib.gpunb_b200_nccl_finalize()

    # device-resident predictor (xbpredall.f + gpunb_send_ without the upload); F = force/2, FDOT = derivative/6
    def state_all(self, body, x0, x0dot, f, fdot, t0):
        self._need_b200()
        body = _f64(body); nj = body.shape[0]
        a = [_f64(x0, (nj, 3)), _f64(x0dot, (nj, 3)), _f64(f, (nj, 3)), _f64(fdot, (nj, 3)), _f64(t0, (nj,))]
        self.nj = nj
        self.lib.gpunb_b200_state_all_(C.byref(C.c_int(nj)), _dp(body), *[_dp(q) for q in a])

    def state_update(self, idx, body, x0, x0dot, f, fdot, t0):
        self._need_b200()
        idx = np.ascontiguousarray(idx, dtype=np.int32); n = idx.shape[0]
        a = [_f64(body, (n,)), _f64(x0, (n, 3)), _f64(x0dot, (n, 3)), _f64(f, (n, 3)), _f64(fdot, (n, 3)), _f64(t0, (n,))]
        self.lib.gpunb_b200_state_update_(C.byref(C.c_int(n)), idx.ctypes.data_as(_c_int_p), *[_dp(q) for q in a])

    def predict_send(self, nj: int, time: float):
        self._need_b200()
        self.nj = nj
        self.lib.gpunb_b200_predict_send_(C.byref(C.c_int(nj)), C.byref(C.c_double(time)))

    def predict_send_records(self, nj: int, time: float, records_dev: int, stride: int):
        """predict_send from particle records another library keeps on the same device (device address of the record of the
        first j-particle, doubles per record) -- libirr_b200.so's table: IrrLib.particle_records()."""
        self._need_b200()
        self.nj = nj
        self.lib.gpunb_b200_predict_send_records_(C.byref(C.c_int(nj)), C.byref(C.c_double(time)), C.c_void_p(records_dev),
                                                  C.byref(C.c_int(stride)))

    def get_predicted(self, idx):
        self._need_b200()
        idx = np.ascontiguousarray(idx, dtype=np.int32); n = idx.shape[0]
        x = np.zeros((n, 3)); v = np.zeros((n, 3))
        self.lib.gpunb_b200_get_predicted_(C.byref(C.c_int(n)), idx.ctypes.data_as(_c_int_p), _dp(x), _dp(v))
        return x, v

    # neighbour-list bookkeeping after gpunb_regf_ (util_gpu.F:102-111 + regcor_gpu.F:267-470), batched on the device
    def regcor(self, index_i, ifirst: int, n: int, ntot: int, new_list, old_list, rs2, step, smin: float, nnbmax: int,
               freg, fdr, dfirr=None, dfd=None, last_lmax: int = 0):
        """Returns dict(nlist, nbloss, nbgain, jjlist, freg, fdr, dfirr, dfd, nbsmin); argument meaning as
        include/gpunb_b200.h part 3 (old_list None: resident list store; step None: resident steps or no retention).
        new_list None + last_lmax: gpunb_b200_regcor_last_ (the rows of the last gpunb_regf_ call, still on the device)."""
        self._need_b200()
        index_i = np.ascontiguousarray(index_i, dtype=np.int32); ni = index_i.shape[0]
        last = new_list is None
        nl = np.zeros((ni, last_lmax), dtype=np.int32) if last else np.array(new_list, dtype=np.int32, order="C")
        lmax = nl.shape[1]
        ol = None if old_list is None else np.ascontiguousarray(old_list, dtype=np.int32)
        if ol is not None and ol.shape != nl.shape:
            raise ValueError("old_list and new_list differ in shape")
        st = None if step is None else _f64(step)
        rs2 = _f64(rs2, (ni,))
        fr = np.array(freg, dtype=np.float64, order="C"); fd = np.array(fdr, dtype=np.float64, order="C")
        di = np.zeros((ni, 3)) if dfirr is None else np.array(dfirr, dtype=np.float64, order="C")
        dd = np.zeros((ni, 3)) if dfd is None else np.array(dfd, dtype=np.float64, order="C")
        nbloss = np.zeros(ni, dtype=np.int32); nbgain = np.zeros(ni, dtype=np.int32)
        jj = np.zeros((ni, 2 * lmax), dtype=np.int32)
        nbsmin = C.c_int(0)
        ip = lambda a: a.ctypes.data_as(_c_int_p)
        fn = self.lib.gpunb_b200_regcor_last_ if last else self.lib.gpunb_b200_regcor_
        fn(C.byref(C.c_int(ni)), ip(index_i), C.byref(C.c_int(ifirst)), C.byref(C.c_int(n)),
                                    C.byref(C.c_int(ntot)), C.byref(C.c_int(lmax)), ip(nl), None if ol is None else ip(ol),
                                    _dp(rs2), None if st is None else _dp(st), C.byref(C.c_double(smin)),
                                    C.byref(C.c_int(nnbmax)), _dp(fr), _dp(fd), _dp(di), _dp(dd), ip(nbloss), ip(nbgain),
                                    ip(jj), C.byref(nbsmin))
        return dict(nlist=nl, nbloss=nbloss, nbgain=nbgain, jjlist=jj, freg=fr, fdr=fd, dfirr=di, dfd=dd,
                    nbsmin=int(nbsmin.value))

    def lists_put(self, index_i, lists):
        self._need_b200()
        index_i = np.ascontiguousarray(index_i, dtype=np.int32); lists = np.ascontiguousarray(lists, dtype=np.int32)
        self.lib.gpunb_b200_lists_put_(C.byref(C.c_int(index_i.shape[0])), index_i.ctypes.data_as(_c_int_p),
                                       C.byref(C.c_int(lists.shape[1])), lists.ctypes.data_as(_c_int_p))

    def lists_get(self, index_i, lmax: int):
        self._need_b200()
        index_i = np.ascontiguousarray(index_i, dtype=np.int32)
        out = np.zeros((index_i.shape[0], lmax), dtype=np.int32)
        self.lib.gpunb_b200_lists_get_(C.byref(C.c_int(index_i.shape[0])), index_i.ctypes.data_as(_c_int_p),
                                       C.byref(C.c_int(lmax)), out.ctypes.data_as(_c_int_p))
        return out

    def steps_all(self, step):
        self._need_b200()
        step = _f64(step)
        self.lib.gpunb_b200_steps_all_(C.byref(C.c_int(step.shape[0])), _dp(step))

    def steps_update(self, idx, step):
        self._need_b200()
        idx = np.ascontiguousarray(idx, dtype=np.int32); step = _f64(step, (idx.shape[0],))
        self.lib.gpunb_b200_steps_update_(C.byref(C.c_int(idx.shape[0])), idx.ctypes.data_as(_c_int_p), _dp(step))

    def set_tuning(self, nslot: int = 0, nsub: int = 0):
        """Pipeline depth: slots of a resident sweep / sub-blocks of one gpunb_regf_ call (0 = leave unchanged)."""
        self._need_b200()
        self.lib.gpunb_b200_set_tuning(nslot, nsub)

    def pin_host(self, *arrays) -> bool:
        """Pin caller-owned numpy arrays (what a Fortran caller does once for its static arrays): gpunb_send_ then
        uploads without the staging copy and gpunb_regf_ results are written straight into them.  Returns True when
        every array was pinned.  Call unpin_host before the arrays are released."""
        self._need_b200()
        ok = True
        for a in arrays:
            ok &= self.lib.gpunb_b200_pin_host_(C.c_void_p(a.ctypes.data), C.byref(C.c_longlong(a.nbytes))) == 0
        return ok

    def unpin_host(self, *arrays):
        self._need_b200()
        for a in arrays:
            self.lib.gpunb_b200_unpin_host_(C.c_void_p(a.ctypes.data))

    def set_resort_every(self, k: int):
        """Hilbert order across snapshots: 0 = adaptive (default), 1 = always sort, k > 1 = sort every k-th snapshot."""
        self._need_b200()
        self.lib.gpunb_b200_set_resort_every(k)

    def set_islice(self, on: int):
        """i-slice mode of the multi-process library: gpunb_regf_ becomes a collective call, every rank passes its own
        i-slice (possibly empty) and receives the results of that slice only."""
        self._need_b200()
        self.lib.gpunb_b200_set_islice(int(on))

    def set_sub_pairs(self, pairs: float):
        """Pairs a sub-block of gpunb_regf_ must keep for the call to be split (default 1.5e8)."""
        self._need_b200()
        self.lib.gpunb_b200_set_sub_pairs(float(pairs))

    def set_isort_pairs(self, pairs: float):
        """gpunb_regf_ calls below this many pairs skip the Morton sort of the i-block (default 2.5e7; 0 = always sort)."""
        self._need_b200()
        self.lib.gpunb_b200_set_isort_pairs(float(pairs))

    def set_send_scatter(self, min_nj: int):
        """One process per GPU: snapshots of at least min_nj particles are uploaded in R slices and all-gathered over NVLink
        (negative: never).  The same value on every rank."""
        self._need_b200()
        self.lib.gpunb_b200_set_send_scatter(int(min_nj))

    def set_regf_oversub(self, k: int):
        """Work items per resident warp slot of an unsplit gpunb_regf_ call (1 ... 4, default 4)."""
        self._need_b200()
        self.lib.gpunb_b200_set_regf_oversub(int(k))

    def set_taper(self, on: int):
        """Sub-block sizes of one gpunb_regf_ call: equal (0, default) or tapering (1)."""
        self._need_b200()
        self.lib.gpunb_b200_set_taper(on)

    def set_near_exact(self, on: int):
        """A/B: 1 = NEAR tiles through the scalar pair body, 0 = the packed body (same bits), -1 = environment."""
        self._need_b200()
        self.lib.gpunb_b200_set_near_exact(on)

    def fp32_microbench(self, mode: int, iters: int = 4096) -> float:
        self._need_b200()
        return float(self.lib.gpunb_b200_fp32_microbench(mode, iters))


def load() -> ForceLib:
    """Load this repo's CUDA library; raises LibraryMissing (no fallback) when it is not built."""
    return ForceLib(lib_path())
