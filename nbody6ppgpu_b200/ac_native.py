"""ctypes front of the native Ahmad-Cohen driver (csrc/ac_driver.cpp -> libac_driver.so): the C++ twin of hermite_ac.py,
used where the wall-clock of an integration has to measure the libraries and not a numpy harness (bench.py --time-unit).
The libraries behind it are named by PATH and dlopen'ed by the driver itself."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent


class ACParams(C.Structure):
    _fields_ = [("nnbopt", C.c_int), ("lmax", C.c_int), ("m_flag", C.c_int), ("use_predictor", C.c_int), ("use_regcor", C.c_int),
                ("eta_i", C.c_double), ("eta_r", C.c_double), ("dtmax", C.c_double), ("dtmin", C.c_double), ("t_end", C.c_double),
                ("rs0", C.c_double)]


class ACStats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("t", "wall_total", "wall_send", "wall_regf", "wall_irr", "wall_regcor", "wall_energy",
                                          "wall_init", "e0", "e1")] + \
               [(k, C.c_longlong) for k in ("irr_steps", "reg_steps", "block_steps", "reg_blocks", "regf_calls", "overflow_retries")] + \
               [("mean_nnb", C.c_double)]


def lib_path() -> Path:
    return _HERE / "libac_driver.so"


def run(gpunb_so, irr_so, m, x, v, t_end, *, nnbopt=40, lmax=128, eta_i=0.02, eta_r=0.02, dtmax=0.125, dtmin=2.0 ** -22, m_flag=0,
        rs0=0.0, use_predictor=False, use_regcor=False):
    """Integrate t_end N-body time units; returns (stats dict, x, v).  irr_so None: irregular sums on the host in fp64.
    use_predictor: 0 host predictor + gpunb_send_, 1 device-resident predictor with its own state (gpunb_b200_state_*),
    2 device-resident predictor reading the irregular-force library's particle table (one copy of the state on the device).
    use_regcor: 0 host list bookkeeping, 1 gpunb_b200_regcor_ with the old lists passed by the driver, 2 with the old lists in
    the library's device-resident list store."""
    so = lib_path()
    if not so.exists():
        raise RuntimeError(f"{so} not found -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(str(so))
    m = np.ascontiguousarray(m, dtype=np.float64); n = m.shape[0]
    x = np.ascontiguousarray(x, dtype=np.float64); v = np.ascontiguousarray(v, dtype=np.float64)
    xo = np.zeros((n, 3)); vo = np.zeros((n, 3))
    p = ACParams(nnbopt, lmax, m_flag, int(use_predictor), int(use_regcor), eta_i, eta_r, dtmax, dtmin, t_end, rs0)
    st = ACStats()
    dp = C.POINTER(C.c_double)
    lib.ac_driver_run.restype = C.c_int
    rc = lib.ac_driver_run(str(gpunb_so).encode(), str(irr_so).encode() if irr_so else None, C.c_int(n), m.ctypes.data_as(dp),
                           x.ctypes.data_as(dp), v.ctypes.data_as(dp), C.byref(p), C.byref(st), xo.ctypes.data_as(dp), vo.ctypes.data_as(dp))
    if rc != 0:
        raise RuntimeError(f"ac_driver_run failed ({rc})")
    out = {k: getattr(st, k) for k, _ in ACStats._fields_}
    out["dE_over_E"] = (st.e1 - st.e0) / abs(st.e0)
    return out, xo, vo
