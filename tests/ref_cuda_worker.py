"""Worker for tests/test_reference_cuda_gpu.py (own process: the reference's CUDA library asserts / aborts on its own
errors).  Runs the reference's gpunb.velocity.cu + gpupot.gpu.cu, built UNMODIFIED for sm_100 into
oracle/_ref/libgpunb_ref_gpu.so, and this repo's library on identical snapshots through the identical C-ABI on the
same GPU: parity at N=16384 and the regular-force rate of both at N=1M ("the kernel to beat").  Prints one JSON line."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("OMP_NUM_THREADS", "8")
os.environ.setdefault("GPU_LIST", "0")
import numpy as np  # noqa: E402

import oracle_lib  # noqa: E402
from nbody6ppgpu_b200 import load, snapshots as S  # noqa: E402


def main():
    ref = oracle_lib.ref_gpu()
    b200 = load()
    ref.devinit(0); b200.devinit(0)
    out = {}
    # ---- parity, N = 16384, both neighbour criteria
    n = 16384
    m, x, v = S.plummer(n, 8, "kroupa")
    for m_flag in (0, 1):
        h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 100.0), 0.125, m_flag)
        res = {}
        for name, lib in (("b200", b200), ("ref", ref)):
            lib.open(n + 10, 0)
            lib.send(m, x, v)
            res[name] = [a.copy() for a in lib.regf(h2[:1024], dtr[:1024], x[:1024], v[:1024], 400, 350, m_flag)]
            lib.close()
        bad = oracle_lib.list_rows_equal(res["b200"][3], res["ref"][3])
        out[f"parity_mflag{m_flag}"] = {
            "rows_differing": len(bad), "mean_nnb": float(res["ref"][3][:, 0].mean()),
            "acc_relerr": oracle_lib.relerr(res["b200"][0], res["ref"][0]),
            "jrk_relerr": oracle_lib.relerr(res["b200"][1], res["ref"][1]),
            "pot_relerr": oracle_lib.relerr(res["b200"][2], res["ref"][2])}
    phi_b = b200.gpupot(1, n, m, x); phi_r = ref.gpupot(1, n, m, x)
    out["gpupot_relerr"] = float(np.max(np.abs(phi_b - phi_r) / np.abs(phi_r)))
    # ---- rate at N = 1M, 1024 i per call, host arrays (both through the same caller)
    n = int(os.environ.get("REFCUDA_N", "1000000"))
    calls = 24
    m, x, v = S.plummer(n, 1, "kroupa")
    h2, dtr = S.radii_nnb(x, m, 200.0)
    lists = {}
    for name, lib in (("ref", ref), ("b200", b200)):
        lists[name] = []
        lib.open(n + 10, 0)
        lib.send(m, x, v)
        call = lib.block_caller(h2, dtr, x, v, 1024, 600, 550, 0)
        for b in range(4):
            call(b * 1024, 1024)
        t0 = time.perf_counter()
        nnb = 0
        for b in range(calls):
            lst = call((4 + b) * 1024, 1024)[3]
            nnb += int(lst[:, 0].sum())
        t = time.perf_counter() - t0
        for b in range(calls):                     # untimed second pass: keep the lists for the comparison below
            lists[name].append(call((4 + b) * 1024, 1024)[3].copy())
        t0 = time.perf_counter()
        lib.send(m, x, v)
        ts = time.perf_counter() - t0
        # a steady-state regular block (intgrt.F:912-974): upload the predicted snapshot, then ONE call with few i
        blockcost = {}
        for ni in (64, 256):
            t0 = time.perf_counter()
            for b in range(5):
                lib.send(m, x, v)
                call((30 + b) * 1024, ni)
            blockcost[f"send+regf_ni{ni}_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        if lib.is_b200:                            # the same block with the device-resident predictor
            z3 = np.zeros((n, 3))
            lib.state_all(m, x, v, z3, z3, np.zeros(n))
            for ni in (64, 256):
                t0 = time.perf_counter()
                for b in range(5):
                    lib.predict_send(n, 0.0)
                    call((30 + b) * 1024, ni)
                blockcost[f"predict_send+regf_ni{ni}_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        out[f"regular_block_1M_{name}"] = blockcost
        lib.close()
        out[f"rate_{name}"] = {"gint_per_s": 1024.0 * calls * n / t * 1e-9, "us_per_call": t / calls * 1e6,
                               "send_ms": ts * 1e3, "mean_nnb": nnb / (1024.0 * calls), "n": n}
    # ---- small active blocks (config samples/N10k_B1k.input: N = 10k, KS-heavy, <Ni> small): microseconds per
    # gpunb_regf_ call against a fixed j-set, same caller for both libraries
    n = 10000
    m2, x2, v2 = S.plummer(n, 4, "kroupa")
    h22, dtr2 = S.radii_nnb(x2, m2, 100.0)
    small = {}
    for name, lib in (("ref", ref), ("b200", b200)):
        lib.open(n + 10, 0)
        lib.send(m2, x2, v2)
        call = lib.block_caller(h22, dtr2, x2, v2, 1024, 400, 350, 0)
        small[name] = {}
        for ni in (1, 8, 32, 64, 256, 1024):
            for b in range(10):
                call((37 * b) % (n - ni), ni)
            reps = 200
            t0 = time.perf_counter()
            for b in range(reps):
                call((97 * b) % (n - ni), ni)
            small[name][str(ni)] = (time.perf_counter() - t0) / reps * 1e6
        t0 = time.perf_counter()
        for b in range(20):
            lib.send(m2, x2, v2)
        small[name]["send"] = (time.perf_counter() - t0) / 20 * 1e6
        lib.close()
    out["small_blocks_N10k_us_per_call"] = small
    # BASELINE config 2 (samples/N16k.input: N = 16000, m_flag = 1): device-resident sweep of this library as a fraction of
    # the FP32 roofline (the block of a resident sweep is free: 4736 = 32 x 148 i-tiles, four launch groups)
    n = 16000
    m3, x3, v3 = S.plummer(n, 6, "kroupa")
    h23, dtr3 = S.radii_nnb(x3, m3, 100.0, 0.125, 1)
    b200.open(n + 10, 0)
    b200.send(m3, x3, v3); b200.set_radii(h23, dtr3)
    for _ in range(3):
        b200.sweep_resident(0, n, 4736, 400, 350, 1)
    ms = min(b200.sweep_resident(0, n, 4736, 400, 350, 1) for _ in range(5))
    b200.close()
    out["sweep_N16k_mflag1"] = {"ms": ms, "gint_per_s": float(n) * n / ms * 1e-6, "frac_of_fp32_roofline": float(n) * n / ms * 1e-6 / 1240.8}
    # list parity at N = 1M over the timed calls; every differing pair is reported with its distance from the boundary
    diffs = []
    for b in range(calls):
        for r in oracle_lib.list_rows_equal(lists["b200"][b], lists["ref"][b]):
            i = (4 + b) * 1024 + r
            la, lb = lists["b200"][b][r], lists["ref"][b][r]
            sa, sb = set(la[1:1 + la[0]].tolist()), set(lb[1:1 + lb[0]].tolist())
            for j in sorted(sa ^ sb):
                xi32, xj32 = x[i].astype(np.float32).astype(np.float64), x[j].astype(np.float32).astype(np.float64)
                dv = v[j].astype(np.float32).astype(np.float64) - v[i].astype(np.float32).astype(np.float64)
                d = xj32 - xi32
                dp = d + float(np.float32(dtr[i])) * dv
                lim = float(np.float32(h2[i]))
                diffs.append({"i": int(i), "j": int(j), "in": "b200" if j in sa else "ref",
                              "min_r2_over_h2_minus_1": float(min((d * d).sum(), (dp * dp).sum()) / lim - 1.0)})
    out["lists_1M"] = {"rows_compared": 1024 * calls, "pairs_differing": len(diffs), "detail": diffs[:8]}
    print("REFCUDA " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
