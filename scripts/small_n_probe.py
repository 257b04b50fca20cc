"""Latency of the synchronous ABI at small N (BASELINE config N10k_B1k: KS-heavy, small active blocks; N16k): microseconds per
gpunb_send_ and per gpunb_regf_ call with the library's own host / device buckets, pageable and pinned caller arrays.
Usage: python scripts/small_n_probe.py [N=10000] [out.json]"""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from nbody6ppgpu_b200 import load, snapshots as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
lib = load(); lib.devinit(0)
m, x, v = S.plummer(n, 4, "kroupa")
h2, dtr = S.radii_nnb(x, m, 100.0)
lib.open(n + 10, 0)
out = {"n": n}
for pin in (0, 1):
    call = lib.block_caller(h2, dtr, x, v, 1024, 400, 350, 0)
    arrs = [m, x, v, *call.outputs]
    if pin:
        assert lib.pin_host(*arrs)
    tag = "pinned" if pin else "pageable"
    for _ in range(5):
        lib.send(m, x, v)
    lib.reset_counters()
    t0 = time.perf_counter()
    for _ in range(50):
        lib.send(m, x, v)
    t = (time.perf_counter() - t0) / 50
    c = lib.counters()
    out[f"send_{tag}"] = {"us": t * 1e6, "stage_us": c["send_stage_ms"] / 50 * 1e3, "tiles_device_us": c["send_tiles_ms"] / 50 * 1e3}
    print(f"{tag:8s} gpunb_send_ N={n}: {t * 1e6:7.1f} us (staging + upload {c['send_stage_ms'] / 50 * 1e3:6.1f}, device tile construction {c['send_tiles_ms'] / 50 * 1e3:6.1f})", flush=True)
    for ni in (1, 8, 32, 64, 256, 1024):
        for b in range(10):
            call((37 * b) % (n - ni), ni)
        reps = 200
        lib.reset_counters()
        t0 = time.perf_counter()
        for b in range(reps):
            call((97 * b) % (n - ni), ni)
        t = (time.perf_counter() - t0) / reps
        c = lib.counters()
        k = 1e3 / reps
        out[f"regf_{tag}_ni{ni}"] = {"us": t * 1e6, "pack": c["host_pack_ms"] * k, "enqueue": c["host_enqueue_ms"] * k, "wait": c["host_wait_ms"] * k,
                                     "scatter": c["host_scatter_ms"] * k, "pair_kernel": c["grav_ms"] * k, "merge": c["merge_ms"] * k}
        print(f"{tag:8s} gpunb_regf_ ni={ni:5d}: {t * 1e6:7.1f} us | host: pack {c['host_pack_ms'] * k:5.1f} enqueue {c['host_enqueue_ms'] * k:5.1f} "
              f"wait {c['host_wait_ms'] * k:6.1f} scatter {c['host_scatter_ms'] * k:5.1f} | device: pair kernel {c['grav_ms'] * k:6.1f} merge+rows {c['merge_ms'] * k:5.1f}", flush=True)
    if pin:
        lib.unpin_host(*arrs)
# full sweep rate (FPOLY0 pattern) device-resident and through the ABI
lib.send(m, x, v); lib.set_radii(h2, dtr)
for blk in (1024, 2048, 4736, 9472):
    for _ in range(3):
        lib.sweep_resident(0, n, blk, 400, 350, 0)
    ms = min(lib.sweep_resident(0, n, blk, 400, 350, 0) for _ in range(5))
    out[f"sweep_gint_s_block{blk}"] = float(n) * n / ms * 1e-6
    print(f"resident sweep N={n} block {blk}: {ms * 1e3:7.1f} us = {float(n) * n / ms * 1e-6:7.1f} Gint/s ({float(n) * n / ms * 1e-6 / 1240.8 * 100:4.1f} % of the FP32 roofline)", flush=True)
lib.close()
if len(sys.argv) > 2:
    Path(sys.argv[2]).write_text(json.dumps(out, indent=1))
