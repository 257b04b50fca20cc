"""Tuning run: resident-sweep throughput of every regf kernel variant (one process per variant)."""
import os, subprocess, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import numpy as np
    from nbody6ppgpu_b200 import load, snapshots as S
    import oracle_lib
    n = int(sys.argv[2])
    lib = load(); lib.devinit(0)
    # parity spot check
    m, x, v = S.plummer(4099, 1, "kroupa"); h2, dtr = S.radii(x, m, S.rs0_for_nnb(4099, 60.0))
    lib.open(4200, 0); lib.send(m, x, v)
    acc, jrk, pot, lst = lib.regf(h2[:700], dtr[:700], x[:700], v[:700], 400, 350, 0)
    lib.close()
    o = oracle_lib.Oracle()
    a64, j64, p64, l64, band, _ = o.regf_f64(m, x, v, h2[:700], dtr[:700], x[:700], v[:700], 400, 350, 0)
    ok = not oracle_lib.list_rows_equal(lst, l64)
    errs = (oracle_lib.relerr(acc, a64), oracle_lib.relerr_scaled(jrk, j64, o.scale[:, 1]), oracle_lib.relerr(jrk, j64), oracle_lib.relerr(pot, p64))
    m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0)
    lib.open(n + 10, 0); lib.send(m, x, v); lib.set_radii(h2, dtr)
    nis = min(n, 131072)
    best = 0
    for rep in range(4):
        ms = lib.sweep_resident(0, nis, 1024, 600, 550, 0)
        best = max(best, nis * n / ms * 1e-6)
    c = lib.counters()
    nearfrac = c["near_tiles"] / c["all_tiles"] if c.get("all_tiles", 0) > 0 else None
    small = {}
    for ni in (32, 128, 512):
        lib.reset_counters()
        for r in range(20):
            lib.regf(h2[:ni], dtr[:ni], x[:ni], v[:ni], 600, 550, 0)
        c = lib.counters()
        small[ni] = ni * n * 20 / c["grav_ms"] * 1e-6
    lib.close()
    print(json.dumps({"variant": os.environ.get("GPUNB_B200_VARIANT", "default"), "n": n, "gint_s": best, "lists_exact": ok,
                      "err_acc_jrkS_jrk_pot": errs, "small_ni_kernel_gints": small, "near_frac": nearfrac}))
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "262144"
    for var in sys.argv[2:] or ["it2", "it2b3", "it1", "it1b5", "it1b6"]:
        env = dict(os.environ, GPUNB_B200_VARIANT=var)
        r = subprocess.run([sys.executable, __file__, "--child", n], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("FAILED " + var + r.stderr[-500:]), flush=True)
