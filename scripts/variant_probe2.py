"""Round-2 kernel variants on one GPU, one process per variant (the variant is chosen once, at gpunb_devinit_):
strict and scaled jerk error, acc / pot error and list parity against the oracle at N = 2048 / 16384 / 1M, the rate of the
resident sweep and of single launches at N = 1M, and the share of NEAR / transposed tile visits.
Usage: python scripts/variant_probe2.py [out.json] [variant ...]"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def child():
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["GPUNB_B200_STATS"] = "1"
    import numpy as np
    from nbody6ppgpu_b200 import load, snapshots as S
    import oracle_lib
    o = oracle_lib.Oracle()
    lib = load(); lib.devinit(0)
    out = {"variant": os.environ.get("GPUNB_B200_VARIANT", "default"), "errors": {}}

    def errs(m, x, v, h2, dtr, idx, lmax, nnbmax, m_flag):
        acc, jrk, pot, lst = lib.regf(h2[idx], dtr[idx], x[idx], v[idx], lmax, nnbmax, m_flag)
        a64, j64, p64, l64, band, _ = o.regf_f64(m, x, v, h2[idx], dtr[idx], x[idx], v[idx], lmax, nnbmax, m_flag, 4.0)
        bad = oracle_lib.list_rows_equal(lst, l64)
        outside = [i for i in bad if band[i] > 4.0]
        ok = lst[:, 0] >= 0
        if bad:
            a64, j64, p64 = o.regf_f64_given_list(m, x, v, x[idx], v[idx], np.where(ok[:, None], lst, l64))
        dj = np.linalg.norm(jrk - j64, axis=1) / np.linalg.norm(j64, axis=1)
        return {"acc": oracle_lib.relerr(acc, a64), "pot": oracle_lib.relerr(pot, p64), "jrk_strict": float(dj.max()),
                "jrk_strict_p99": float(np.quantile(dj, 0.99)), "jrk_strict_median": float(np.median(dj)),
                "jrk_scaled": oracle_lib.relerr_scaled(jrk, j64, o.scale[:, 1]),
                "rows_differing_in_band": len(bad), "rows_differing_outside_band": len(outside)}

    for n, imf, m_flag in ((2048, "equal", 0), (16384, "kroupa", 0), (16384, "kroupa", 1)):
        m, x, v = S.plummer(n, 1, imf)
        h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 100.0), 0.125, m_flag)
        lib.open(n + 10, 0); lib.send(m, x, v)
        for i0 in (0, n - 1024):
            out["errors"][f"N{n}_{imf}_mflag{m_flag}_i{i0}"] = errs(m, x, v, h2, dtr, slice(i0, i0 + 1024), 400, 350, m_flag)
        lib.close()
    n = int(os.environ.get("PROBE_N", "1000000"))
    m, x, v = S.plummer(n, 1, "kroupa")
    h2, dtr = S.radii_nnb(x, m, 200.0)
    r = np.sqrt((x ** 2).sum(1)); order = np.argsort(r)
    rng = np.random.default_rng(3)
    lib.open(n + 10, 0); lib.send(m, x, v)
    for name, idx in (("core", np.sort(order[:256])), ("halo", np.sort(order[-256:])), ("gather", np.sort(rng.choice(n, 256, replace=False)))):
        out["errors"][f"N{n}_{name}"] = errs(m, x, v, h2, dtr, idx, 600, 550, 0)
    # rates
    lib.set_radii(h2, dtr)
    nblk = 96
    lib.sweep_resident(0, 1024 * 16, 1024, 600, 550, 0)
    lib.reset_counters()
    ms = min(lib.sweep_resident(0, 1024 * nblk, 1024, 600, 550, 0) for _ in range(3))
    c = lib.counters()
    out["sweep_gint_s"] = 1024.0 * nblk * n / ms * 1e-6
    out["near_frac"] = c["near_tiles"] / max(c["all_tiles"], 1.0)
    out["transposed_frac"] = c["transposed_tiles"] / max(c["all_tiles"], 1.0)
    lib.set_tuning(0, 1)
    for ni in (1024, 256, 32):
        lib.reset_counters()
        for b in range(24):
            i0 = b * 1024
            lib.regf(h2[i0:i0 + ni], dtr[i0:i0 + ni], x[i0:i0 + ni], v[i0:i0 + ni], 600, 550, 0)
        c = lib.counters()
        out[f"launch_ms_ni{ni}"] = c["grav_ms"] / c["grav_launches"]
        out[f"launch_gint_s_ni{ni}"] = ni * float(n) / (c["grav_ms"] / c["grav_launches"]) * 1e-6
    # a spatially COMPACT i-block (the 1024 particles nearest the centre): every lane of a warp shares its NEAR tiles
    idx = np.sort(order[:1024])
    lib.reset_counters()
    for _ in range(8):
        lib.regf(h2[idx], dtr[idx], x[idx], v[idx], 600, 550, 0)
    c = lib.counters()
    out["launch_ms_compact_core_1024"] = c["grav_ms"] / c["grav_launches"]
    out["compact_near_frac"] = c["near_tiles"] / max(c["all_tiles"], 1.0)
    out["compact_transposed_frac"] = c["transposed_tiles"] / max(c["all_tiles"], 1.0)
    lib.close()
    print("VARIANT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
    else:
        outp = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/variant_probe2.json"
        res = []
        for var in sys.argv[2:] or ["it1b4", "it1b4n", "it1b4t", "it1b4nt"]:
            env = dict(os.environ, GPUNB_B200_VARIANT=var.split("+")[0])
            for opt in var.split("+")[1:]:
                if opt == "x":                     # work items mapped so that the warps of a CTA take different i-tiles
                    env["GPUNB_B200_ITMAP"] = "1"
                if opt.startswith("o"):            # grid oversubscription: o2 = two work items per resident warp slot
                    env["GPUNB_B200_OVERSUB"] = opt[1:]
            r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
            lines = [l for l in r.stdout.splitlines() if l.startswith("VARIANT ")]
            if not lines:
                print("FAILED", var, r.stdout[-1500:], r.stderr[-3000:], flush=True)
                continue
            d = json.loads(lines[-1][8:])
            d["variant"] = var
            res.append(d)
            worst = {k: max(e[k] for e in d["errors"].values()) for k in ("acc", "pot", "jrk_strict", "jrk_scaled")}
            print(var, {k: f"{v:.2e}" for k, v in worst.items()},
                  "rows outside band:", sum(e["rows_differing_outside_band"] for e in d["errors"].values()),
                  f"sweep {d['sweep_gint_s']:.1f} Gint/s, launch ni1024 {d['launch_ms_ni1024']:.4f} ms ({d['launch_gint_s_ni1024']:.1f}), "
                  f"ni256 {d['launch_gint_s_ni256']:.1f}, ni32 {d['launch_gint_s_ni32']:.1f}, compact {d['launch_ms_compact_core_1024']:.4f} ms, "
                  f"near {d['near_frac']:.4f} transposed {d['transposed_frac']:.4f} compact near {d['compact_near_frac']:.4f} tr {d['compact_transposed_frac']:.4f}", flush=True)
        Path(outp).write_text(json.dumps(res, indent=1))
