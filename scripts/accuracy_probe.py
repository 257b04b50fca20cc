"""Accuracy probe on GPU: worst errors over several snapshots/blocks for the current build/env."""
import sys, json, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from nbody6ppgpu_b200 import load, snapshots as S
import oracle_lib
o = oracle_lib.Oracle()
lib = load(); lib.devinit(0)
worst = dict(acc=0, jrkS=0, jrk=0, pot=0); rows = 0; bandrows = 0
for n, imf, m_flag, seed in [(2048, "equal", 0, 1), (16384, "kroupa", 0, 1), (16384, "kroupa", 1, 1), (16384, "kroupa", 1, 2), (65536, "kroupa", 0, 3)]:
    m, x, v = S.plummer(n, seed, imf)
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 100.0), 0.125, m_flag)
    lib.open(n + 10, 0); lib.send(m, x, v)
    for i0 in (0, n // 2, n - 1024):
        sel = slice(i0, i0 + 1024)
        acc, jrk, pot, lst = lib.regf(h2[sel], dtr[sel], x[sel], v[sel], 400, 350, m_flag)
        a64, j64, p64, l64, band, _ = o.regf_f64(m, x, v, h2[sel], dtr[sel], x[sel], v[sel], 400, 350, m_flag, 4.0)
        bad = oracle_lib.list_rows_equal(lst, l64); bandrows += len(bad); rows += 1024
        assert not [i for i in bad if band[i] > 4.0]
        e = dict(acc=oracle_lib.relerr(acc, a64), jrkS=oracle_lib.relerr_scaled(jrk, j64, o.scale[:, 1]), jrk=oracle_lib.relerr(jrk, j64), pot=oracle_lib.relerr(pot, p64))
        for k in worst: worst[k] = max(worst[k], e[k])
        print(n, imf, m_flag, i0, {k: f"{v_:.2e}" for k, v_ in e.items()}, flush=True)
    lib.close()
print("WORST", os.environ.get("GPUNB_B200_FLUSH"), os.environ.get("GPUNB_B200_VARIANT"), {k: f"{v_:.2e}" for k, v_ in worst.items()}, "rows differing (in band):", bandrows, "of", rows)
