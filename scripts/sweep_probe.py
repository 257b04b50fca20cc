"""Resident-sweep rate against pipeline slots and sweep length (blocks of 1024 and larger) for the variant chosen by
GPUNB_B200_VARIANT.  Usage: python scripts/sweep_probe.py [N]"""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from nbody6ppgpu_b200 import load, snapshots as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
lib = load(); lib.devinit(0)
m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0)
lib.open(n + 10, 0); lib.send(m, x, v); lib.set_radii(h2, dtr)
var = os.environ.get("GPUNB_B200_VARIANT", "default")
for block in (1024, 2048, 4736):
    for nslot in (1, 2, 3, 4):
        lib.set_tuning(nslot, 0)
        for nblk in (96, 977 * 1024 // block):
            ni = min(n, block * nblk)
            lib.sweep_resident(0, min(ni, block * 16), block, 600, 550, 0)
            ms = min(lib.sweep_resident(0, ni, block, 600, 550, 0) for _ in range(2))
            print(f"{var} block {block} nslot {nslot} blocks {nblk:4d}: {ms / nblk * 1e3:7.1f} us per block  {float(ni) * n / ms * 1e-6:8.1f} Gint/s", flush=True)
lib.close()
