import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["GPUNB_B200_MULTI"] = "1"; os.environ["GPU_LIST"] = "0 1"; os.environ.setdefault("OMP_NUM_THREADS", "8")
import numpy as np
import oracle_lib
from nbody6ppgpu_b200 import load, snapshots as S, sharding
lib = load(); lib.devinit(0)
o = oracle_lib.Oracle()
n = 20011
m, x, v = S.plummer(n, 31, "kroupa")
ref = o.pot_f64(1, n, m, x)
for rep in range(2):
    phi = lib.gpupot(1, n, m, x)
    err = np.abs(phi - ref) / ref
    print("rep", rep, "max err", err.max(), "n bad", int((err > 1e-6).sum()), "first bad", np.nonzero(err > 1e-6)[0][:8], flush=True)
    bad = np.nonzero(err > 1e-6)[0]
    if bad.size:
        i = bad[0]
        print(" phi", phi[i], "ref", ref[i], "ratio", phi[i] / ref[i])
        # partial sums of the oracle per shard
        for r in range(2):
            mem = sharding.shard_members(x, r, 2)
            d = x[mem] - x[i]; rr = np.sqrt((d * d).sum(1)); ok = rr > 0
            print("  shard", r, "partial", float((m[mem][ok] / rr[ok]).sum()))
