"""One-shot GPU probe: FP32 pipe microbenchmark + a first timing of the regf kernel (not the bench)."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from nbody6ppgpu_b200 import load, snapshots as S
lib = load(); lib.devinit(0)
names = {0: "FFMA", 1: "FFMA2", 2: "FADD2", 3: "FMUL2", 4: "MUFU.RSQ", 5: "FFMA2+ALU"}
res = {}
for mode in range(6):
    best = max(lib.fp32_microbench(mode, 8192) for _ in range(3))
    res[names[mode]] = best
    print(f"microbench {names[mode]:10s} {best:8.2f} T(fl)op/s", flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
m, x, v = S.plummer(n, 1, "kroupa")
h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 200.0))
lib.open(n + 10, 0)
t = time.time(); lib.send(m, x, v); print("send s", time.time() - t)
lib.set_radii(h2, dtr)
for rep in range(3):
    ms = lib.sweep_resident(0, min(n, 65536), 1024, 600, 550, 0)
    inter = min(n, 65536) * n
    print(f"resident sweep: {ms:.3f} ms  {inter/ms*1e-6:.1f} Gint/s", flush=True)
    res["resident_gints"] = inter / ms * 1e-6
lib.reset_counters()
t = time.time()
for b in range(0, 8192, 1024):
    lib.regf(h2[b:b+1024], dtr[b:b+1024], x[b:b+1024], v[b:b+1024], 600, 550, 0)
dt = time.time() - t
c = lib.counters()
print("abi: wall", dt, "Gint/s", 8192 * n / dt * 1e-9, "kernel-only Gint/s", c["interactions"] / c["grav_ms"] * 1e-6, c)
lib.profile(0)
lib.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(res, open(ROOT / "gpurun_out" / "probe.json", "w"))
