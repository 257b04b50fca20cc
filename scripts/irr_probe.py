"""Irregular-force library probe (for the GPU validation of the rank-3 draft): microseconds per irr_simd_firr_vec_ call
of this repo's libirr_b200.so for active blocks of 1 ... 4096 particles with ~NNB neighbours each, at N particles.
Usage: python scripts/irr_probe.py [N=100000] [NNB=64]
(The reference's AVX library can be timed with the same loop from tests/, where oracle/_ref may be loaded.)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from scipy.spatial import cKDTree
from nbody6ppgpu_b200 import irr, snapshots as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nnb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(1)
m, x, v = S.plummer(n, 1, "kroupa")
a2 = 0.5 * rng.normal(size=(n, 3)); j6 = rng.normal(size=(n, 3)) / 6.0
t0 = np.zeros(n)
idx = cKDTree(x).query(x, k=nnb + 1)[1][:, 1:]
lib = irr.IrrLib(irr.lib_path())
lib.open(n, 8 * ((nnb + 7) // 8) + 8, 0)
t = time.perf_counter()
for i in range(n):
    lib.set_jp(i + 1, x[i], v[i], a2[i], j6[i], m[i], t0[i])
    lib.set_list(i + 1, irr.pad_list(np.sort(idx[i]) + 1))
print(f"set_jp + set_list of {n} particles: {time.perf_counter() - t:.2f} s (python loop)")
for ni in (1, 8, 64, 512, 4096):
    addr = np.sort(rng.choice(n, ni, replace=False)).astype(np.int32) + 1
    lib.firr_vec(0.01, addr)
    reps = 50
    t = time.perf_counter()
    for _ in range(reps):
        acc, jrk, nn = lib.firr_vec(0.01, addr)
    dt = (time.perf_counter() - t) / reps
    print(f"ni {ni:5d}: {dt * 1e6:8.1f} us per irr_simd_firr_vec_ call, {ni * nnb / dt * 1e-9:7.3f} Gint/s")
a64, j64, n64 = irr.firr_f64(0.01, addr, [np.sort(r) + 1 for r in idx], x, v, a2, j6, m, t0)
print("max rel err acc", float(np.max(np.linalg.norm(acc - a64, axis=1) / np.linalg.norm(a64, axis=1))),
      "nearest neighbour ids equal:", bool(np.array_equal(nn, n64)))
lib.profile(0); lib.close(0)
