"""Resident-sweep throughput against j-sets of N/G particles on ONE GPU (what each rank of a G-GPU run computes per
i-block), to size the per-call overheads (isort, merge) that bound strong scaling."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nbody6ppgpu_b200 import load, snapshots as S
import os
from nbody6ppgpu_b200.gpunb import ForceLib
lib = ForceLib(os.environ["GPUNB_PROBE_LIB"]) if os.environ.get("GPUNB_PROBE_LIB") else load()
lib.devinit(0)
for n in [int(a) for a in os.environ.get('GPUNB_PROBE_N', '125000,250000,500000,1000000').split(',')]:
    m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0 * n / 1e6 if len(sys.argv) > 1 else 200.0)
    lib.open(n + 10, 0); lib.send(m, x, v); lib.set_radii(h2, dtr)
    nis = min(n, 65536)
    best = 1e30
    for rep in range(4):
        best = min(best, lib.sweep_resident(0, nis, 1024, 600, 550, 0))
    calls = nis // 1024
    lib.reset_counters()
    for b in range(16):
        lib.regf(h2[b*1024:(b+1)*1024], dtr[b*1024:(b+1)*1024], x[b*1024:(b+1)*1024], v[b*1024:(b+1)*1024], 600, 550, 0)
    c = lib.counters()
    print(f"nj {n:8d}: sweep {nis * n / best * 1e-6:7.1f} Gint/s, {best / calls * 1e3:7.1f} us per i-block of 1024; "
          f"regf_kernel {c['grav_ms'] / 16 * 1e3:7.1f} us, merge {c['merge_ms'] / 16 * 1e3:6.1f} us", flush=True)
    lib.close()
