import os, sys, ctypes as C
os.environ["GPUNB_B200_STATS"] = "2"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from nbody6ppgpu_b200 import load, snapshots as S
from nbody6ppgpu_b200.gpunb import ForceLib
lib = ForceLib(os.environ["GPUNB_PROBE_LIB"]) if os.environ.get("GPUNB_PROBE_LIB") else load()
lib.devinit(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0 * n / 1e6)
lib.open(n + 10, 0); lib.send(m, x, v)
for b in range(3):
    lib.regf(h2[b*1024:(b+1)*1024], dtr[b*1024:(b+1)*1024], x[b*1024:(b+1)*1024], v[b*1024:(b+1)*1024], 600, 550, 0)
buf = np.zeros(3 * 4096, dtype=np.uint64)
lib.lib.gpunb_b200_debug_wtimes.argtypes = [C.c_void_p, C.c_int]
k = lib.lib.gpunb_b200_debug_wtimes(buf.ctypes.data, 4096)
t = buf.reshape(-1, 3)
t0 = (t[:, 0]).astype(np.int64); t1 = t[:, 1].astype(np.int64); near = t[:, 2].astype(np.int64)
ok = t0 > 0
t0, t1, near = t0[ok], t1[ok], near[ok]
start = t0.min()
dur = (t1 - t0) * 1e-3
print("items", ok.sum(), "kernel span us", (t1.max() - start) * 1e-3)
print("start offset us: max", (t0.max() - start) * 1e-3)
print("duration us: min %.1f mean %.1f max %.1f std %.1f" % (dur.min(), dur.mean(), dur.max(), dur.std()))
print("end time us: min %.1f mean %.1f max %.1f" % ((t1.min() - start) * 1e-3, (t1.mean() - start) * 1e-3, (t1.max() - start) * 1e-3))
print("near tiles per item: min", near.min(), "mean", near.mean(), "max", near.max(), "corr(dur,near)", np.corrcoef(dur, near)[0, 1])
# per-SM-slot view: items are launched 16 per SM; finishing histogram
e = np.sort((t1 - start) * 1e-3)
print("end-time percentiles us:", [round(float(np.percentile(e, q)), 1) for q in (0, 5, 25, 50, 75, 95, 100)])
c = lib.counters(); print(c)
lib.close()
