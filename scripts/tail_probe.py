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
lib.set_tuning(0, int(os.environ.get("PROBE_NSUB", "1")))      # one pair-kernel launch per call: the isolated launch of roofline.frac
for b in range(3):
    lib.regf(h2[b*1024:(b+1)*1024], dtr[b*1024:(b+1)*1024], x[b*1024:(b+1)*1024], v[b*1024:(b+1)*1024], 600, 550, 0)
buf = np.zeros(3 * 4096, dtype=np.uint64)
lib.lib.gpunb_b200_debug_wtimes.argtypes = [C.c_void_p, C.c_int]
k = lib.lib.gpunb_b200_debug_wtimes(buf.ctypes.data, 4096)
t = buf.reshape(-1, 3)
t0 = (t[:, 0]).astype(np.int64); t1 = t[:, 1].astype(np.int64); near = (t[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
smid = (t[:, 2] >> np.uint64(32)).astype(np.int64)
ok = t0 > 0
t0, t1, near, smid = t0[ok], t1[ok], near[ok], smid[ok]
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
# per SM: items resident, first start, last end, mean duration -- is the spread between SMs or inside them?
sms = np.unique(smid)
per = np.array([[np.sum(smid == s_), (t0[smid == s_].min() - start) * 1e-3, (t1[smid == s_].max() - start) * 1e-3, dur[smid == s_].mean()] for s_ in sms])
print("SMs used", sms.size, "items per SM: min", per[:, 0].min(), "max", per[:, 0].max())
print("per-SM first start us: min %.1f max %.1f | last end us: min %.1f mean %.1f max %.1f | mean duration us: min %.1f max %.1f" % (
    per[:, 1].min(), per[:, 1].max(), per[:, 2].min(), per[:, 2].mean(), per[:, 2].max(), per[:, 3].min(), per[:, 3].max()))
print("within-SM duration spread (max-min) us: mean %.1f" % np.mean([dur[smid == s_].max() - dur[smid == s_].min() for s_ in sms]))
order = np.argsort(per[:, 2])
print("slowest SMs (id, items, start, end, mean dur):", [(int(sms[o]), int(per[o, 0]), round(per[o, 1], 1), round(per[o, 2], 1), round(per[o, 3], 1)) for o in order[-6:]])
print("fastest SMs:", [(int(sms[o]), int(per[o, 0]), round(per[o, 1], 1), round(per[o, 2], 1), round(per[o, 3], 1)) for o in order[:6]])
c = lib.counters(); print({k: c[k] for k in ("grav_ms", "grav_launches")})
lib.close()
