"""One process driving G GPUs (GPUNB_B200_MULTI=1: the mode an unmodified, non-MPI NBODY6++ binary uses, like the reference's
gpunb.velocity.cu with its OpenMP thread per GPU): microseconds per gpunb_regf_ call and Gint/s at N = 1M.
Usage: python scripts/inproc_probe.py G [N]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
G = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
os.environ["GPUNB_B200_MULTI"] = "1"
os.environ["GPU_LIST"] = " ".join(str(g) for g in range(G))
import numpy as np
from nbody6ppgpu_b200 import load, snapshots as S
lib = load(); lib.devinit(0)
assert lib.num_devices() == G
m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0)
lib.open(n + 10, 0)
for pin in (0, 1):
    call = lib.block_caller(h2, dtr, x, v, 2048, 600, 550, 0)
    arrs = [m, x, v, *call.outputs]
    if pin:
        assert lib.pin_host(*arrs)
    sends = {}
    for mode, thr in (("whole snapshot to every device", -1), ("slice per device + NVLink pushes", 0)):
        lib.set_send_scatter(thr)
        lib.send(m, x, v); lib.send(m, x, v)            # first sends of this arm: allocations, sort buffers
        t0 = time.perf_counter()
        for _ in range(6):
            lib.send(m, x, v)
        ts = (time.perf_counter() - t0) / 6
        a = [q.copy() for q in lib.regf(h2[:256], dtr[:256], x[:256], v[:256], 600, 550, 0)]
        sends[mode] = (ts, a)
        print(f"inproc x{G} {'pinned  ' if pin else 'pageable'} gpunb_send_ {mode:34s}: {ts * 1e3:7.3f} ms", flush=True)
    ra, rb = sends["whole snapshot to every device"][1], sends["slice per device + NVLink pushes"][1]
    assert all(np.array_equal(p_, q_) for p_, q_ in zip(ra, rb)), "gpunb_regf_ results differ between the two send paths"
    for ni in (1024, 2048, 64):
        for b in range(4):
            call(b * ni, ni)
        nb = 48
        lib.reset_counters()
        t0 = time.perf_counter()
        for b in range(nb):
            call((4 + b) * ni, ni)
        t = (time.perf_counter() - t0) / nb
        c = lib.counters()
        print(f"inproc x{G} {'pinned  ' if pin else 'pageable'} ni {ni:5d}: {t * 1e6:8.1f} us per gpunb_regf_ call = {ni * float(n) / t * 1e-9:8.1f} Gint/s "
              f"(pair kernel on device 0 {c['grav_ms'] / nb * 1e3:7.1f} us, merge+combine {c['merge_ms'] / nb * 1e3:6.1f} us, scatter {c['host_scatter_ms'] / nb * 1e3:5.1f} us); gpunb_send_ {ts * 1e3:6.2f} ms", flush=True)
    if pin:
        lib.unpin_host(*arrs)
lib.close()
