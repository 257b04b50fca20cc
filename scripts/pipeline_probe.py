"""Pipeline depth probe on one GPU (or one rank per GPU under torchrun): device-resident sweep rate for
nslot = 1..4 pipeline slots, and microseconds per gpunb_regf_ call (host arrays) for nsub = 1..4 sub-blocks.
Usage: python scripts/pipeline_probe.py [N] [blocks]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
import numpy as np
import torch
os.environ["GPU_LIST"] = str(local)
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nbody6ppgpu_b200 import load, snapshots as S
lib = load(); lib.devinit(rank)
if dist:
    from nbody6ppgpu_b200.sharding import nccl_bootstrap
    nccl_bootstrap(lib, rank, world)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 192
m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0)
lib.open(n + 10, rank); lib.send(m, x, v); lib.set_radii(h2, dtr)
lib.reset_counters()
for _ in range(5):
    lib.send(m, x, v)
c = lib.counters()
print(f"rank {rank}/{world} gpunb_send_ N={n}: {c['send_ms'] / c['sends']:6.2f} ms per call (pinned staging + upload loop {c['send_stage_ms'] / c['sends']:6.2f} ms, "
      f"tile construction on the device {c['send_tiles_ms'] / c['sends']:6.2f} ms)", flush=True)
z3 = np.zeros((n, 3))
lib.state_all(m, x, v, z3, z3, np.zeros(n))
lib.reset_counters()
t0 = time.perf_counter()
for k in range(5):
    lib.predict_send(n, 1e-3 * (k + 1))
tp = (time.perf_counter() - t0) / 5
idx = np.arange(0, n, n // 2000, dtype=np.int32)[:2000]
t0 = time.perf_counter()
for k in range(5):
    lib.state_update(idx, m[idx], x[idx], v[idx], z3[idx], z3[idx], np.zeros(idx.size))
tu = (time.perf_counter() - t0) / 5
c = lib.counters()
print(f"rank {rank}/{world} gpunb_b200_predict_send_ N={n}: {tp * 1e3:6.2f} ms per call (device tile construction {c['send_tiles_ms'] / c['sends']:6.2f} ms); "
      f"state_update of {idx.size} particles {tu * 1e3:6.3f} ms", flush=True)
lib.set_resort_every(8)
lib.predict_send(n, 0.0)
lib.reset_counters()
t0 = time.perf_counter()
for k in range(7):
    lib.predict_send(n, 1e-3 * (k + 1))
tp = (time.perf_counter() - t0) / 7
c = lib.counters()
print(f"rank {rank}/{world} gpunb_b200_predict_send_ with the Hilbert order kept (GPUNB_B200_RESORT_EVERY=8): {tp * 1e3:6.2f} ms per call "
      f"(device tile construction {c['send_tiles_ms'] / c['sends']:6.2f} ms)", flush=True)
lib.set_resort_every(0)
lib.send(m, x, v)
if os.environ.get("PROBE_ONLY_SEND"):
    lib.close()
    sys.exit(0)
for nslot in (1, 2, 3, 4, 3):
    lib.set_tuning(nslot, 0)
    lib.sweep_resident(0, 1024 * 16, 1024, 600, 550, 0)
    if dist: dist.barrier()
    ms = min(lib.sweep_resident(0, 1024 * nblk, 1024, 600, 550, 0) for _ in range(2))
    print(f"rank {rank}/{world} nslot {nslot}: {ms / nblk * 1e3:7.1f} us per block  {1024.0 * nblk * n / ms * 1e-6:8.1f} Gint/s", flush=True)
for ne in ((1, 0) if lib.lib.gpunb_b200_has_near_scalar_ab() else (0,)):
    lib.set_near_exact(ne)
    lib.sweep_resident(0, 1024 * 16, 1024, 600, 550, 0)
    ms = min(lib.sweep_resident(0, 1024 * nblk, 1024, 600, 550, 0) for _ in range(2))
    print(f"rank {rank}/{world} NEAR tiles {'scalar pair body' if ne else 'packed f32x2 pair body'}: {ms / nblk * 1e3:7.1f} us per block  {1024.0 * nblk * n / ms * 1e-6:8.1f} Gint/s", flush=True)
lib.set_near_exact(-1)
if os.environ.get("GPUNB_B200_STATS"):
    lib.reset_counters()
    lib.sweep_resident(0, 1024 * 64, 1024, 600, 550, 0)
    c = lib.counters()
    print(f"rank {rank}/{world} tile visits: NEAR {c['near_tiles'] / c['all_tiles'] * 100:5.2f} %", flush=True)
call = lib.block_caller(h2, dtr, x, v, 1024, 600, 550, 0)
for nsub, taper in ((1, 1), (2, 0), (2, 1), (3, 0), (3, 1), (4, 0), (4, 1)):
    lib.set_tuning(0, nsub); lib.set_taper(taper)
    for rep in range(2):
        if dist: dist.barrier()
        torch.cuda.synchronize()
        lib.reset_counters()
        t0 = time.perf_counter()
        for b in range(48):
            call(b * 1024, 1024)
        t = time.perf_counter() - t0
    c = lib.counters()
    print(f"rank {rank}/{world} nsub {nsub} {'tapering' if taper else 'equal   '}: {t / 48 * 1e6:7.1f} us per gpunb_regf_ call  {1024.0 * 48 * n / t * 1e-9:8.1f} Gint/s | host us/call: "
          f"pack {c['host_pack_ms'] / 48 * 1e3:5.1f} enqueue {c['host_enqueue_ms'] / 48 * 1e3:5.1f} wait {c['host_wait_ms'] / 48 * 1e3:6.1f} "
          f"scatter {c['host_scatter_ms'] / 48 * 1e3:5.1f} | device us/call: pair kernels {c['grav_ms'] / 48 * 1e3:6.1f} tail {c['merge_ms'] / 48 * 1e3:5.1f}", flush=True)
lib.reset_counters()
t0 = time.perf_counter()
phi = lib.gpupot(1, n, m, x)
tp = time.perf_counter() - t0
c = lib.counters()
print(f"rank {rank}/{world} gpupot_ N={n}: {tp * 1e3:7.1f} ms per call, kernels {c['pot_ms']:7.1f} ms = {float(n) * n / world / c['pot_ms'] * 1e-6:7.1f} Gpair/s per GPU", flush=True)
lib.close()
if dist:
    dist.barrier(); lib.nccl_finalize(); dist.destroy_process_group()
