"""gpupot_ at N = 1M: wall-clock and kernel time per call, one process (optionally GPUNB_B200_MULTI=1) or one rank per GPU
under torchrun (partial potentials of the j-shards, all-gather, sum in rank order).  Usage: python scripts/pot_probe.py [N]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
import numpy as np
import torch
if "GPUNB_B200_MULTI" not in os.environ:
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
m, x, v = S.plummer(n, 1, "kroupa")
lib.open(n + 10, rank)
lib.gpupot(1, 4096, m, x)
ref = None
for rep in range(2):
    if dist: dist.barrier()
    lib.reset_counters()
    t0 = time.perf_counter()
    phi = lib.gpupot(1, n, m, x)
    tp = time.perf_counter() - t0
    c = lib.counters()
    print(f"rank {rank}/{world} devices {lib.num_devices()} gpupot_ N={n} call {rep}: wall {tp * 1e3:7.1f} ms, kernels {c['pot_ms']:7.1f} ms = "
          f"{float(n) * n / c['pot_ms'] * 1e-6:7.1f} Gpair/s in total", flush=True)
print(f"rank {rank}: phi[0..2] = {phi[:3]}, sum m phi = {float((m * phi).sum()):.12f}", flush=True)
lib.close()
if dist:
    dist.barrier(); lib.nccl_finalize(); dist.destroy_process_group()
