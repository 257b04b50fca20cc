"""gpunb_send_ under torchrun: whole snapshot per rank against the scattered upload + all-gather.  Usage: torchrun ... send_probe.py [N]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
import numpy as np
import torch
import torch.distributed as dist
os.environ["GPU_LIST"] = str(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nbody6ppgpu_b200 import load, snapshots as S
from nbody6ppgpu_b200.sharding import nccl_bootstrap
lib = load(); lib.devinit(rank)
nccl_bootstrap(lib, rank, world)
for n in (int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 262144, 65536):
    m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0)
    lib.open(n + 10, rank)
    res = {}
    for pin in (False, True):
        if pin:
            assert lib.pin_host(m, x, v)
        for mode, thr in (("whole snapshot on every rank", -1), ("1/R per rank + all-gather", 0)):
            lib.set_send_scatter(thr)
            lib.send(m, x, v); lib.send(m, x, v)
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(10):
                lib.send(m, x, v)
            t = (time.perf_counter() - t0) / 10
            acc, jrk, pot, lst = lib.regf(h2[:256], dtr[:256], x[:256], v[:256], 600, 550, 0)
            res[(pin, mode)] = (t, acc.copy(), lst.copy())
            if rank == 0:
                print(f"world {world} N={n} {'pinned  ' if pin else 'pageable'} gpunb_send_ {mode:30s}: {t * 1e3:7.3f} ms", flush=True)
        if pin:
            lib.unpin_host(m, x, v)
    a = res[(False, "whole snapshot on every rank")]; b = res[(False, "1/R per rank + all-gather")]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), "results differ between the two send paths"
    lib.close()
dist.barrier(); lib.nccl_finalize(); dist.destroy_process_group()
