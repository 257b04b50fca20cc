"""Per-kernel device timeline of a resident sweep (GPUNB_B200_TIMELINE=1), one process per GPU under torchrun
or a single process.  Prints mean microseconds per i-block of isort / regf / merge / exchange on every rank."""
import os, sys
from pathlib import Path
os.environ["GPUNB_B200_TIMELINE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
import torch
os.environ["GPU_LIST"] = str(local)
torch.cuda.set_device(local)
dist = None
force = os.environ.get("GPUNB_FORCE_SHARD") == "1"      # exercise the exchange path with a single rank
if world > 1 or force:
    import torch.distributed as dist
    if force and "MASTER_ADDR" not in os.environ:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29555", RANK="0", WORLD_SIZE="1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nbody6ppgpu_b200 import load, snapshots as S
from nbody6ppgpu_b200.gpunb import ForceLib
lib = ForceLib(os.environ["GPUNB_PROBE_LIB"]) if os.environ.get("GPUNB_PROBE_LIB") else load()
lib.devinit(rank)
if dist:
    from nbody6ppgpu_b200.sharding import nccl_bootstrap
    nccl_bootstrap(lib, rank, world)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 128
m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0)
lib.open(n + 10, rank); lib.send(m, x, v); lib.set_radii(h2, dtr)
lib.sweep_resident(0, 1024 * 16, 1024, 600, 550, 0)
lib.reset_counters()
if dist: dist.barrier()
ms = lib.sweep_resident(0, 1024 * nblk, 1024, 600, 550, 0)
c = lib.counters(); b = c["tl_blocks"]
print(f"rank {rank}/{world}: {ms / nblk * 1e3:7.1f} us per block | isort {c['tl_isort_ms'] / b * 1e3:6.1f}  regf {c['tl_regf_ms'] / b * 1e3:7.1f}  "
      f"merge {c['tl_merge_ms'] / b * 1e3:6.1f}  exchange {c['tl_exch_ms'] / b * 1e3:6.1f}  (sum {(c['tl_isort_ms'] + c['tl_regf_ms'] + c['tl_merge_ms'] + c['tl_exch_ms']) / b * 1e3:7.1f})", flush=True)
lib.close()
if dist:
    dist.barrier(); lib.nccl_finalize(); dist.destroy_process_group()
