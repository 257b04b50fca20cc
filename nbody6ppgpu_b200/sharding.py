"""Host-side mirror of the multi-GPU exchange step (used by bench.py's launcher logic and by the CPU tests).

The j-set is split into contiguous index ranges exactly like the reference
(gpunb.velocity.cu:713-715: joff[id] = id*nbody/numGPU); every shard produces, per i-particle, partial
sums, a signed neighbour count and an ascending row of global j indices.  ``combine_shards`` restates what
``combine_kernel`` does on the GPU: fp64 sum in rank order, counts scanned in rank order, rows concatenated
(rank order IS ascending j), overflow -> -(sum |count_r|).
"""
from __future__ import annotations

import numpy as np


def shard_range(rank: int, nranks: int, nj: int) -> tuple[int, int]:
    return (rank * nj) // nranks, ((rank + 1) * nj) // nranks


def combine_shards(f_parts, lists, nnbmax: int):
    """f_parts: list over ranks of [ni,7] float64; lists: list over ranks of [ni,lmax] int32 rows whose
    indices are already GLOBAL.  Returns (f[ni,7], list[ni,lmax])."""
    R = len(f_parts)
    ni, lmax = lists[0].shape
    f = np.zeros((ni, 7))
    for r in range(R):                       # fixed rank order: deterministic
        f += f_parts[r]
    out = np.zeros((ni, lmax), dtype=np.int32)
    for i in range(ni):
        cnts = [int(lists[r][i, 0]) for r in range(R)]
        total = sum(abs(c) for c in cnts)
        if any(c < 0 for c in cnts) or total > nnbmax:
            out[i, 0] = -total
            continue
        out[i, 0] = total
        k = 1
        for r in range(R):
            out[i, k:k + cnts[r]] = lists[r][i, 1:1 + cnts[r]]
            k += cnts[r]
    return f, out


def nccl_bootstrap(lib, rank: int, world: int, device=None):
    """Create the library's NCCL communicator using torch.distributed only to broadcast the unique id."""
    import torch
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.frombuffer(bytearray(lib.nccl_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(uid, src=0)
    lib.nccl_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
