"""Host-side mirror of the multi-GPU exchange step (used by bench.py's launcher logic and by the CPU tests).

The reference splits the j ARRAY into contiguous index ranges (gpunb.velocity.cu:713-715: joff[id] =
id*nbody/numGPU).  This library sorts the j-set along a Hilbert curve, cuts it into tiles of 64 and gives shard r
the tiles r, r+R, r+2R, ... (``shard_tiles``): every tile is as compact as on one GPU, so the tile culling of the
pair kernel works as well on a shard as on the whole set, and every shard samples every region of the cluster.  Every shard produces, per i-particle, partial sums, a signed
neighbour count and an ascending row of global j indices.  ``combine_shards`` restates what ``combine_kernel``
does on the GPU: fp64 sum in rank order, rows gathered and sorted ascending, overflow -> -(sum |count_r|).
"""
from __future__ import annotations

import numpy as np

TJ = 64


def shard_range(rank: int, nranks: int, nj: int) -> tuple[int, int]:
    """The reference's index-range split (kept for comparison and for the gpupot/legacy tests)."""
    return (rank * nj) // nranks, ((rank + 1) * nj) // nranks


def shard_tiles(rank: int, nranks: int, nj: int) -> range:
    """Tiles of shard `rank`: every nranks-th tile of the Hilbert-sorted j-set (mirror of shard_tiles() in
    gpunb_b200.cu)."""
    T = (nj + TJ - 1) // TJ
    return range(rank, T, nranks)


def hilbert_keys(x: np.ndarray) -> np.ndarray:
    """63-bit Hilbert keys (Skilling's transpose algorithm, 21 bits per axis) -- numpy mirror of morton_key() in
    gpunb_b200.cu, same fp32 quantisation."""
    x32 = np.asarray(x, dtype=np.float64).astype(np.float32)
    H = np.float32(max(float(np.abs(x32).max()), 1e-30))
    sc = np.float32(1048575.5) / H
    X = [np.clip(x32[:, c] * sc + np.float32(1048576.0), np.float32(0.0), np.float32(2097151.0)).astype(np.uint64) for c in range(3)]
    M = 1 << 20
    Q = M
    while Q > 1:
        P = np.uint64(Q - 1)
        for i in range(3):
            hit = (X[i] & np.uint64(Q)) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            t = np.where(hit, np.uint64(0), (X[0] ^ X[i]) & P)
            X[0] = X[0] ^ t
            X[i] = X[i] ^ t
        Q >>= 1
    X[1] ^= X[0]
    X[2] ^= X[1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & np.uint64(Q)) != 0, t ^ np.uint64(Q - 1), t)
        Q >>= 1
    X = [v ^ t for v in X]

    def spread(v):
        v = v & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    return (spread(X[0]) << np.uint64(2)) | (spread(X[1]) << np.uint64(1)) | spread(X[2])


def hilbert_bits(n: int) -> int:
    """Bits per axis the library sorts on (mirror of hilbert_bits() in gpunb_b200.cu)."""
    lg = 0
    while (1 << lg) < n:
        lg += 1
    return min(max((lg + 2) // 3 + 5, 8), 21)


def hilbert_order(x: np.ndarray) -> np.ndarray:
    """The library's j order: stable sort on the leading 3*hilbert_bits(n) bits of the 63-bit Hilbert keys."""
    shift = np.uint64(63 - 3 * hilbert_bits(x.shape[0]))
    return np.argsort(hilbert_keys(x) >> shift, kind="stable")


def shard_members(x: np.ndarray, rank: int, nranks: int) -> np.ndarray:
    """Global j indices (ascending) of the particles in shard `rank`."""
    nj = x.shape[0]
    order = hilbert_order(x)
    parts = [order[t * TJ:min((t + 1) * TJ, nj)] for t in shard_tiles(rank, nranks, nj)]
    return np.sort(np.concatenate(parts)) if parts else np.zeros(0, dtype=np.int64)


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
        out[i, 1:1 + total] = np.sort(np.concatenate([lists[r][i, 1:1 + cnts[r]] for r in range(R)]))
    return f, out


def send_slice(rank: int, nranks: int, nj: int) -> tuple[int, int, int]:
    """Scattered gpunb_send_ (lib_send, one process per GPU): rank r uploads particles [lo, hi) of the snapshot into its chunk of
    the gather buffer; chunk = ceil(nj / R) particles per rank (the last chunks may be short or empty).  Returns (lo, hi, chunk)."""
    chunk = (nj + nranks - 1) // nranks
    lo = min(nj, rank * chunk)
    return lo, min(nj, lo + chunk), chunk


def send_pack(rank: int, nranks: int, m, x, v) -> np.ndarray:
    """This rank's 7 * chunk doubles of the gather buffer: m[chunk] | x[chunk][3] | v[chunk][3] of its slice (tail unspecified)."""
    nj = m.shape[0]
    lo, hi, chunk = send_slice(rank, nranks, nj)
    buf = np.full(7 * chunk, np.nan)
    k = hi - lo
    buf[:k] = m[lo:hi]
    buf[chunk:chunk + 3 * k] = x[lo:hi].ravel()
    buf[4 * chunk:4 * chunk + 3 * k] = v[lo:hi].ravel()
    return buf


def send_unpack(gathered: np.ndarray, nranks: int, nj: int):
    """Mirror of send_unpack_kernel: the all-gathered buffer [R][7 chunk] -> the packed snapshot (m[nj], x[nj][3], v[nj][3])."""
    chunk = (nj + nranks - 1) // nranks
    g = np.asarray(gathered).reshape(nranks, 7 * chunk)
    j = np.arange(nj)
    r, o = j // chunk, j % chunk
    m = g[r, o]
    x = np.stack([g[r, chunk + 3 * o + c] for c in range(3)], axis=1)
    v = np.stack([g[r, 4 * chunk + 3 * o + c] for c in range(3)], axis=1)
    return m, x, v


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
