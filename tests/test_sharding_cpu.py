"""N>1 host logic on CPU: world_size-2 gloo processes, each computing ITS j-shard with the oracle, exchanged
with torch.distributed all_gather and merged by the Python mirror of combine_kernel; the result must equal the
single-process oracle (lists identical, sums equal to fp64 rounding)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from nbody6ppgpu_b200 import snapshots as S  # noqa: E402
from nbody6ppgpu_b200.sharding import combine_shards, send_pack, send_slice, send_unpack, shard_members, shard_range, shard_tiles  # noqa: E402


def test_shard_range_matches_reference_split():
    for nj in (1, 7, 1000, 16385):
        for R in (1, 2, 3, 8):
            edges = [shard_range(r, R, nj) for r in range(R)]
            assert edges[0][0] == 0 and edges[-1][1] == nj
            assert all(edges[r][1] == edges[r + 1][0] for r in range(R - 1))
            assert all(e == ((r * nj) // R, ((r + 1) * nj) // R) for r, e in enumerate(edges))


def test_shard_tiles_partition_the_curve():
    m, x, v = S.plummer(5003, 5, "kroupa")
    T = (5003 + 63) // 64
    for R in (1, 2, 3, 8):
        tiles = [list(shard_tiles(r, R, 5003)) for r in range(R)]
        assert sorted(t for ts in tiles for t in ts) == list(range(T))                   # every tile owned once
        assert max(len(ts) for ts in tiles) - min(len(ts) for ts in tiles) <= 1
        members = [shard_members(x, r, R) for r in range(R)]
        allm = np.concatenate(members)
        assert allm.size == 5003 and np.array_equal(np.sort(allm), np.arange(5003))      # a partition of the j-set
        assert all(np.all(np.diff(mm) > 0) for mm in members)
    # tiles are compact whatever the shard count: 64 consecutive particles of the curve span a small box
    from nbody6ppgpu_b200.sharding import hilbert_keys
    order = np.argsort(hilbert_keys(x), kind="stable")
    inner = [order[t * 64:(t + 1) * 64] for t in range(T - 1)]
    ext = np.array([np.ptp(x[ix], axis=0).max() for ix in inner])
    assert np.median(ext) < 0.25 * np.ptp(x[np.linalg.norm(x, axis=1) < 2.0], axis=0).max()


def _worker(rank, world, port, nnbmax, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    o = oracle_lib.Oracle()
    n, ni, lmax = 3001, 200, 128
    m, x, v = S.plummer(n, 21, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 60.0))
    mem = shard_members(x, rank, world)       # a contiguous range of the Hilbert-sorted tiles, like lib_send
    a, j, p, l, band, _ = o.regf_f64(m[mem], x[mem], v[mem], h2[:ni], dtr[:ni], x[:ni], v[:ni], lmax, nnbmax, 0)
    for i in range(ni):                       # shard-local indices -> global, as regf_kernel does through jidx
        c = l[i, 0]
        if c > 0:
            l[i, 1:1 + c] = mem[l[i, 1:1 + c]]
    f = torch.from_numpy(np.concatenate([a, j, p[:, None]], axis=1))
    lt = torch.from_numpy(l)
    fs = [torch.zeros_like(f) for _ in range(world)]
    ls = [torch.zeros_like(lt) for _ in range(world)]
    dist.all_gather(fs, f); dist.all_gather(ls, lt)
    fc, lc = combine_shards([t.numpy() for t in fs], [t.numpy() for t in ls], nnbmax)
    a1, j1_, p1, l1, _, _ = o.regf_f64(m, x, v, h2[:ni], dtr[:ni], x[:ni], v[:ni], lmax, nnbmax, 0)
    ok = not oracle_lib.list_rows_equal(lc, l1)
    err = max(oracle_lib.relerr(fc[:, 0:3], a1), oracle_lib.relerr(fc[:, 6], p1))
    # every rank holds the complete result (replicated semantics of the NCCL mode)
    res = torch.tensor([1.0 if ok else 0.0, err, float((l1[:, 0] < 0).sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(gathered, res)
    if rank == 0:
        np.save(out, np.stack([g.numpy() for g in gathered]))
    dist.destroy_process_group()


@pytest.mark.parametrize("nnbmax", [100, 40])
def test_two_rank_gloo_shard_combine(tmp_path, nnbmax):
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 2000) + nnbmax
    mp.spawn(_worker, args=(2, port, nnbmax, out), nprocs=2, join=True)
    res = np.load(out)
    assert (res[:, 0] == 1.0).all(), "combined lists differ from the single-process oracle"
    assert (res[:, 1] < 1e-13).all()
    if nnbmax == 40:
        assert res[0, 2] > 0, "the small-nnbmax case must exercise overflow rows"


def test_send_slices_cover_the_snapshot():
    for nj in (1, 7, 64, 1000, 20011, 40000):
        for R in (2, 3, 8):
            sl = [send_slice(r, R, nj) for r in range(R)]
            assert sl[0][0] == 0 and sl[-1][1] == nj and all(sl[r][1] == sl[r + 1][0] for r in range(R - 1))
            assert all(hi - lo <= chunk for lo, hi, chunk in sl) and len({c for _, _, c in sl}) == 1


def _send_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for n in (20011, 64, 5):                   # ragged: 20011 = 2 * 10006 - 1; fewer particles than ranks * chunk
        m, x, v = S.plummer(max(n, 8), 21, "kroupa")
        m, x, v = m[:n], x[:n], v[:n]
        mine = torch.from_numpy(np.nan_to_num(send_pack(rank, world, m, x, v), nan=-7.0))   # every rank has the whole host snapshot,
        parts = [torch.zeros_like(mine) for _ in range(world)]                               # uploads only its slice
        dist.all_gather(parts, mine)
        gm, gx, gv = send_unpack(torch.cat(parts).numpy(), world, n)
        ok &= np.array_equal(gm, m) and np.array_equal(gx, x) and np.array_equal(gv, v)
    res = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64)
    gathered = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(gathered, res)
    if rank == 0:
        np.save(out, np.stack([g.numpy() for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_gloo_scattered_send(tmp_path):
    """The scattered gpunb_send_ of the one-process-per-GPU mode (1/R of the snapshot per rank, all-gather, unpack) with its
    Python mirror over gloo: every rank ends up with the complete packed snapshot, bit for bit."""
    out = str(tmp_path / "send.npy")
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_send_worker, args=(2, port, out), nprocs=2, join=True)
    assert (np.load(out) == 1.0).all()
