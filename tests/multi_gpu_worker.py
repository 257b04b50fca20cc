"""Worker for tests/test_multi_gpu.py.  Modes:
   inproc G   one process driving G GPUs (GPUNB_B200_MULTI=1), checked against the oracle
   nccl       launched under torchrun: one process per GPU joined with gpunb_b200_nccl_init
Exit code 0 on success; prints one summary line."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("OMP_NUM_THREADS", "8")
import numpy as np  # noqa: E402


def check_pinned(lib, tag, rank=0):
    """Caller-pinned arrays (gpunb_b200_pin_host_) through the exchange step: combine_kernel writes the ABI layout straight
    into the caller's arrays; bit for bit the staged path."""
    from nbody6ppgpu_b200 import snapshots as S
    n = 20011
    m, x, v = S.plummer(n, 31, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 150.0))
    lib.open(n + 10, rank)
    staged, pinned = {}, []
    try:
        for mode in ("staged", "pinned"):
            call = lib.block_caller(h2, dtr, x, v, 2048, 400, 350, 0)
            if mode == "pinned":
                pinned = [m, x, v, *call.outputs]
                assert lib.pin_host(*pinned)
            lib.send(m, x, v)
            for nsub, i0, ni in ((1, 0, 1024), (-2, 5000, 2048), (1, 300, 20)):
                lib.set_tuning(0, nsub)
                res = [a.copy() for a in call(i0, ni)]
                if mode == "staged":
                    staged[(nsub, i0, ni)] = res
                else:
                    for q in range(4):
                        assert np.array_equal(res[q], staged[(nsub, i0, ni)][q]), (tag, nsub, i0, ni, q)
    finally:
        if pinned:
            lib.unpin_host(*pinned)
        lib.set_tuning(0, 4)
        lib.close()
    print(f"{tag} rank {rank}: pinned ok", flush=True)


def check(lib, tag, rank=0):
    import oracle_lib
    from nbody6ppgpu_b200 import snapshots as S
    o = oracle_lib.Oracle()
    n = 20011
    m, x, v = S.plummer(n, 31, "kroupa")
    for m_flag, lmax, nnbmax in ((0, 400, 350), (1, 400, 350), (0, 128, 60)):
        h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 150.0), 0.125, m_flag)
        lib.open(n + 10, rank)
        lib.send(m, x, v)
        for isel in (slice(0, 1024), slice(n - 700, n), slice(5000, 5003)):
            lib.set_tuning(0, -2 if isel.start == 0 else 4)      # first block: forced sub-blocks (two exchange steps per call)
            acc, jrk, pot, lst = lib.regf(h2[isel], dtr[isel], x[isel], v[isel], lmax, nnbmax, m_flag)
            a64, j64, p64, l64, band, _ = o.regf_f64(m, x, v, h2[isel], dtr[isel], x[isel], v[isel], lmax, nnbmax, m_flag, 4.0)
            bad = [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > 4.0]
            assert not bad, (tag, m_flag, bad[:5])
            ea, ep = oracle_lib.relerr(acc, a64), oracle_lib.relerr(pot, p64)
            ej = oracle_lib.relerr_scaled(jrk, j64, o.scale[:, 1])
            assert max(ea, ej, ep) <= 1e-6, (tag, ea, ej, ep)
        # resident sweep path uses the same exchange step, launched back to back
        lib.set_radii(h2, dtr)
        # 12 blocks: more than the XSLOTS = 8 exchange slots, so slot reuse (acks) is exercised
        lib.sweep_resident(0, 12 * 1024, 1024, lmax, nnbmax, m_flag)
        a, j, p, l = lib.fetch_last(lmax)
        lib.set_tuning(0, 1)                        # one pair-kernel launch per call: same summation order as the sweep
        lib.set_isort_pairs(0.0)                    # ... and the same Morton order of the i-block (small calls skip the sort)
        a2, j2, p2, l2 = lib.regf(h2[11264:12288], dtr[11264:12288], x[11264:12288], v[11264:12288], lmax, nnbmax, m_flag)
        lib.set_tuning(0, 4)
        lib.set_isort_pairs(2.5e7)
        assert np.array_equal(a, a2) and np.array_equal(p, p2) and not oracle_lib.list_rows_equal(l, l2)
        lib.close()
    # gpupot over the same j-shards (partials summed in rank order)
    for istart, ni in ((1, n), (4097, 3001)):
        phi = lib.gpupot(istart, ni, m, x)
        ref = o.pot_f64(istart, ni, m, x)
        assert float(np.max(np.abs(phi - ref) / ref)) <= 1e-6, (tag, istart, ni)
    print(f"{tag} rank {rank}: ok", flush=True)


def check_auto_subblocks(lib, tag, rank, world):
    """the number of sub-blocks (= exchange steps) of one gpunb_regf_ call must be the same on every rank even
    though the shards differ by a tile.  n = 20011 gives 313 tiles (157 / 156 at two ranks); with the threshold lowered to
    3e6 pairs, ni = 600 put the rank-local pair counts on either side of a multiple of the threshold (the old rule cut the
    call into 2 sub-blocks on rank 0 and 1 on rank 1: wrong rows or a hang)."""
    import oracle_lib
    from nbody6ppgpu_b200 import snapshots as S
    o = oracle_lib.Oracle()
    n = 20011
    m, x, v = S.plummer(n, 31, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 150.0))
    lib.open(n + 10, rank)
    lib.send(m, x, v)
    lib.set_tuning(0, 4)
    try:
        for thr, ni in ((3.0e6, 600), (1.5e6 * 4 / world, 600), (2.0e6, 900), (1.0e6, 1200)):
            lib.set_sub_pairs(thr)
            isel = slice(100, 100 + ni)
            acc, jrk, pot, lst = lib.regf(h2[isel], dtr[isel], x[isel], v[isel], 400, 350, 0)
            a64, j64, p64, l64, band, _ = o.regf_f64(m, x, v, h2[isel], dtr[isel], x[isel], v[isel], 400, 350, 0, 4.0)
            bad = [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > 4.0]
            assert not bad, (tag, thr, ni, bad[:5])
            assert max(oracle_lib.relerr(acc, a64), oracle_lib.relerr(pot, p64)) <= 1e-6, (tag, thr, ni)
    finally:
        lib.set_sub_pairs(1.5e8)
        lib.close()
    print(f"{tag} rank {rank}: auto sub-blocks ok", flush=True)


def check_islice(lib, tag, rank, world):
    """i-slice mode: gpunb_regf_ as a collective call -- every rank passes ITS OWN i-slice (ragged sizes, one rank empty)
    and receives its own rows; checked against the oracle on the rank's slice, staged and with pinned caller arrays, with
    one and several sub-blocks (exchange steps), for both neighbour criteria."""
    import oracle_lib
    from nbody6ppgpu_b200 import snapshots as S
    o = oracle_lib.Oracle()
    n = 20011
    m, x, v = S.plummer(n, 37, "kroupa")
    sizes_all = [[1024] * world, [1 + (311 * (r + 1)) % 1000 for r in range(world)], [0 if r == world - 1 else 700 + r for r in range(world)],
                 [2048 if r == 0 else 5 for r in range(world)], [3 if r == 1 % world else 0 for r in range(world)], [0] * world]
    for m_flag, lmax, nnbmax in ((0, 400, 350), (1, 128, 60)):
        h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 150.0), 0.125, m_flag)
        lib.open(n + 10, rank)
        lib.set_islice(1)
        lib.send(m, x, v)
        try:
            for case, sizes in enumerate(sizes_all):
                start = 17 * case + sum(sizes[:rank]) + 50 * rank           # disjoint slices, different on every rank
                ni = sizes[rank]
                idx = (start + 3 * np.arange(ni)) % n                          # a strided gather
                for nsub in (1, -3):
                    lib.set_tuning(0, nsub)
                    acc, jrk, pot, lst = lib.regf(h2[idx], dtr[idx], x[idx], v[idx], lmax, nnbmax, m_flag)
                    if ni == 0:
                        continue
                    a64, j64, p64, l64, band, _ = o.regf_f64(m, x, v, h2[idx], dtr[idx], x[idx], v[idx], lmax, nnbmax, m_flag, 4.0)
                    bad = [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > 4.0]
                    assert not bad, (tag, "islice", case, nsub, m_flag, bad[:5])
                    ok = lst[:, 0] >= 0
                    if oracle_lib.list_rows_equal(lst, l64):
                        a64, j64, p64 = o.regf_f64_given_list(m, x, v, x[idx], v[idx], np.where(ok[:, None], lst, l64))
                    ea, ep = oracle_lib.relerr(acc, a64), oracle_lib.relerr(pot, p64)
                    ej = oracle_lib.relerr_scaled(jrk, j64, o.scale[:, 1])
                    assert max(ea, ej, ep) <= 1e-6, (tag, "islice", case, nsub, ea, ej, ep)
            # pinned caller arrays: bit for bit the staged rows
            ni = 300 + 100 * rank
            i0 = 4000 + 1000 * rank
            call = lib.block_caller(h2, dtr, x, v, 2048, lmax, nnbmax, m_flag)
            lib.set_tuning(0, 1)
            staged = [a.copy() for a in call(i0, ni)]
            pinned = list(call.outputs)
            assert lib.pin_host(*pinned)
            try:
                for a in call.outputs:
                    a[...] = 0
                res = [a.copy() for a in call(i0, ni)]
            finally:
                lib.unpin_host(*pinned)
            for q in range(4):
                assert np.array_equal(res[q], staged[q]), (tag, "islice pinned", q)
        finally:
            lib.set_islice(0)
            lib.set_tuning(0, 4)
            lib.close()
    print(f"{tag} rank {rank}: i-slice mode ok", flush=True)


def check_regcor(lib, tag, rank=0):
    """The list bookkeeping after a SHARDED gpunb_regf_: combine_kernel leaves the device copy of the rows that
    gpunb_b200_regcor_last_ reads; same results as the call that uploads the rows, and as the CPU restatement."""
    import oracle_lib
    import regcor_cases as RC
    c = RC.make_case(n_tot=20011, ni=900, ifirst=5, lmax=400, nnb_mean=60.0, seed=13)
    o = oracle_lib.Oracle()
    rows_j = c["index_i"] - c["ifirst"]
    h2 = c["rs2"].copy()
    lib.open(c["m"].shape[0] + 10, rank)
    try:
        lib.send(c["m"], c["x"], c["v"])
        for nsub in (1, -2):
            lib.set_tuning(0, nsub)
            new = lib.regf(h2, np.full(900, 0.01), c["x"][rows_j], c["v"][rows_j], 400, 350, 0)[3].copy()
            assert (new[:, 0] >= 0).all()
            args = (c["index_i"], c["ifirst"], c["n"], c["ntot"])
            tail = (c["rs2"], c["step"], c["smin"], c["nnbmax"], c["freg"], c["fdr"])
            last = lib.regcor(*args, None, c["old"], *tail, last_lmax=400)
            up = lib.regcor(*args, new, c["old"], *tail)
            ora = o.regcor(*args, new, c["old"], c["m"], c["x"], c["v"], *tail)
            for k in ("nbloss", "nbgain", "freg", "fdr", "dfirr", "dfd"):
                assert np.array_equal(last[k], up[k]) and np.array_equal(last[k], ora[k]), (tag, nsub, k)
            for r in range(900):
                assert np.array_equal(last["nlist"][r, :last["nlist"][r, 0] + 1], ora["nlist"][r, :ora["nlist"][r, 0] + 1]), (tag, nsub, r)
            assert last["nbsmin"] == ora["nbsmin"] and ora["nbloss"].sum() > 0 and ora["nbgain"].sum() > 0
    finally:
        lib.set_tuning(0, 4)
        lib.close()
    print(f"{tag} rank {rank}: regcor after a sharded regf ok", flush=True)


def main():
    mode = sys.argv[1]
    if mode == "inproc":
        G = int(sys.argv[2])
        os.environ["GPUNB_B200_MULTI"] = "1"
        os.environ["GPU_LIST"] = " ".join(str(g) for g in range(G))
        from nbody6ppgpu_b200 import load
        lib = load(); lib.devinit(0)
        assert lib.num_devices() == G
        check_pinned(lib, f"inproc x{G}")
        # from here on gpunb_send_ scatters (device g uploads slice g, pushes it to every peer over NVLink; the default only
        # above 40000 particles): staged and caller-pinned sources, then every parity check below runs behind it
        lib.set_send_scatter(1000)
        check_pinned(lib, f"inproc x{G} scattered send")
        if not os.environ.get("WORKER_ONLY_PINNED"):
            check(lib, f"inproc x{G}")
            check_regcor(lib, f"inproc x{G}")
    elif mode == "nccl":
        import torch
        import torch.distributed as dist
        rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        os.environ["GPU_LIST"] = str(local)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from nbody6ppgpu_b200 import load
        from nbody6ppgpu_b200.sharding import nccl_bootstrap
        lib = load(); lib.devinit(rank)
        nccl_bootstrap(lib, rank, world)
        check_pinned(lib, f"nccl x{world}", rank)
        # from here on gpunb_send_ uploads 1/R of the snapshot per rank and all-gathers it over NVLink (the default only above
        # 40000 particles): staged and caller-pinned sources, then every parity check below runs behind it
        lib.set_send_scatter(1000)
        check_pinned(lib, f"nccl x{world} scattered send", rank)
        if not os.environ.get("WORKER_ONLY_PINNED"):
            check(lib, f"nccl x{world}", rank)      # every rank checks: results are replicated
            check_auto_subblocks(lib, f"nccl x{world}", rank, world)
            check_islice(lib, f"nccl x{world}", rank, world)
            check_regcor(lib, f"nccl x{world}", rank)
        dist.barrier()
        lib.nccl_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
