#!/usr/bin/env python
"""bench.py -- regular-force interactions/s (Gint/s) on a synthetic Plummer sphere, N=1M Kroupa IMF.

Metric and config are BASELINE.json's: `regular-force interactions/sec (Gint/s) at N=1M`, workload
`synthetic Plummer N=1M Kroupa IMF, regular-force sweep` (configs[4]; the largest single-GPU config).
A "step" is one FPOLY0-style pass (reference: src/Main/fpoly0.F:72-125): gpunb_send of all N
j-particles, then gpunb_regf over i-blocks of 1024 until `--ni-total` i-particles are done (default:
all N -> 977 calls, 1e12 interactions).  interactions = sum ni*nj exactly as the reference counts
(gpunb.velocity.cu:747), self and neighbour pairs included.

  value   device-resident leg: the j snapshot, radii and i-blocks already in HBM; the blocks cycle through the
          library's pipeline slots (pair kernel of block b+1 beside merge/exchange of block b), timed with CUDA
          events on the library's main stream around the whole sweep (gpunb_b200_sweep_resident).
  e2e     the same sweep through the reference-facing C-ABI (gpunb_send_ + gpunb_regf_) with HOST buffers:
          H2D of the snapshot and of every i-block, D2H of forces and neighbour lists inside the timed
          region.  Default: caller-owned arrays pinned ONCE with gpunb_b200_pin_host_ (what a Fortran
          caller does for its COMMON blocks); the same leg with plain pageable arrays is reported beside
          it (`e2e.pageable`).  Under torchrun the ranks use the library's i-slice mode (the calling
          pattern of NBODY6++'s MPI build, intgrt.F:982-1231): every rank passes its own blocks of 1024
          and receives its own rows.
  parity_check  one block of 1024 i-particles through gpunb_regf_ against the oracle (oracle/, the checker)
          BEFORE the timed region: lists exact outside the 4-ulp band, acc / pot <= 1e-6, strict and scaled
          jerk error reported; at world > 1 also bitwise equality of the replicated results across ranks and
          the i-slice mode against the oracle.  A failure exits non-zero.
  roofline  FP32-FMA bound (no tensor cores: pairwise sum, not a contraction).  achieved = 60 flop x
          interactions of one launch / mean launch duration of regf_kernel (CUDA events around each
          launch, on its stream); peak = 2*128*148*sm_max_mhz nominal AND the FFMA rate measured by the
          library's microbenchmark in the same run.  HBM GB/s is reported only to show the kernel is far
          from memory bound.
  cpu_baseline  the reference's own AVX library (oracle/_ref, kind "reference") on the host cores, on a
          bounded sample of the same workload.
  time_unit / wall_s_per_time_unit  the second half of BASELINE.json's metric on a bounded sample: 1/8 N-body time unit at
          N = 16 000 (samples/N16k.input settings) behind this library with the native Ahmad-Cohen driver
          (csrc/ac_driver.cpp), wall seconds per time unit, buckets per library call, dE/E.  `--time-unit` integrates
          a whole time unit behind this library (device paths / reference ABI only) and both reference libraries.

`--impl reference` times that AVX library alone (same metric/config), each step a bounded sample.
Under torchrun (N>1) the j-set is sharded over ranks (every R-th Hilbert tile); partial sums and neighbour rows are
exchanged by the library's own kernels over NVLink peer memory (flags + peer pulls); NCCL only
bootstraps the cudaIpc handles.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "regular-force interactions/sec at N=1M"
UNIT = "Gint/s"
LMAX, NNBMAX, BLOCK = 600, 550, 1024          # --with-par=1m: LMAX=600 (configure.ac:390-394); NNBMAX=min(N/2,LMAX-50)
NNB_TARGET = 200.0
FLOP_PER_INT = 60.0                           # reference convention, gpunb.velocity.cu:894
NSLOT = int(os.environ.get("GPUNB_B200_NSLOT", "0"))     # pipeline slots of the resident sweep (0: library default, 3)
NSUB = int(os.environ.get("GPUNB_B200_NSUB", "4"))       # sub-blocks of one gpunb_regf_ call


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--ni-total", type=int, default=0, help="i-particles swept per step (0 = all N)")
    ap.add_argument("--cpu-blocks", type=int, default=256, help="i-blocks of 1024 in the CPU baseline sample (~12 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--m-flag", type=int, default=0)
    ap.add_argument("--pageable", action="store_true", help="headline e2e leg with pageable caller arrays (default: the caller pins "
                    "its static arrays once with gpunb_b200_pin_host_; the other kind is reported beside it)")
    ap.add_argument("--quick", action="store_true", help="skip the secondary legs (pageable e2e, other configs, reference CUDA library)")
    ap.add_argument("--ref-cuda-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--time-unit", action="store_true", help="second half of the metric alone: wall s per N-body time unit of the "
                    "Ahmad-Cohen driver at N=16k (samples/N16k.input: NNBOPT=100, KZ(39)=2) behind this library, the reference CUDA "
                    "library and the reference AVX library, with dE/E for each; prints one JSON line")
    ap.add_argument("--time-unit-probe", default="", help=argparse.SUPPRESS)      # (subprocess) one library: b200 | b200_host | ref_cuda | ref_avx
    ap.add_argument("--tu-n", type=int, default=16000)
    ap.add_argument("--tu-t", type=float, default=1.0, help="N-body time units integrated by --time-unit (multiple of --tu-dtmax)")
    ap.add_argument("--tu-dtmax", type=float, default=0.125, help="largest block step (SMAX); the run ends synchronised at multiples of it")
    ap.add_argument("--tu-nnbopt", type=int, default=100)
    ap.add_argument("--tu-lmax", type=int, default=400)
    ap.add_argument("--tu-mflag", type=int, default=1)
    ap.add_argument("--tu-arms", default="b200,b200_host,ref_cuda,ref_avx")
    return ap.parse_args()


def make_snapshot(n, m_flag):
    import numpy as np
    from nbody6ppgpu_b200 import snapshots as S
    m, x, v = S.plummer(n, 1, "kroupa")
    # neighbour spheres with <nnb> ~ NNBOPT = 200 everywhere (the state the RS control converges to)
    h2, dtr = S.radii_nnb(x, m, NNB_TARGET, 0.125, m_flag)
    return m, x, v, h2, dtr, float(np.sqrt(h2.min() * (m.mean() if m_flag else 1.0)))


def bench_config(n, m_flag, rs0, world):
    """The workload, identical for both arms (`--impl b200` and `--impl reference`): what is computed, not how much of it a
    step samples (that is the top-level `sample` / `run`)."""
    return {"workload": f"synthetic Plummer N={n} Kroupa IMF, regular-force sweep", "nj": n, "block": BLOCK, "lmax": LMAX,
            "nnbmax": NNBMAX, "m_flag": m_flag, "rs_min": rs0, "nnb_target": NNB_TARGET,
            "radii": "RS_i from the local Plummer density for <nnb> ~ 200 (snapshots.radii_nnb), dtr as fpoly0.F:53-56",
            "l2": "flushed between timed steps (256 MB fill) on the GPU arm; the j-set (56 MB fp64) exceeds the host caches on the CPU arm",
            "parallelism": f"j-shard x{world}"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w": sorted(pw)[len(pw) // 2],
                "samples": len(sm), "reasons": sorted(reasons)}


_THREADS = None


def cpu_threads():
    """Host threads for the AVX baseline, read ONCE (OMP_PROC_BIND later narrows the main thread's affinity)."""
    global _THREADS
    if _THREADS is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        _THREADS = max(1, min(n, 32))         # reg.avx.cpp:7,103 asserts threads <= TMAX = 32
    return _THREADS


def load_reference_avx():
    """The reference's own CPU library (oracle/_ref): only used as the baseline being timed."""
    from nbody6ppgpu_b200.gpunb import ForceLib
    so = ROOT / "oracle" / "_ref" / "libgpunb_ref_avx.so"
    if not so.exists():
        return None
    return ForceLib(so)


def time_reference(ref, m, x, v, h2, dtr, blocks, m_flag, first_block=0):
    """One bounded sample: send + `blocks` regf calls of 1024 on the AVX library.  Returns (s, interactions)."""
    n = m.shape[0]
    call = ref.block_caller(h2, dtr, x, v, BLOCK, LMAX, NNBMAX, m_flag)   # static caller arrays, as for the b200 arm
    t0 = time.perf_counter()
    ref.send(m, x, v)
    inter = 0
    for b in range(blocks):
        i0 = ((first_block + b) * BLOCK) % max(n - BLOCK - 8, 1)      # the AVX library reads 3 rows past ni
        call(i0, BLOCK)
        inter += BLOCK * n
    return time.perf_counter() - t0, inter


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    threads = cpu_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    ref = load_reference_avx()
    if ref is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libgpunb_ref_avx.so not built"})
        return
    m, x, v, h2, dtr, rs0 = make_snapshot(args.n, args.m_flag)
    ref.open(args.n + 10, 0)
    blocks = max(1, min(args.cpu_blocks // 4, 64))
    for w in range(args.warmup):
        time_reference(ref, m, x, v, h2, dtr, 1, args.m_flag)
    t_tot, inter_tot = 0.0, 0
    for k in range(args.steps):
        t, inter = time_reference(ref, m, x, v, h2, dtr, blocks, args.m_flag, first_block=k * blocks)
        t_tot += t; inter_tot += inter
    ref.close()
    val = inter_tot / t_tot * 1e-9
    sample = f"send + {blocks} regf calls of {BLOCK} i-particles against all {args.n} j per step"
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_tot / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.n, args.m_flag, rs0, args.gpus), "sample": sample,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


_JSON_FD = None


def quiet_stdout():
    """Native libraries (NCCL's version banner, the force libraries' '# Open ...' lines) write to fd 1; the driver wants
    ONE JSON line on stdout.  Everything but emit() goes to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, line)


def sweep_block(lib, world):
    """i-particles per launch group of the resident sweep: 1024 on one GPU.  With the j-set sharded over R ranks the block
    grows with R, so that one launch keeps about the same number of pairs and its fixed costs stay amortised:
    32 * (W / S) with W the resident warps and S the largest integer <= W / (32 R) that divides W (2048 / 4736 / 9472
    at R = 2 / 4 / 8 on a B200) -- every resident warp gets a work item."""
    if world == 1:
        return BLOCK
    W = lib.resident_warps()
    S = max(1, W // (32 * world))
    while S > 1 and W % S:
        S -= 1
    return min(32 * (W // S), 16384)


def valid_rows_bytes(lst):
    import numpy as np
    return b"".join(lst[i, :max(int(lst[i, 0]), 0) + 1].tobytes() for i in range(lst.shape[0]))


def parity_check(lib, dist, rank, world, m, x, v, h2, dtr, m_flag):
    """One block of 1024 i-particles through gpunb_regf_ BEFORE the timed region, against the oracle (the checker under
    oracle/, on rank 0): lists exact outside the 4-ulp band of the RS boundary, acc / pot <= 1e-6, jerk reported strictly
    and under the cancellation-aware norm of tests/test_regf_gpu.py.  world > 1: the replicated results must be bitwise
    equal on every rank, and the i-slice mode (rank r passes rows [r ni / R, (r+1) ni / R) of the same block) must give
    the same lists and forces within the same bars."""
    import hashlib
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    i0, ni = 4096, BLOCK
    sel = slice(i0, i0 + ni)
    lib.set_tuning(0, NSUB)
    rep = [a.copy() for a in lib.regf(h2[sel], dtr[sel], x[sel], v[sel], LMAX, NNBMAX, m_flag)]
    out = {"block_i0": i0, "ni": ni, "tolerance": 1e-6, "band_ulp": 4.0}
    gathered = None
    if world > 1:
        import torch
        dig = hashlib.sha256(rep[0].tobytes() + rep[1].tobytes() + rep[2].tobytes() + valid_rows_bytes(rep[3])).digest()
        t = torch.tensor(list(dig[:16]), dtype=torch.int64, device="cuda")
        alld = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(alld, t)
        out["replicated_bitwise_equal_across_ranks"] = bool(all(torch.equal(alld[0], q) for q in alld))
        lo, hi = i0 + rank * ni // world, i0 + (rank + 1) * ni // world
        lib.set_islice(1)
        mine = [a.copy() for a in lib.regf(h2[lo:hi], dtr[lo:hi], x[lo:hi], v[lo:hi], LMAX, NNBMAX, m_flag)]
        lib.set_islice(0)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        o = oracle_lib.Oracle()
        a64, j64, p64, l64, band, nband = o.regf_f64(m, x, v, h2[sel], dtr[sel], x[sel], v[sel], LMAX, NNBMAX, m_flag, 4.0)
        scale = o.scale[:, 1].copy()

        def judge(res, tag):
            acc, jrk, pot, lst = res
            bad = oracle_lib.list_rows_equal(lst, l64)
            outside = [i for i in bad if band[i] > 4.0]
            a, j, pp = a64, j64, p64
            if bad:
                okrow = lst[:, 0] >= 0
                a, j, pp = o.regf_f64_given_list(m, x, v, x[sel], v[sel], np.where(okrow[:, None], lst, l64))
            dj = np.linalg.norm(jrk - j, axis=1) / np.linalg.norm(j, axis=1)
            r = {"list_rows_differing_in_band": len(bad), "list_rows_differing_outside_band": len(outside),
                 "acc_relerr": oracle_lib.relerr(acc, a), "pot_relerr": oracle_lib.relerr(pot, pp),
                 "jerk_scaled_relerr": float(np.max(np.linalg.norm(jrk - j, axis=1) / np.maximum(np.linalg.norm(j, axis=1), scale))),
                 "jerk_strict_relerr_max": float(dj.max()), "jerk_strict_relerr_p99": float(np.quantile(dj, 0.99)),
                 "jerk_strict_relerr_median": float(np.median(dj))}
            r["ok"] = (not outside) and max(r["acc_relerr"], r["pot_relerr"], r["jerk_scaled_relerr"]) <= 1e-6 and r["jerk_strict_relerr_max"] <= 1e-5
            out[tag] = r
            return r["ok"]
        ok = judge(rep, "regf_vs_oracle")
        if gathered is not None:
            ok &= out["replicated_bitwise_equal_across_ranks"]
            isl = [np.concatenate([g[q] for g in gathered]) for q in range(4)]
            ok &= judge(isl, "islice_vs_oracle")
            out["islice_lists_equal_replicated"] = not oracle_lib.list_rows_equal(isl[3], rep[3])
            ok &= out["islice_lists_equal_replicated"]
        out["oracle"] = "oracle/regf_oracle.c: fp64 statement of regint.f:39-79 with the reference FP32 membership"
        out["jerk_note"] = ("jerk bar: 1e-6 of max(|J_i|, R_i), R_i the quadrature sum of the pair terms; the strict |dJ|/|J| is reported, "
                            "bounded at 1e-5 (cancellation outliers; the reference's FP32 path leaves 5e-6..2e-5)")
    if world > 1:
        import torch
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    out["ok"] = ok
    return out


def ref_cuda_probe(args):
    """(subprocess) the reference's own CUDA library (gpunb.velocity.cu built unmodified for sm_100, oracle/_ref) on the
    same workload through the same caller: 32 gpunb_regf_ calls of 1024 i with host arrays."""
    from nbody6ppgpu_b200.gpunb import ForceLib
    so = ROOT / "oracle" / "_ref" / "libgpunb_ref_gpu.so"
    if not so.exists():
        emit({"unavailable": "oracle/_ref/libgpunb_ref_gpu.so not built"})
        return
    ref = ForceLib(so)
    ref.devinit(0)
    n = args.n
    m, x, v, h2, dtr, rs0 = make_snapshot(n, args.m_flag)
    ref.open(n + 10, 0)
    t0 = time.perf_counter(); ref.send(m, x, v); ts = time.perf_counter() - t0
    call = ref.block_caller(h2, dtr, x, v, BLOCK, LMAX, NNBMAX, args.m_flag)
    nb = max(1, min(32, n // BLOCK - 1))
    for b in range(min(4, nb)):
        call(b * BLOCK, BLOCK)
    t0 = time.perf_counter()
    for b in range(nb):
        call(b * BLOCK, BLOCK)
    t = time.perf_counter() - t0
    ref.close()
    emit({"gint_per_s": float(BLOCK) * nb * n / t * 1e-9, "us_per_call": t / nb * 1e6, "send_ms": ts * 1e3, "calls": nb, "n": n,
          "library": "reference gpunb.velocity.cu, sm_100, same gpunb_regf_ calls with host arrays"})


def run_ref_cuda_probe(n, m_flag, local):
    try:
        env = dict(os.environ, GPU_LIST=str(local), OMP_NUM_THREADS="8")
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--ref-cuda-probe", "--n", str(n), "--m-flag", str(m_flag)],
                           capture_output=True, text=True, timeout=300, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"unavailable": (r.stderr or "no output")[-200:]}
    except Exception as e:                                   # never fatal for the bench line
        return {"unavailable": repr(e)[:200]}


TU_WORKLOAD = ("Plummer N={n} Kroupa IMF, NNBOPT=100, LMAX=400, mass-weighted neighbour criterion (m_flag=1, KZ(39)=2), SMAX=0.125, "
               "ETAI=ETAR=0.02 as samples/N16k.input; Ahmad-Cohen block-step driver nbody6ppgpu_b200/csrc/ac_driver.cpp (no KS, no stellar evolution)")


def time_unit_probe(args):
    """(subprocess) integrate --tu-t N-body time units behind ONE regular-force library with the native driver
    (csrc/ac_driver.cpp; its Python twin hermite_ac.py gives the same integration bit for bit).  The irregular force always goes
    through this repo's libirr_b200.so (the same for every arm, so the arms differ in the regular-force library only);
    `b200` also uses the device-resident predictor and gpunb_b200_regcor_, `b200_host` keeps those on the host like the
    reference arms do."""
    from nbody6ppgpu_b200 import ac_native, irr, lib_path, snapshots as S
    kind = args.time_unit_probe
    so = {"ref_cuda": ROOT / "oracle" / "_ref" / "libgpunb_ref_gpu.so", "ref_avx": ROOT / "oracle" / "_ref" / "libgpunb_ref_avx.so"}.get(kind, lib_path())
    if not so.exists():
        emit({"unavailable": f"{so.name} not built"})
        return
    n, T = args.tu_n, args.tu_t
    m, x, v = S.plummer(n, 5, "kroupa")
    dev = kind == "b200"
    st, _, _ = ac_native.run(so, irr.lib_path(), m, x, v, T, nnbopt=args.tu_nnbopt, lmax=args.tu_lmax, m_flag=args.tu_mflag,
                             dtmax=args.tu_dtmax, use_predictor=2 if dev else 0, use_regcor=2 if dev else 0)
    libs = st["wall_send"] + st["wall_regf"] + st["wall_regcor"] + st["wall_irr"]
    emit({"library": {"b200": "libgpunb_b200.so + device-resident predictor on the irr library's particle table + gpunb_b200_regcor_ on the device-resident list store", "b200_host": "libgpunb_b200.so (reference ABI only)",
                      "ref_cuda": "reference gpunb.velocity.cu + gpupot.gpu.cu (sm_100, oracle/_ref)",
                      "ref_avx": f"reference reg.avx.cpp + pot.avx.cpp (oracle/_ref, {os.environ.get('OMP_NUM_THREADS')} threads)"}[kind],
          "driver": "native (nbody6ppgpu_b200/csrc/ac_driver.cpp); irregular force through libirr_b200.so for every arm",
          "wall_s_per_time_unit": st["wall_total"] / T, "t_integrated": T, "dE_over_E": st["dE_over_E"], "E0": st["e0"],
          "block_steps": st["block_steps"], "irr_steps": st["irr_steps"], "reg_steps": st["reg_steps"], "reg_blocks": st["reg_blocks"],
          "regf_calls": st["regf_calls"], "overflow_retries": st["overflow_retries"], "mean_nnb": st["mean_nnb"], "init_s": st["wall_init"],
          "wall_breakdown_s": {"gpunb_send_or_predict_send": st["wall_send"], "gpunb_regf": st["wall_regf"], "regcor": st["wall_regcor"],
                               "irr_firr_vec": st["wall_irr"], "driver_host": st["wall_total"] - libs}})


def run_time_unit_probe(kind, n, T, local, timeout=1500, extra=()):
    try:
        env = dict(os.environ, GPU_LIST=str(local), OMP_NUM_THREADS=str(cpu_threads()))
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--time-unit-probe", kind, "--tu-n", str(n), "--tu-t", str(T), *extra],
                           capture_output=True, text=True, timeout=timeout, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {"unavailable": (r.stderr or "no output")[-300:]}
    except Exception as e:                                   # never fatal for the bench line
        return {"unavailable": repr(e)[:200]}


def main():
    args = parse()
    quiet_stdout()
    # before torch/numpy pull in libgomp: the reference AVX library asserts threads <= 32 (reg.avx.cpp:7,103)
    os.environ["OMP_NUM_THREADS"] = str(cpu_threads())
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.ref_cuda_probe:
        ref_cuda_probe(args)
        return
    if args.time_unit_probe:
        time_unit_probe(args)
        return
    if args.time_unit:
        if rank == 0:
            extra = ("--tu-dtmax", str(args.tu_dtmax), "--tu-nnbopt", str(args.tu_nnbopt), "--tu-lmax", str(args.tu_lmax), "--tu-mflag", str(args.tu_mflag))
            arms = {k: run_time_unit_probe(k, args.tu_n, args.tu_t, local, extra=extra) for k in args.tu_arms.split(",")}
            workload = TU_WORKLOAD.format(n=args.tu_n)
            if (args.tu_nnbopt, args.tu_lmax, args.tu_mflag, args.tu_dtmax) != (100, 400, 1, 0.125):
                workload = (f"Plummer N={args.tu_n} Kroupa IMF, NNBOPT={args.tu_nnbopt}, LMAX={args.tu_lmax}, m_flag={args.tu_mflag}, SMAX={args.tu_dtmax}, "
                            "ETAI=ETAR=0.02; Ahmad-Cohen block-step driver nbody6ppgpu_b200/csrc/ac_driver.cpp (no KS, no stellar evolution)")
            first = arms[args.tu_arms.split(",")[0]]
            emit({"metric": "wall s per N-body time unit", "unit": "s", "higher_is_better": False, "n_gpus": 1,
                  "config": {"workload": workload}, "value": first.get("wall_s_per_time_unit"), "arms": arms})
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    os.environ["GPU_LIST"] = str(local)              # same mechanism as the reference (gpunb.velocity.cu:582-591)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nbody6ppgpu_b200 import load
    lib = load()
    lib.devinit(rank)
    if world > 1:                                    # j sharded over ranks
        from nbody6ppgpu_b200.sharding import nccl_bootstrap
        nccl_bootstrap(lib, rank, world)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(val):
        if dist is None:
            return val
        t = torch.tensor([val], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(val):
        if dist is None:
            return val
        t = torch.tensor([val], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = args.n
    ni_total = args.ni_total if args.ni_total > 0 else n
    m, x, v, h2, dtr, rs0 = make_snapshot(n, args.m_flag)
    lib.open(n + 10, rank)
    lib.set_tuning(NSLOT, NSUB)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    interactions_scale = 1.0 / world                 # roofline per GPU: each rank's kernel sums nj/world j

    def flush_l2():
        flush.fill_(rank + 1)
        torch.cuda.synchronize()

    lib.send(m, x, v)
    # ---------------- parity before anything is timed ----------------
    parity = parity_check(lib, dist, rank, world, m, x, v, h2, dtr, args.m_flag) if n > 4096 + BLOCK else {"ok": True, "skipped": "n too small"}
    if not parity["ok"]:
        if rank == 0:
            emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "parity_check": parity, "error": "parity check failed"})
        raise SystemExit(3)

    # ---------------- device-resident leg: `value` ----------------
    sblock = sweep_block(lib, world)
    lib.set_radii(h2, dtr)
    for _ in range(args.warmup):
        lib.sweep_resident(0, ni_total, sblock, LMAX, NNBMAX, args.m_flag)
    lib.reset_counters()
    sampler = ClockSampler(local)
    ms_steps = []
    for _ in range(args.steps):
        flush_l2()
        barrier()
        ms_steps.append(max_over_ranks(lib.sweep_resident(0, ni_total, sblock, LMAX, NNBMAX, args.m_flag)))
        barrier()
    clocks = sampler.stop()
    c_res = lib.counters()
    inter_step = float(ni_total) * n
    ms_per_step = sum(ms_steps) / len(ms_steps)
    value = inter_step / (ms_per_step * 1e-3) * 1e-9
    launches_res = c_res["launches"]
    sweep_blocks = (ni_total + sblock - 1) // sblock

    # per-launch duration of the dominant kernel (regf_kernel): CUDA events around each launch on its stream,
    # taken from a timed pass of ABI calls (the resident sweep does not break the stream to read events)
    lib.reset_counters()
    lib.set_tuning(0, 1)                             # ONE pair-kernel launch per call for this probe
    nprobe = min(32, (ni_total + BLOCK - 1) // BLOCK)
    for b in range(nprobe):
        i0 = b * BLOCK
        lib.regf(h2[i0:i0 + BLOCK], dtr[i0:i0 + BLOCK], x[i0:i0 + BLOCK], v[i0:i0 + BLOCK], LMAX, NNBMAX, args.m_flag)
    c_probe = lib.counters()
    kern_ms = c_probe["grav_ms"] / c_probe["grav_launches"]
    merge_ms = c_probe["merge_ms"] / c_probe["grav_launches"]
    lib.set_tuning(NSLOT, NSUB)
    int_per_launch = float(BLOCK) * n * interactions_scale

    # ---------------- end-to-end leg through the C-ABI with host buffers: `e2e` ----------------
    # caller-owned arrays and by-reference scalars set up once (the Fortran caller's static arrays)
    regf_call = lib.block_caller(h2, dtr, x, v, BLOCK, LMAX, NNBMAX, args.m_flag)
    nblk = (ni_total + BLOCK - 1) // BLOCK

    def abi_step():
        lib.send(m, x, v)
        nnb_sum = 0
        if world == 1:
            for i0 in range(0, ni_total, BLOCK):
                acc, jrk, pot, lst = regf_call(i0, min(BLOCK, ni_total - i0))
                nnb_sum += int(lst[:, 0].sum())      # the step's result is read on the host
        else:
            # i-slice mode: block b belongs to rank b mod R (NBODY6++'s MPI build slices the i-block the same way,
            # intgrt.F:982-1231); every collective call takes one block of 1024 from every rank
            for c in range((nblk + world - 1) // world):
                b = c * world + rank
                if b < nblk:
                    acc, jrk, pot, lst = regf_call(b * BLOCK, min(BLOCK, ni_total - b * BLOCK))
                    nnb_sum += int(lst[:, 0].sum())
                else:
                    regf_call(0, 0)                  # nothing left for this rank: it still joins the collective call
        return nnb_sum

    def abi_leg(pin):
        """Returns (Gint/s, seconds per step, counters, host kind, sum of neighbour counts over all ranks)."""
        pinned_arrays = [m, x, v, *regf_call.outputs] if pin else []
        kind = "pinned" if (pinned_arrays and lib.pin_host(*pinned_arrays)) else "pageable"
        if world > 1:
            lib.set_islice(1)
        e2e_warm = max(1, min(args.warmup, 1)) if ni_total >= 500_000 else args.warmup
        for _ in range(e2e_warm):
            abi_step()
        lib.reset_counters()
        torch.cuda.synchronize()
        t_tot, nnb = 0.0, 0
        for _ in range(args.steps):
            flush_l2()
            barrier()
            t0 = time.perf_counter()
            nnb = abi_step()
            torch.cuda.synchronize()
            t_tot += max_over_ranks(time.perf_counter() - t0)
            barrier()
        c = lib.counters()
        if world > 1:
            lib.set_islice(0)
        if pinned_arrays:
            lib.unpin_host(*pinned_arrays)
        return inter_step * args.steps / t_tot * 1e-9, t_tot / args.steps, c, kind, sum_over_ranks(float(nnb))

    e2e_val, e2e_s, c_e2e, host_kind, nnb_sum = abi_leg(not args.pageable)
    e2e_other = None
    if not args.quick:
        ov, os_, oc, ok_, _ = abi_leg(args.pageable)
        e2e_other = {"value": ov, "unit": UNIT, "ms_per_step": os_ * 1e3, "host_arrays": ok_}
    h2d_step = sum_over_ranks(c_e2e["h2d_bytes"]) / args.steps
    d2h_step = sum_over_ranks(c_e2e["d2h_bytes"]) / args.steps
    lib.profile(rank)

    # ---------------- FP32 pipe microbenchmark (roofline denominator measured in the same run) ----------------
    # best of the FMA shapes that do not starve on register-file bandwidth (packed FFMA2 with a shared / repeated
    # operand; a scalar FFMA with three distinct registers only reaches ~70 % of the lane rate on this part)
    ffma_tflops = max(lib.fp32_microbench(mode, 8192) for mode in (1, 8, 10) for _ in range(2))
    lib.close()

    # ---------------- the smaller BASELINE configs, quick (one GPU only) ----------------
    def quick_config(n2, m_flag2):
        m2, x2, v2, h22, dtr2, _ = make_snapshot(n2, m_flag2)
        lib.open(n2 + 10, rank)
        lib.send(m2, x2, v2)
        lib.set_radii(h22, dtr2)
        # small j-sets: fewer, larger launch groups (the block of a resident sweep is free; 4736 = 32 x 148 i-tiles)
        sb2 = BLOCK if n2 > 65536 else 4736
        for _ in range(2):
            lib.sweep_resident(0, n2, sb2, LMAX, NNBMAX, m_flag2)
        ms = min(lib.sweep_resident(0, n2, sb2, LMAX, NNBMAX, m_flag2) for _ in range(3))
        call2 = lib.block_caller(h22, dtr2, x2, v2, BLOCK, LMAX, NNBMAX, m_flag2)
        pinned2 = [m2, x2, v2, *call2.outputs]
        kind2 = "pinned" if lib.pin_host(*pinned2) else "pageable"

        def step2():
            lib.send(m2, x2, v2)
            for i0 in range(0, n2, BLOCK):
                call2(i0, min(BLOCK, n2 - i0))
        step2()
        reps = 3 if n2 > 100_000 else 10
        t0 = time.perf_counter()
        for _ in range(reps):
            step2()
        te = (time.perf_counter() - t0) / reps
        lib.unpin_host(*pinned2)
        lib.close()
        val2 = float(n2) * n2 / (ms * 1e-3) * 1e-9
        return {"n": n2, "m_flag": m_flag2, "sweep_block": sb2, "value": val2, "e2e": float(n2) * n2 / te * 1e-9, "unit": UNIT, "ms_per_sweep": ms,
                "frac_of_sweep": val2 / (2.0 * 128 * 148 * 1965e6 / FLOP_PER_INT * 1e-9), "host_arrays": kind2}

    configs = None
    ref_cuda = None
    time_unit = None
    if world == 1 and not args.quick and n >= 500_000:
        configs = {"N256k_mflag0": quick_config(262144, 0), "N16k_mflag1": quick_config(16000, 1)}
        ref_cuda = run_ref_cuda_probe(n, args.m_flag, local)
        # second half of BASELINE.json's metric, on a bounded sample: 1/8 time unit at N=16k behind this library
        # (`python bench.py --time-unit` integrates a whole unit behind this and both reference libraries)
        time_unit = run_time_unit_probe("b200", 16000, 0.125, local, timeout=300)
        time_unit["workload"] = TU_WORKLOAD.format(n=16000)

    if world > 1:
        barrier()
        lib.nccl_finalize()
        dist.destroy_process_group()
        if rank != 0:
            return

    peaks, peak_src = measured_peaks()
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    fp32_nominal = 2.0 * 128 * 148 * sm_max * 1e6 * 1e-12
    roof_gint = fp32_nominal * 1e12 / FLOP_PER_INT * 1e-9
    achieved_tflops = FLOP_PER_INT * int_per_launch / (kern_ms * 1e-3) * 1e-12
    mean_nnb = nnb_sum / float(ni_total)
    # j tiles once (3408 B per 64 j) + i-block in + partial sums/lists out
    alg_bytes = n * interactions_scale * (3408.0 / 64.0) + BLOCK * (64.0 + 56.0 + 4.0 * (1 + mean_nnb))
    traffic, traffic_src = None, None                # dram bytes of one regf_kernel launch from the committed ncu capture
    try:
        prof = json.loads((ROOT / "profiles" / "regf_kernel_ncu_latest.json").read_text())
        if world == 1 and prof.get("nj") == n:
            traffic = float(prof["dram_bytes_read"]) + float(prof["dram_bytes_write"])
            traffic_src = f"profiles/regf_kernel_ncu_latest.json (ncu --set full capture {prof.get('tag')}, kernel {prof.get('kernel')})"
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    gint_kernel = int_per_launch / (kern_ms * 1e-3) * 1e-9
    in_sweep_ms = ms_per_step / sweep_blocks
    in_sweep_gint = float(sblock) * n * interactions_scale / (in_sweep_ms * 1e-3) * 1e-9 if sweep_blocks > 1 else value / world
    lane_peak = 128.0 * 148 * sm_max * 1e6           # FP32 lane-instructions per second
    roofline = {
        "bound": "fp32", "kernel": "regf_kernel", "achieved": achieved_tflops, "peak": fp32_nominal, "unit": "TFLOP/s",
        "frac": achieved_tflops / fp32_nominal,
        "convention": "60 flop per interaction, the reference's own (gpunb.velocity.cu:894); see fp32_pipe_slot_frac for the executed work",
        "peak_source": f"nominal 2*128 lanes*148 SM*{sm_max:.0f} MHz (MEASURED_PEAKS.json has no FP32 entry)",
        "peak_measured_ffma": ffma_tflops, "peak_measured_ffma_how": "library microbenchmark, packed FFMA2 with <= 2 distinct register operands", "frac_of_measured_ffma": achieved_tflops / ffma_tflops,
        "flop_per_interaction": FLOP_PER_INT, "interactions_per_launch": int_per_launch,
        "launch_ms": kern_ms, "launch_ms_how": "isolated launches (one gpunb_regf_ call = one pair-kernel launch), CUDA events on the launching stream",
        "gint_per_s_kernel": gint_kernel, "roofline_gint_per_s": roof_gint,
        # what the FAR body (94 % of the tile visits) executes per pair: 27 FP32 lane-operations (13.5 packed f32x2
        # instructions) + 1 MUFU.RSQ -- the fraction of the FP32 pipe's lane-operation slots the kernel fills
        "fp32_lane_ops_per_interaction": 27.0, "fp32_pipe_slot_frac": 27.0 * gint_kernel * 1e9 / lane_peak,
        "in_sweep_block_ms": in_sweep_ms, "in_sweep_block_i": sblock, "in_sweep_gint_per_s_per_gpu": in_sweep_gint,
        "in_sweep_how": "ms_per_step / blocks of the pipelined resident sweep: consecutive launches overlap their tails",
        "traffic": traffic, "traffic_source": traffic_src,
        # the same 60 flop/interaction over the WHOLE pipelined sweep (`value`), where the tail of every launch is
        # filled by the next block's CTAs -- per GPU
        "frac_of_sweep": value / world / roof_gint,
        "fp32_pipe_slot_frac_of_sweep": 27.0 * value / world * 1e9 / lane_peak,
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) * 1e-9,
                "peak_gbs": hbm_peak, "peak_source": peak_src, "frac": alg_bytes / (kern_ms * 1e-3) * 1e-9 / hbm_peak},
        "merge_kernel_ms": merge_ms,
    }

    if world == 1:
        e2e_api = (f"gpunb_send_ + {nblk} gpunb_regf_ calls of {BLOCK} i (ctypes, {host_kind} caller-owned host arrays allocated once; result rows "
                   "written by the kernels over PCIe, d2h = bytes of valid rows)")
    else:
        e2e_api = (f"every rank: gpunb_send_ + {(nblk + world - 1) // world} collective gpunb_regf_ calls in i-slice mode (rank r passes block c R + r of "
                   f"{BLOCK} i and receives its own rows; ctypes, {host_kind} caller-owned host arrays); bytes summed over ranks")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(n, args.m_flag, rs0, world),
        "sample": f"send + every i-block of the sweep ({nblk} blocks of {BLOCK}) against all {n} j per step",
        "run": {"ni_per_step": ni_total, "sweep_block": sblock, "mean_nnb": mean_nnb, "interactions_per_step": inter_step,
                "pipeline": {"sweep_slots": NSLOT if NSLOT else 3, "regf_subblocks": NSUB}},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                "ms_per_step": e2e_s * 1e3, "api": e2e_api, "host_arrays": host_kind, "mode": "single process" if world == 1 else "i-slice (collective calls)"},
        "gpu_launches": int(launches_res),
        "roofline": roofline,
        "parity_check": parity,
        "jerk_strict": {"relerr_max": parity.get("regf_vs_oracle", {}).get("jerk_strict_relerr_max"), "tolerance_used": 1e-5,
                        "north_star_tolerance": 1e-6, "scaled_relerr": parity.get("regf_vs_oracle", {}).get("jerk_scaled_relerr"),
                        "note": "strict per-particle |dJ|/|J| at N=1M on the parity block; the 1e-6 bar is met under the cancellation-aware norm only"},
    }
    if e2e_other is not None:
        out["e2e"]["pageable" if e2e_other["host_arrays"] == "pageable" else "pinned"] = e2e_other
    if configs is not None:
        out["configs"] = configs
    if ref_cuda is not None:
        out["ref_cuda"] = ref_cuda
    if time_unit is not None:
        out["time_unit"] = time_unit
        out["wall_s_per_time_unit"] = time_unit.get("wall_s_per_time_unit")

    if not args.no_cpu_baseline and rank == 0:
        threads = cpu_threads()
        os.environ["OMP_NUM_THREADS"] = str(threads)
        ref = load_reference_avx()
        if ref is not None:
            ref.open(n + 10, 0)
            time_reference(ref, m, x, v, h2, dtr, 1, args.m_flag)
            t, inter = time_reference(ref, m, x, v, h2, dtr, args.cpu_blocks, args.m_flag)
            ref.close()
            out["cpu_baseline"] = {"value": inter / t * 1e-9, "unit": UNIT, "cores": threads, "kind": "reference",
                                   "sample": f"reference reg.avx.cpp (oracle/_ref): send + {args.cpu_blocks} regf calls of {BLOCK} i against all {n} j, {t:.1f} s"}
        else:
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref not built"}
    emit(out)


if __name__ == "__main__":
    main()
