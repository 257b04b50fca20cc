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
  e2e     the same sweep through the reference-facing C-ABI (gpunb_send_ + gpunb_regf_) with HOST
          (pageable numpy) buffers: H2D of the snapshot and of every i-block, D2H of forces and
          neighbour lists inside the timed region.
  roofline  FP32-FMA bound (no tensor cores: pairwise sum, not a contraction).  achieved = 60 flop x
          interactions of one launch / mean launch duration of regf_kernel (CUDA events around each
          launch, on its stream); peak = 2*128*148*sm_max_mhz nominal AND the FFMA rate measured by the
          library's microbenchmark in the same run.  HBM GB/s is reported only to show the kernel is far
          from memory bound.
  cpu_baseline  the reference's own AVX library (oracle/_ref, kind "reference") on the host cores, on a
          bounded sample of the same workload.

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
NSLOT = int(os.environ.get("GPUNB_B200_NSLOT", "0"))     # pipeline slots of the resident sweep (0: library default, 2 on one GPU / 3 sharded)
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
    ap.add_argument("--pageable", action="store_true", help="e2e leg with pageable caller arrays (default: the caller pins "
                    "its static arrays once with gpunb_b200_pin_host_)")
    return ap.parse_args()


def make_snapshot(n, m_flag):
    import numpy as np
    from nbody6ppgpu_b200 import snapshots as S
    m, x, v = S.plummer(n, 1, "kroupa")
    # neighbour spheres with <nnb> ~ NNBOPT = 200 everywhere (the state the RS control converges to)
    h2, dtr = S.radii_nnb(x, m, NNB_TARGET, 0.125, m_flag)
    return m, x, v, h2, dtr, float(np.sqrt(h2.min() * (m.mean() if m_flag else 1.0)))


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
        "config": {"workload": f"synthetic Plummer N={args.n} Kroupa IMF, regular-force sweep", "nj": args.n, "block": BLOCK,
                   "lmax": LMAX, "nnbmax": NNBMAX, "m_flag": args.m_flag, "rs_min": rs0, "sample": sample},
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


def main():
    args = parse()
    quiet_stdout()
    # before torch/numpy pull in libgomp: the reference AVX library asserts threads <= 32 (reg.avx.cpp:7,103)
    os.environ["OMP_NUM_THREADS"] = str(cpu_threads())
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    if world > 1:                                    # j sharded over ranks, every rank makes identical calls
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

    # ---------------- device-resident leg: `value` ----------------
    lib.send(m, x, v)
    lib.set_radii(h2, dtr)
    for _ in range(args.warmup):
        lib.sweep_resident(0, ni_total, BLOCK, LMAX, NNBMAX, args.m_flag)
    lib.reset_counters()
    sampler = ClockSampler(local)
    ms_steps = []
    for _ in range(args.steps):
        flush_l2()
        barrier()
        ms_steps.append(max_over_ranks(lib.sweep_resident(0, ni_total, BLOCK, LMAX, NNBMAX, args.m_flag)))
        barrier()
    clocks = sampler.stop()
    c_res = lib.counters()
    inter_step = float(ni_total) * n
    ms_per_step = sum(ms_steps) / len(ms_steps)
    value = inter_step / (ms_per_step * 1e-3) * 1e-9
    launches_res = c_res["launches"]

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
    # ... and pinned once, like COMMON blocks that live for the whole run (gpunb_b200_pin_host_): the snapshot is then
    # uploaded without a staging copy and result rows land straight in the caller's arrays.  --pageable: plain arrays.
    pinned_arrays = [] if args.pageable else [m, x, v, *regf_call.outputs]
    host_kind = "pinned" if (pinned_arrays and lib.pin_host(*pinned_arrays)) else "pageable"

    def abi_step():
        lib.send(m, x, v)
        nnb_sum = 0
        for i0 in range(0, ni_total, BLOCK):
            acc, jrk, pot, lst = regf_call(i0, min(BLOCK, ni_total - i0))
            nnb_sum += int(lst[:, 0].sum())          # the step's result is read on the host
        return nnb_sum

    e2e_warm = max(1, min(args.warmup, 1)) if ni_total >= 500_000 else args.warmup
    for _ in range(e2e_warm):
        abi_step()
    lib.reset_counters()
    torch.cuda.synchronize()
    t_e2e = 0.0
    e2e_steps = args.steps
    nnb_sum = 0
    for _ in range(e2e_steps):
        flush_l2()
        barrier()
        t0 = time.perf_counter()
        nnb_sum = abi_step()
        torch.cuda.synchronize()
        t_e2e += max_over_ranks(time.perf_counter() - t0)
        barrier()
    c_e2e = lib.counters()
    if pinned_arrays:
        lib.unpin_host(*pinned_arrays)
    e2e_val = inter_step * e2e_steps / t_e2e * 1e-9
    lib.profile(rank)

    # ---------------- FP32 pipe microbenchmark (roofline denominator measured in the same run) ----------------
    # best of the FMA shapes that do not starve on register-file bandwidth (packed FFMA2 with a shared / repeated
    # operand; a scalar FFMA with three distinct registers only reaches ~70 % of the lane rate on this part)
    ffma_tflops = max(lib.fp32_microbench(mode, 8192) for mode in (1, 8, 10) for _ in range(2))
    lib.close()
    if world > 1:
        barrier()
        lib.nccl_finalize()
        dist.destroy_process_group()
        if rank != 0:
            return

    peaks, peak_src = measured_peaks()
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    fp32_nominal = 2.0 * 128 * 148 * sm_max * 1e6 * 1e-12
    achieved_tflops = FLOP_PER_INT * int_per_launch / (kern_ms * 1e-3) * 1e-12
    mean_nnb = nnb_sum / float(ni_total)
    # j tiles once (3392 B per 64 j) + i-block in + partial sums/lists out
    alg_bytes = n * interactions_scale * (3392.0 / 64.0) + BLOCK * (64.0 + 56.0 + 4.0 * (1 + mean_nnb))
    traffic = None                                   # dram bytes of one regf_kernel launch from the committed ncu capture
    try:
        prof = json.loads((ROOT / "profiles" / "regf_kernel_ncu_latest.json").read_text())
        if world == 1 and prof.get("nj") == n:
            traffic = float(prof["dram_bytes_read"]) + float(prof["dram_bytes_write"])
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {
        "bound": "fp32", "kernel": "regf_kernel", "achieved": achieved_tflops, "peak": fp32_nominal, "unit": "TFLOP/s",
        "frac": achieved_tflops / fp32_nominal,
        "peak_source": f"nominal 2*128 lanes*148 SM*{sm_max:.0f} MHz (MEASURED_PEAKS.json has no FP32 entry)",
        "peak_measured_ffma": ffma_tflops, "peak_measured_ffma_how": "library microbenchmark, packed FFMA2 with <= 2 distinct register operands", "frac_of_measured_ffma": achieved_tflops / ffma_tflops,
        "flop_per_interaction": FLOP_PER_INT, "interactions_per_launch": int_per_launch, "launch_ms": kern_ms,
        "gint_per_s_kernel": int_per_launch / (kern_ms * 1e-3) * 1e-9,
        "roofline_gint_per_s": fp32_nominal * 1e12 / FLOP_PER_INT * 1e-9,
        "traffic": traffic,
        # the same 60 flop/interaction over the WHOLE pipelined sweep (`value`), where the tail of every launch is
        # filled by the next block's CTAs -- per GPU
        "frac_of_sweep": value / world / (fp32_nominal * 1e12 / FLOP_PER_INT * 1e-9),
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) * 1e-9,
                "peak_gbs": hbm_peak, "peak_source": peak_src, "frac": alg_bytes / (kern_ms * 1e-3) * 1e-9 / hbm_peak},
        "merge_kernel_ms": merge_ms,
    }

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic Plummer N={n} Kroupa IMF, regular-force sweep", "nj": n, "ni_per_step": ni_total,
                   "block": BLOCK, "lmax": LMAX, "nnbmax": NNBMAX, "m_flag": args.m_flag, "rs_min": rs0, "mean_nnb": mean_nnb,
                   "interactions_per_step": inter_step, "l2": "flushed between timed steps (256 MB fill)",
                   "parallelism": f"j-shard x{world}",
                   "pipeline": {"sweep_slots": NSLOT if NSLOT else (2 if world == 1 else 3), "regf_subblocks": NSUB}},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": c_e2e["h2d_bytes"] / e2e_steps,
                "d2h_bytes_per_step": c_e2e["d2h_bytes"] / e2e_steps, "ms_per_step": t_e2e / e2e_steps * 1e3,
                "api": f"gpunb_send_ + gpunb_regf_ (ctypes, {host_kind} caller-owned host arrays allocated once; result rows "
                       "written by the kernels over PCIe, d2h = bytes of valid rows)", "host_arrays": host_kind},
        "gpu_launches": int(launches_res),
        "roofline": roofline,
    }

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
