#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of regf_kernel (read HERE with ncu -i): writes profiles/<tag>_regf_kernel_ncu_full.md
and profiles/regf_kernel_ncu_latest.json (dram bytes per launch, used by bench.py's roofline.traffic).
Usage: ncu_summary.py gpurun_out/prof_<tag>.ncu-rep <tag> <nj>"""
import csv, io, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
rep, tag, nj = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
r = data[-1]
def g(k):
    return r[hdr.index(k)] if k in hdr else None
def unit(k):
    return units[hdr.index(k)] if k in hdr else ""
def to_bytes(k):
    v, u = float(g(k)), unit(k).lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_tma_ld.sum"]
lines = [f"# ncu --set full: {g('Kernel Name')} ({tag})", "",
         f"`ncu --set full --clock-control none --import-source on -k regex:regf_kernel` on one B200, N = nj = {nj}, ni = 1024, launch #{len(data)} of the capture.", "",
         "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if g(k) is not None:
        lines.append(f"| {k} | {g(k)} | {unit(k)} |")
stalls = []
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
        try:
            stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
tot = sum(v for v, _ in stalls) or 1.0
lines += ["", "Warp-state samples (all warps):", "", "| state | share |", "|---|---|"]
for v, n in sorted(stalls, reverse=True)[:10]:
    lines.append(f"| {n} | {100 * v / tot:.1f} % |")
inst = float(g("smsp__inst_executed.sum")); pairs = 1024.0 * nj / 32
lines += ["", f"Warp instructions per (warp, j) pair step: {inst / pairs:.2f}; cycles per pair step and sub-partition: "
          f"{float(g('gpu__time_duration.sum')) * (1e3 if unit('gpu__time_duration.sum') == 'ms' else 1.0) * 1e-6 * 1.965e9 * 592 / pairs:.1f} (at 1965 MHz)."]
(ROOT / "profiles" / f"{tag}_regf_kernel_ncu_full.md").write_text("\n".join(lines) + "\n")
(ROOT / "profiles" / "regf_kernel_ncu_latest.json").write_text(json.dumps({
    "tag": tag, "nj": nj, "ni": 1024, "kernel": g("Kernel Name"), "duration": g("gpu__time_duration.sum") + " " + unit("gpu__time_duration.sum"),
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum")}, indent=1) + "\n")
print("\n".join(lines[:40]))
