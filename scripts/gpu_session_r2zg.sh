#!/bin/bash
# round 2, session zg: single-GPU validation at HEAD: GPU test-suite, smoke, bench + reference arm
TAG=r2zg
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.txt 2>&1; echo "smoke rc $?"; grep "smoke ok" gpurun_out/smoke_$TAG.txt | cut -c1-300
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $? in $SECONDS s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2zg.json"))
r = d["roofline"]
print("value %.1f e2e %.1f (pageable %.1f) frac %.4f launch_ms %.4f frac_of_sweep %.4f ref_cuda %.1f cpu %.1f/%d wall_s_per_time_unit %s parity %s" % (
    d["value"], d["e2e"]["value"], d["e2e"].get("pageable", {}).get("value", -1), r["frac"], r["launch_ms"], r["frac_of_sweep"],
    d["ref_cuda"]["gint_per_s"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d.get("wall_s_per_time_unit"), d["parity_check"]["ok"]))
PY
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2>/dev/null; echo "reference arm in $SECONDS s"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
