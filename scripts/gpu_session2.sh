#!/bin/bash
# Full 1-GPU session: smoke, tests, probe, bench (+ reference arm, + the smaller configs), ncu launch list + full capture.
# Usage: scripts/gpu_session2.sh <tag>
TAG=${1:-r01p}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^#\|^\[R" | tail -3 | tee gpurun_out/smoke_$TAG.log
GPUNB_DRIFT_OUT=gpurun_out/energy_drift_$TAG.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_$TAG.json timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log; tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/pipeline_probe.py 1000000 192 > gpurun_out/probe_$TAG.log 2>&1; grep -v "^#\|^\[R" gpurun_out/probe_$TAG.log | tail -16
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_$TAG.json
# the smaller configurations of BASELINE.json (not bench lines: recorded beside the parity tests)
timeout 200 python bench.py --n 262144 --steps 3 --warmup 3 --cpu-blocks 16 > gpurun_out/bench_256k_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_256k_$TAG.json
timeout 200 python bench.py --n 16000 --m-flag 1 --steps 5 --warmup 3 --cpu-blocks 15 > gpurun_out/bench_16k_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_16k_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --ni-total 16384 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:regf_kernel -s 4 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --ni-total 8192 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -14
