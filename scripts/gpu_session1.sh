#!/bin/bash
# Session: tests + pipeline probe + bench (1 GPU).  Usage: scripts/gpu_session1.sh <tag>
TAG=${1:-r01m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
GPUNB_DRIFT_OUT=gpurun_out/energy_drift_$TAG.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_$TAG.json timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log; tail -8 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/pipeline_probe.py 1000000 192 > gpurun_out/probe_$TAG.log 2>&1; grep -v "^#\|^\[R" gpurun_out/probe_$TAG.log | tail -16
GPUNB_B200_STATS=1 timeout 300 python scripts/pipeline_probe.py 1000000 32 2>&1 | grep "tile visits" | tee -a gpurun_out/probe_$TAG.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
