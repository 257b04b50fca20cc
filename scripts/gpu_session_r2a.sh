#!/bin/bash
# Round 2, first 1-GPU session: far-body instruction shapes, irregular-force validation, kernel variants (Newton step in the
# FAR body, transposed NEAR lanes) with their strict jerk errors, the GPU test-suite and the bench at HEAD.
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
timeout 120 build/farbody_exp packed > gpurun_out/farbody_$TAG.txt 2>&1; tail -42 gpurun_out/farbody_$TAG.txt
IRR_B200_VALIDATE=1 timeout 300 python -m pytest tests/test_irr_cpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_irr_$TAG.log
timeout 300 python scripts/irr_probe.py 100000 64 2>&1 | grep -v "^#" | tee gpurun_out/irr_probe_$TAG.log
timeout 600 python scripts/variant_probe2.py gpurun_out/variant_probe2_$TAG.json 2>&1 | tee gpurun_out/variant_probe2_$TAG.log
GPUNB_DRIFT_OUT=gpurun_out/energy_drift_$TAG.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_$TAG.json timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log; tail -6 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
ls -la gpurun_out | tail -12
