#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
for n in 10000 16000; do timeout 200 python scripts/small_n_probe.py $n gpurun_out/small_n_${n}_$TAG.json 2>&1 | grep -v "^#\|^\[R" | tee -a gpurun_out/small_n_$TAG.txt; done
GPUNB_B200_NOISORT=1 timeout 200 python scripts/small_n_probe.py 10000 2>&1 | grep -v "^#\|^\[R" | grep "1024\|256" | sed 's/^/NOISORT /' | tee -a gpurun_out/small_n_$TAG.txt
GPUNB_DRIFT_OUT=gpurun_out/energy_drift_$TAG.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_$TAG.json timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > gpurun_out/pytest_$TAG.log; tail -6 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
