#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
for v in it1b4t it1b4; do GPUNB_B200_VARIANT=$v timeout 300 python scripts/sweep_probe.py 2>&1 | grep -v "^#\|^\[R" | tee -a gpurun_out/sweep_probe_$TAG.txt; done
for n in 10000 16000; do timeout 200 python scripts/small_n_probe.py $n gpurun_out/small_n_${n}_$TAG.json 2>&1 | grep -v "^#\|^\[R" | tee -a gpurun_out/small_n_$TAG.txt; done
