#!/bin/bash
# ncu --set full of one regf_kernel launch at N=1M (ni=1024), source view included. Usage: scripts/ncu_regf.sh <tag> [variant]
TAG=${1:-x}; export GPUNB_B200_VARIANT=${2:-it1}
ncu --set full --clock-control none --import-source on -k regex:regf_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --ni-total 4096 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
