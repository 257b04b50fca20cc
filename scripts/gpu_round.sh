#!/bin/bash
# One GPU session: tests, bench, ncu launch list + full capture of the top kernel. Usage: scripts/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json
# launch list of the same command (short: 16 i-blocks per step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --ni-total 16384 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
# full capture of the top kernel
ncu --set full --clock-control none --import-source on -k regex:regf_kernel -s 4 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --ni-total 8192 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
