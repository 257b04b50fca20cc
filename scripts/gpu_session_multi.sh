#!/bin/bash
# Multi-GPU session: exchange tests, pipeline probe and bench under torchrun.  Usage: scripts/gpu_session_multi.sh <tag> <ngpu> [skiptests]
TAG=${1:-r01m}; G=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
if [ -z "$3" ]; then
  timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_multi_${TAG}_$G.log; tail -8 gpurun_out/pytest_multi_${TAG}_$G.log
fi
timeout 300 $TR --master-port 29711 scripts/pipeline_probe.py 1000000 192 > gpurun_out/probe_${TAG}_$G.log 2>&1; grep "nslot\|nsub" gpurun_out/probe_${TAG}_$G.log | grep "rank 0" | tail -12
timeout 400 $TR --master-port 29712 bench.py --gpus $G --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_$G.json 2> gpurun_out/bench_${TAG}_$G.err; cat gpurun_out/bench_${TAG}_$G.json; tail -3 gpurun_out/bench_${TAG}_$G.err
