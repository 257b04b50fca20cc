#!/bin/bash
# 8-GPU session (charged 8x): one exchange test at R=8 and the bench.  Usage: scripts/gpu_session_8.sh <tag>
TAG=${1:-r01n}; G=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "nccl and 8" 2>&1 | tail -15 > gpurun_out/pytest_multi_${TAG}_$G.log; tail -4 gpurun_out/pytest_multi_${TAG}_$G.log
timeout 300 $TR --master-port 29712 bench.py --gpus $G --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_$G.json 2> gpurun_out/bench_${TAG}_$G.err; cat gpurun_out/bench_${TAG}_$G.json; tail -3 gpurun_out/bench_${TAG}_$G.err
