#!/bin/bash
# Round 2: N-GPU session -- the whole GPU suite (multi-GPU cases up to N devices run), the 1-GPU bench with the new
# fields, the bench under torchrun.  Usage: scripts/gpu_session_r2b.sh <tag> <ngpu>
TAG=${1:-r2b}; G=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
GPUNB_DRIFT_OUT=gpurun_out/energy_drift_$TAG.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_$TAG.json timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_${TAG}_$G.log; tail -12 gpurun_out/pytest_${TAG}_$G.log
if [ -z "$SKIP_1GPU" ]; then
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_1.json 2> gpurun_out/bench_${TAG}_1.err; cat gpurun_out/bench_${TAG}_1.json; tail -3 gpurun_out/bench_${TAG}_1.err
fi
timeout 600 $TR --master-port 29712 bench.py --gpus $G --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_$G.json 2> gpurun_out/bench_${TAG}_$G.err; cat gpurun_out/bench_${TAG}_$G.json; tail -5 gpurun_out/bench_${TAG}_$G.err
