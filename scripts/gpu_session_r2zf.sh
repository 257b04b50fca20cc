#!/bin/bash
# round 2, session zf (G GPUs): the in-process multi-GPU mode with one enqueue thread per device -- parity case, then A/B of
# GPUNB_B200_ENQUEUE_THREADS=0/1 on the same box.  Usage: gpu_session_r2zf.sh G
G=${1:-2}; TAG=r2zf
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -$G > gpurun_out/gpu_${TAG}_$G.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "inprocess and $G" > gpurun_out/pytest_inproc_${TAG}_$G.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_inproc_${TAG}_$G.log
for th in ${THREADS_ARMS:-0 1}; do
  GPUNB_B200_ENQUEUE_THREADS=$th timeout 300 python scripts/inproc_probe.py $G 2>&1 | grep "^inproc" | sed "s/^/threads=$th /" | tee -a gpurun_out/inproc_${TAG}_$G.txt
done
