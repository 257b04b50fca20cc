#!/bin/bash
# round 2, session zd: register-file microbenchmark of packed FFMA2 operand patterns (scripts/exp/rfbank_exp.cu, built into build/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_r2zd.txt
timeout 240 ./build/rfbank_exp > gpurun_out/rfbank_r2zd.txt 2>&1; echo "rfbank rc $?"
tail -5 gpurun_out/rfbank_r2zd.txt
