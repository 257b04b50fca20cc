#!/bin/bash
# round 2, session p: regcor on the device (parity + latency), irregular-force library (parity + table against the AVX library)
mkdir -p gpurun_out
export GPUNB_REGCOR_OUT=gpurun_out/regcor_r2p.json GPUNB_IRR_OUT=gpurun_out/irr_table_r2p.json
timeout 900 python -m pytest tests/test_regcor_gpu.py tests/test_irr_gpu.py tests/test_irr_cpu.py -m gpu -x -q -s > gpurun_out/pytest_r2p.log 2>&1
echo "pytest rc $?"; tail -5 gpurun_out/pytest_r2p.log
