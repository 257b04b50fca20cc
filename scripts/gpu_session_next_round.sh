#!/bin/bash
# First GPU session of the next round: validate what round 1 could only draft.
#   1. the irregular-force library (tests gated behind IRR_B200_VALIDATE=1) + its probe
#   2. the golden-fixture GPU test (added after the GPU budget of round 1 was spent)
#   3. the usual full session
TAG=${1:-r02b}
mkdir -p gpurun_out
IRR_B200_VALIDATE=1 timeout 300 python -m pytest tests/test_irr_cpu.py tests/test_zz_golden_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_new_$TAG.log
timeout 300 python scripts/irr_probe.py 100000 64 2>&1 | grep -v "^#" | tee gpurun_out/irr_probe_$TAG.log
bash scripts/gpu_session2.sh $TAG
