#!/bin/bash
# round 2, session t: grid oversubscription against the single-wave tail, regcor_last, irr latencies after the sync removal
mkdir -p gpurun_out
timeout 900 python scripts/variant_probe2.py gpurun_out/variant_probe_r2t.json it1b4tq it1b4tq+o2 it1b4tq+o4 it1b4tq+o2+x it1b4tq+o4+x > gpurun_out/variant_probe_r2t.txt 2>&1
cat gpurun_out/variant_probe_r2t.txt
export GPUNB_REGCOR_OUT=gpurun_out/regcor_r2t.json GPUNB_IRR_OUT=gpurun_out/irr_table_r2t.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_r2t.json
timeout 900 python -m pytest tests/test_regcor_gpu.py tests/test_irr_gpu.py tests/test_reference_cuda_gpu.py tests/test_hermite_ac.py -m gpu -x -q -s > gpurun_out/pytest_r2t.log 2>&1
echo "pytest rc $?"; grep -E "regcor_last|regcor, 1024|passed|failed|Error" gpurun_out/pytest_r2t.log | tail -12
