#!/bin/bash
# round 2, session u: oversubscribed gpunb_regf_ launches (e2e per call), regcor buckets, gpupot without the Newton step, GPU suite
mkdir -p gpurun_out
export GPUNB_B200_REGCOR_PROFILE=1
export GPUNB_POT_OUT=gpurun_out/pot_r2u.json GPUNB_REGCOR_OUT=gpurun_out/regcor_r2u.json GPUNB_IRR_OUT=gpurun_out/irr_table_r2u.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_r2u.json
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_r2u.log 2>&1
echo "pytest rc $?"; grep -E "regcor_last|regcor, 1024|gpunb_b200_regcor:|passed|failed|Error|Pot.A\] Ni 1000000" gpurun_out/pytest_r2u.log | tail -20
unset GPUNB_B200_REGCOR_PROFILE
for o in 1 2 4 8; do
  GPUNB_B200_REGF_OVERSUB=$o timeout 600 python bench.py --quick --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_r2u_o$o.json 2> gpurun_out/bench_r2u_o$o.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2u_o$o.json"))
print("regf oversub $o: value %.1f e2e %.1f launch_ms %.4f frac %.4f merge_ms %.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["roofline"]["merge_kernel_ms"]))
PY
done
grep -E "gpupot 1M" gpurun_out/pytest_r2u.log
