#!/bin/bash
# round 2, session zj: the native driver after the parallel list-copy loops -- driver tests on the device paths, time unit at N = 16k
TAG=r2zj
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hermite_ac.py tests/test_predictor_gpu.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --time-unit --tu-t 1.0 > gpurun_out/time_unit_$TAG.json 2> gpurun_out/time_unit_$TAG.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/time_unit_r2zj.json"))
for k, a in d["arms"].items():
    if "wall_s_per_time_unit" not in a: print(k, a); continue
    print(k, "wall/tu %.2f dE/E %.3e" % (a["wall_s_per_time_unit"], a["dE_over_E"]), {q: round(v, 3) for q, v in a["wall_breakdown_s"].items()}, a["block_steps"], a["irr_steps"], a["reg_steps"], a["reg_blocks"], a["regf_calls"])
PY
