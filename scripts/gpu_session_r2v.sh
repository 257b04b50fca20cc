#!/bin/bash
# round 2, session v: shared particle table / resident lists behind the native driver, time unit (4 arms), full bench + reference
# arm, ncu launch list and full capture of the pair kernel at HEAD
TAG=r2v
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
export GPUNB_IRR_OUT=gpurun_out/irr_table_$TAG.json
timeout 900 python -m pytest tests/test_hermite_ac.py tests/test_irr_gpu.py tests/test_irr_cpu.py tests/test_regf_gpu.py -m gpu -x -q -s > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc $?"; grep -E "shared state|wall_total|passed|failed|Error" gpurun_out/pytest_$TAG.log | tail -8
timeout 900 python bench.py --time-unit --tu-t 1.0 > gpurun_out/time_unit_$TAG.json 2> gpurun_out/time_unit_$TAG.err
echo "time unit rc $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/time_unit_r2v.json"))
for k, a in d["arms"].items():
    print(k, "wall/tu %.2f dE/E %.2e" % (a.get("wall_s_per_time_unit", -1), a.get("dE_over_E", 0)), {q: round(v, 3) for q, v in a.get("wall_breakdown_s", {}).items()})
PY
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $?"; cut -c1-1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --ni-total 16384 --no-cpu-baseline --quick > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:regf_kernel -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python scripts/ncu_target.py > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
