#!/bin/bash
# round 2, session q: device paths of the AC driver, time-unit probe (short), tail probe of the pair kernel at HEAD
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hermite_ac.py -m gpu -x -q -s -k device_paths > gpurun_out/pytest_r2q.log 2>&1
echo "pytest rc $?"; grep -E "^(host|irr|device) |passed|failed|Error" gpurun_out/pytest_r2q.log | tail -8
for k in b200 b200_host; do
  timeout 600 python bench.py --time-unit-probe $k --tu-n 16000 --tu-t 0.125 > gpurun_out/tu_${k}_r2q.json 2> gpurun_out/tu_${k}_r2q.err
  echo "tu $k rc $?"; cat gpurun_out/tu_${k}_r2q.json
done
timeout 300 python scripts/tail_probe.py > gpurun_out/tail_probe_r2q.txt 2>&1; tail -8 gpurun_out/tail_probe_r2q.txt
