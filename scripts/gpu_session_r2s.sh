#!/bin/bash
# round 2, session s: native driver on the device paths, time unit at N=16k behind all four arms (T = 1), tail probe per SM,
# GPU test-suite with the new default variant
mkdir -p gpurun_out
timeout 300 python scripts/tail_probe.py > gpurun_out/tail_probe_r2s.txt 2>&1; tail -12 gpurun_out/tail_probe_r2s.txt
timeout 600 python -m pytest tests/test_hermite_ac.py -m gpu -x -q -s -k "native or device_paths" > gpurun_out/pytest_r2s_driver.log 2>&1
echo "driver pytest rc $?"; grep -E "wall_total|passed|failed|Error" gpurun_out/pytest_r2s_driver.log | tail -6
timeout 1500 python bench.py --time-unit --tu-t 1.0 > gpurun_out/time_unit_r2s.json 2> gpurun_out/time_unit_r2s.err
echo "time unit rc $?"; cat gpurun_out/time_unit_r2s.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2s.log 2>&1
echo "pytest rc $?"; tail -5 gpurun_out/pytest_r2s.log
