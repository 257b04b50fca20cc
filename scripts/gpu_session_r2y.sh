#!/bin/bash
# round 2, session y: compute-sanitizer (memcheck, racecheck) over the new kernels on small cases; smoke(); driver after the scratch fix
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_regcor_gpu.py -m gpu -x -q -k "bit_for_bit or resident" > gpurun_out/sanitizer_memcheck_regcor_r2y.log 2>&1
echo "memcheck regcor rc $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_regcor_r2y.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_regcor_gpu.py -m gpu -x -q -k "bit_for_bit" > gpurun_out/sanitizer_racecheck_regcor_r2y.log 2>&1
echo "racecheck regcor rc $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_regcor_r2y.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_irr_gpu.py tests/test_irr_cpu.py -m gpu -x -q -k "batch or fp64" > gpurun_out/sanitizer_memcheck_irr_r2y.log 2>&1
echo "memcheck irr rc $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_irr_r2y.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_regf_gpu.py -m gpu -x -q -k "oversubscribed or ragged or overflow" > gpurun_out/sanitizer_memcheck_regf_r2y.log 2>&1
echo "memcheck regf rc $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_memcheck_regf_r2y.log | tail -3
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2y.txt 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke_r2y.txt | cut -c1-400
timeout 300 python bench.py --time-unit-probe b200 --tu-n 16000 --tu-t 1.0 > gpurun_out/tu_b200_r2y.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/tu_b200_r2y.json')); print('b200 arm: wall/tu %.2f' % d['wall_s_per_time_unit'], {k: round(v,3) for k,v in d['wall_breakdown_s'].items()}, d['dE_over_E'])"
timeout 300 python bench.py --time-unit-probe b200_host --tu-n 16000 --tu-t 1.0 > gpurun_out/tu_b200_host_r2y.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/tu_b200_host_r2y.json')); print('b200_host arm: wall/tu %.2f' % d['wall_s_per_time_unit'], {k: round(v,3) for k,v in d['wall_breakdown_s'].items()}, d['dE_over_E'])"
