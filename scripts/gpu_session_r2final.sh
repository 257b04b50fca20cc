#!/bin/bash
# round 2: GPU test-suite and smoke at the final HEAD
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2final.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_r2final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2final.txt 2>&1; echo "smoke rc $?"; grep "smoke ok" gpurun_out/smoke_r2final.txt | cut -c1-300
timeout 600 python bench.py --time-unit-probe b200 --tu-n 16000 --tu-t 1.0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b200 arm: wall/tu %.2f' % d['wall_s_per_time_unit'], {k: round(v,3) for k,v in d['wall_breakdown_s'].items()}, d['dE_over_E'])"
