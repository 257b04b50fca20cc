#!/bin/bash
# round 2, session w: where the driver's regcor time goes (buckets), e2e A/B of the oversubscribed last sub-block on ONE box
mkdir -p gpurun_out
GPUNB_B200_REGCOR_PROFILE=1 timeout 300 python bench.py --time-unit-probe b200 --tu-n 16000 --tu-t 0.125 > gpurun_out/tu_prof_r2w.json 2> gpurun_out/tu_prof_r2w.err
grep "gpunb_b200_regcor:" gpurun_out/tu_prof_r2w.err; python -c "
import json; d=json.load(open('gpurun_out/tu_prof_r2w.json')); print(d['wall_s_per_time_unit'], d['wall_breakdown_s'], d['reg_blocks'], d['reg_steps'])"
for o in 1 4 1 4; do
  GPUNB_B200_REGF_OVERSUB=$o timeout 600 python bench.py --quick --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_r2w_o$o.json 2> gpurun_out/bench_r2w_o$o.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2w_o$o.json"))
print("regf oversub $o: value %.1f e2e %.1f (%.1f us/call) launch_ms %.4f frac %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3 / 977, d["roofline"]["launch_ms"], d["roofline"]["frac"]))
PY
done
