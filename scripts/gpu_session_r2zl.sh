#!/bin/bash
# round 2, session zl: final single-GPU validation after the host-team change: GPU test-suite, smoke, bench, time unit (4 arms)
TAG=r2zl
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.txt 2>&1; echo "smoke rc $?"; grep "smoke ok" gpurun_out/smoke_$TAG.txt | cut -c1-300
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $? in $SECONDS s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2zl.json"))
r = d["roofline"]
print("value %.1f e2e %.1f (pageable %.1f) frac %.4f launch_ms %.4f frac_of_sweep %.4f ref_cuda %.1f cpu %.1f/%d wall_s_per_time_unit %s parity %s" % (
    d["value"], d["e2e"]["value"], d["e2e"].get("pageable", {}).get("value", -1), r["frac"], r["launch_ms"], r["frac_of_sweep"],
    d["ref_cuda"]["gint_per_s"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d.get("wall_s_per_time_unit"), d["parity_check"]["ok"]))
print({k: (round(v["value"], 1), round(v["e2e"], 1)) for k, v in d["configs"].items()})
PY
timeout 900 python bench.py --time-unit --tu-t 1.0 > gpurun_out/time_unit_$TAG.json 2> gpurun_out/time_unit_$TAG.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/time_unit_r2zl.json"))
for k, a in d["arms"].items():
    if "wall_s_per_time_unit" not in a: print(k, a); continue
    print(k, "wall/tu %.2f dE/E %.3e" % (a["wall_s_per_time_unit"], a["dE_over_E"]), {q: round(v, 3) for q, v in a["wall_breakdown_s"].items()}, a["block_steps"], a["irr_steps"], a["reg_steps"], a["reg_blocks"], a["regf_calls"])
PY
