#!/bin/bash
# round 2, final single-GPU session: GPU test-suite, smoke, bench + reference arm, time unit at N = 16k (4 arms) and 256k (3 arms),
# IT = 2 variants at HEAD for the record
TAG=r2zz
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
export GPUNB_POT_OUT=gpurun_out/pot_$TAG.json GPUNB_REGCOR_OUT=gpurun_out/regcor_$TAG.json GPUNB_IRR_OUT=gpurun_out/irr_table_$TAG.json GPUNB_REFCUDA_OUT=gpurun_out/ref_cuda_$TAG.json
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.txt 2>&1; echo "smoke rc $?"; grep "smoke ok" gpurun_out/smoke_$TAG.txt | cut -c1-300
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2zz.json"))
r = d["roofline"]
print("value %.1f e2e %.1f (pageable %.1f) frac %.4f launch_ms %.4f frac_of_sweep %.4f traffic %s ref_cuda %.1f cpu %.1f/%d wall_s_per_time_unit %s" % (
    d["value"], d["e2e"]["value"], d["e2e"].get("pageable", {}).get("value", -1), r["frac"], r["launch_ms"], r["frac_of_sweep"], r["traffic"],
    d["ref_cuda"]["gint_per_s"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d.get("wall_s_per_time_unit")))
print(d["configs"])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref_$TAG.json
timeout 900 python bench.py --time-unit --tu-t 1.0 > gpurun_out/time_unit_$TAG.json 2> gpurun_out/time_unit_$TAG.err
timeout 900 python bench.py --time-unit --tu-n 262144 --tu-t 0.03125 --tu-dtmax 0.03125 --tu-nnbopt 200 --tu-lmax 600 --tu-mflag 0 --tu-arms b200,b200_host,ref_cuda > gpurun_out/time_unit_256k_$TAG.json 2> gpurun_out/time_unit_256k_$TAG.err
python - <<'PY'
import json
for f in ("gpurun_out/time_unit_r2zz.json", "gpurun_out/time_unit_256k_r2zz.json"):
    d = json.load(open(f))
    for k, a in d["arms"].items():
        if "wall_s_per_time_unit" not in a: print(k, a); continue
        print(k, "wall/tu %.2f dE/E %.3e" % (a["wall_s_per_time_unit"], a["dE_over_E"]), {q: round(v, 3) for q, v in a["wall_breakdown_s"].items()}, a["block_steps"], a["irr_steps"], a["reg_steps"], a["reg_blocks"], a["regf_calls"])
PY
timeout 600 python scripts/variant_probe2.py gpurun_out/variant_probe_$TAG.json it1b4tq it2b3 it2 > gpurun_out/variant_probe_$TAG.txt 2>&1; cat gpurun_out/variant_probe_$TAG.txt | cut -c1-420
