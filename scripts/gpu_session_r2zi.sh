#!/bin/bash
# round 2, session zi: the time-unit runs at the final HEAD (N = 16k four arms, N = 256k three arms)
TAG=r2zi
mkdir -p gpurun_out
timeout 900 python bench.py --time-unit --tu-t 1.0 > gpurun_out/time_unit_$TAG.json 2> gpurun_out/time_unit_$TAG.err
timeout 900 python bench.py --time-unit --tu-n 262144 --tu-t 0.03125 --tu-dtmax 0.03125 --tu-nnbopt 200 --tu-lmax 600 --tu-mflag 0 --tu-arms b200,b200_host,ref_cuda > gpurun_out/time_unit_256k_$TAG.json 2> gpurun_out/time_unit_256k_$TAG.err
python - <<'PY'
import json
for f in ("gpurun_out/time_unit_r2zi.json", "gpurun_out/time_unit_256k_r2zi.json"):
    d = json.load(open(f))
    for k, a in d["arms"].items():
        if "wall_s_per_time_unit" not in a: print(k, a); continue
        print(k, "wall/tu %.2f wall %.2f dE/E %.3e" % (a["wall_s_per_time_unit"], a["wall_s_per_time_unit"] * a["t_integrated"], a["dE_over_E"]), {q: round(v, 3) for q, v in a["wall_breakdown_s"].items()}, a["block_steps"], a["irr_steps"], a["reg_steps"], a["reg_blocks"], a["regf_calls"])
PY
