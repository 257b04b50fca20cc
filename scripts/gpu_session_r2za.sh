#!/bin/bash
# round 2, session za: the time unit where the libraries matter -- N = 262144 (BASELINE config 4 sizes: NNBOPT 200, LMAX 600),
# 1/32 time unit (SMAX = 1/32 so that the run ends synchronised), this library on the device paths / reference ABI only, reference CUDA
mkdir -p gpurun_out
timeout 1700 python bench.py --time-unit --tu-n 262144 --tu-t 0.03125 --tu-dtmax 0.03125 --tu-nnbopt 200 --tu-lmax 600 --tu-mflag 0 --tu-arms b200,b200_host,ref_cuda > gpurun_out/time_unit_256k_r2za.json 2> gpurun_out/time_unit_256k_r2za.err
echo "rc $?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/time_unit_256k_r2za.json"))
for k, a in d["arms"].items():
    if "wall_s_per_time_unit" not in a: print(k, a); continue
    print(k, "wall/tu %.1f (run %.2f s) dE/E %.2e init %.2f" % (a["wall_s_per_time_unit"], a["wall_s_per_time_unit"] * a["t_integrated"], a["dE_over_E"], a["init_s"]), {q: round(v, 3) for q, v in a["wall_breakdown_s"].items()}, a["block_steps"], a["irr_steps"], a["reg_steps"], a["reg_blocks"], a["regf_calls"])
PY
tail -3 gpurun_out/time_unit_256k_r2za.err
