#!/bin/bash
# round 2, session zk: does the size of the library's host-side OpenMP teams interfere with the caller's own parallel loops?
TAG=r2zk
mkdir -p gpurun_out
for ht in 4 16; do
  GPUNB_B200_HOST_THREADS=$ht timeout 900 python bench.py --time-unit --tu-t 1.0 --tu-arms b200,b200_host > gpurun_out/time_unit_${TAG}_ht$ht.json 2> gpurun_out/time_unit_${TAG}_ht$ht.err
  python - <<PY
import json
d = json.load(open("gpurun_out/time_unit_${TAG}_ht$ht.json"))
for k, a in d["arms"].items():
    if "wall_s_per_time_unit" not in a: print(k, a); continue
    print("host_threads=$ht", k, "wall/tu %.2f" % a["wall_s_per_time_unit"], {q: round(v, 3) for q, v in a["wall_breakdown_s"].items()})
PY
done
nproc
