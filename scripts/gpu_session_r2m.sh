#!/bin/bash
# Round 2 multi-GPU session at HEAD: all multi-GPU parity cases that fit the box, the bench under torchrun at N = G (and G/2,
# G/4 when asked), the in-process mode.  Usage: scripts/gpu_session_r2m.sh <tag> <ngpu> [also smaller world sizes: 1]
TAG=${1:-r2m}; G=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_${TAG}_$G.txt
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -v 2>&1 | tail -30 > gpurun_out/pytest_multi_${TAG}_$G.log; grep -E "PASS|FAIL|SKIP|passed|failed" gpurun_out/pytest_multi_${TAG}_$G.log
WS="$G"; if [ -n "$3" ]; then WS=""; w=$G; while [ $w -ge 2 ]; do WS="$WS $w"; w=$((w/2)); done; fi
for W in $WS; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29700+W))"
  timeout 600 $TR bench.py --gpus $W --steps 3 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$W.json"))
    print("world $W: value %.1f Gint/s (sweep block %d), e2e %.1f, parity ok %s, in-sweep per GPU %.1f" % (d["value"], d["run"]["sweep_block"], d["e2e"]["value"], d["parity_check"]["ok"], d["roofline"]["in_sweep_gint_per_s_per_gpu"]))
except Exception as e:
    print("world $W: bench failed", e); print(open("gpurun_out/bench_${TAG}_$W.err").read()[-3000:])
PY
done
timeout 300 python scripts/inproc_probe.py $G 2>&1 | grep "^inproc" | tee gpurun_out/inproc_${TAG}_$G.txt
