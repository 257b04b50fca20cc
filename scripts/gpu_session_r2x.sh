#!/bin/bash
# round 2, session x (8 GPUs): every multi-GPU parity case at HEAD, the bench under torchrun at 8 and 2 ranks
TAG=r2x; G=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_${TAG}_$G.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -v 2>&1 | tail -40 > gpurun_out/pytest_multi_${TAG}_$G.log; grep -E "PASS|FAIL|SKIP|passed|failed|Error" gpurun_out/pytest_multi_${TAG}_$G.log
for W in $G 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29700+W))"
  GPUNB_B200_SPIN_TIMEOUT_S=20 timeout 600 $TR bench.py --gpus $W --steps 3 --warmup 3 --quick > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$W.json"))
    print("world $W: value %.1f Gint/s (sweep block %d), e2e %.1f, parity ok %s, in-sweep per GPU %.1f" % (d["value"], d["run"]["sweep_block"], d["e2e"]["value"], d["parity_check"]["ok"], d["roofline"]["in_sweep_gint_per_s_per_gpu"]))
except Exception as e:
    print("world $W: bench failed", e); print(open("gpurun_out/bench_${TAG}_$W.err").read()[-2500:])
PY
done
