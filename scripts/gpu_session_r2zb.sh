#!/bin/bash
# round 2, session zb (8 GPUs): the scattered gpunb_send_ at 8 ranks -- parity behind it, send times, the bench line
TAG=r2zb; G=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -v -k "nccl and $G" 2>&1 | tail -8 > gpurun_out/pytest_multi_${TAG}_$G.log; grep -E "PASS|FAIL|passed|failed|Error" gpurun_out/pytest_multi_${TAG}_$G.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29700+G))"
timeout 300 $TR scripts/send_probe.py 2>&1 | grep -E "gpunb_send_|Error|assert" | tee gpurun_out/send_probe_${TAG}_$G.txt
GPUNB_B200_SPIN_TIMEOUT_S=20 timeout 600 $TR bench.py --gpus $G --steps 3 --warmup 3 --quick > gpurun_out/bench_${TAG}_$G.json 2> gpurun_out/bench_${TAG}_$G.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_$G.json"))
    print("world $G: value %.1f Gint/s (sweep block %d), e2e %.1f (%.2f ms/step), parity ok %s, in-sweep per GPU %.1f" % (d["value"], d["run"]["sweep_block"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity_check"]["ok"], d["roofline"]["in_sweep_gint_per_s_per_gpu"]))
except Exception as e:
    print("world $G: bench failed", e); print(open("gpurun_out/bench_${TAG}_$G.err").read()[-2500:])
PY
