#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus $N" > gpurun_out/8g.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/8g_bench.json 2>> gpurun_out/8g.log
echo "rc=$?" >> gpurun_out/8g.log
tail -3 gpurun_out/8g.log; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/8g_bench.json").read().strip().split("\n")[-1])
    print("N", d["n_gpus"], "value %.4g"%d["value"], "ms %.1f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "e2e ms %.1f"%d["e2e"]["ms_per_step"])
    print(d.get("strong_scaling_fixed_batch"))
except Exception as e: print("ERR", e)
PY
