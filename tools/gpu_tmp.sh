#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 3 --strong-problems 0 --no-cpu-baseline --no-config0 > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/tmp_bench.json").read().strip().split("\n")[-1])
print("N", d["n_gpus"], "value %.4g"%d["value"], "ms %.1f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "e2e ms %.1f"%d["e2e"]["ms_per_step"])
PY
