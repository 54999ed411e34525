#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench.json").read().strip().split("\n")[-1])
print("N", d["n_gpus"], "value %.4g"%d["value"], "ms %.1f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "e2e ms %.1f"%d["e2e"]["ms_per_step"], "h2d %.1f"%d["e2e"]["h2d_ms_per_step_max_rank"], d["e2e"]["phase_ms_per_step_rank0"], "frac %.4f"%d["roofline"]["frac"])
PY
python tools/probe_h2d.py
