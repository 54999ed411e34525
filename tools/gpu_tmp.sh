#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> gpurun_out/tmp.log
timeout 300 python tools/prof_forward.py --module-only --reps 3 2>&1 | tail -7 >> gpurun_out/tmp.log
timeout 300 python tools/prof_forward.py --reps 2 2>&1 | grep -B1 -A12 "^rep 1" | head -14 >> gpurun_out/tmp.log
cat gpurun_out/tmp.log
