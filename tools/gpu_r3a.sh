#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r3a_*
timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r3a.log
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 3 2>&1 | grep "^E=" >> gpurun_out/r3a.log
timeout 200 python tools/prof_create.py 2>&1 | tail -2 >> gpurun_out/r3a.log
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r3a.log
cat gpurun_out/r3a.log
