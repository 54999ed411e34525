#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2x_*
tools/probe/probe_log > gpurun_out/r2x_probe.log 2>&1
timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2x_sweep.log
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 3 2>&1 | grep "^E=" >> gpurun_out/r2x_sweep.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2x_sweep.log
cat gpurun_out/r2x_probe.log gpurun_out/r2x_sweep.log
