#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -2 >> gpurun_out/tmp.log
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 2 2>&1 | grep "^E=" | tail -1 >> gpurun_out/tmp.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/tmp.log
timeout 120 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -1 >> gpurun_out/tmp.log
timeout 300 python tools/prof_neural.py >> gpurun_out/tmp.log 2>&1
cat gpurun_out/tmp.log
