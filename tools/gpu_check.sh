#!/bin/bash
# usage (under gpurun): tools/gpu_check.sh <tag>  -- full GPU test suite, SP sweep timing, tensor-core layer check
cd "$(dirname "$0")/.."
tag=${1:-chk}
mkdir -p gpurun_out; rm -f gpurun_out/${tag}.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 >> gpurun_out/${tag}.log
timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/${tag}.log
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 2 2>&1 | grep "^E=" | tail -1 >> gpurun_out/${tag}.log
timeout 120 python tools/prof_edge_nn.py 2>&1 | tail -4 >> gpurun_out/${tag}.log
timeout 300 python tools/prof_neural.py >> gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
