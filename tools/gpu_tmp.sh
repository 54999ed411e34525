#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or neural or npdnp" 2>&1 | tail -3 >> gpurun_out/tmp.log
timeout 120 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -2 >> gpurun_out/tmp.log
timeout 300 python tools/prof_neural.py >> gpurun_out/tmp.log 2>&1
cat gpurun_out/tmp.log
