#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or neural or npdnp" 2>&1 | tail -15 >> gpurun_out/tmp.log
timeout 120 python tools/prof_edge_nn.py 2>&1 | tail -4 >> gpurun_out/tmp.log
cat gpurun_out/tmp.log
