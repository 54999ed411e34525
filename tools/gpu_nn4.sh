#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "tensor_core or neural or npdnp or cli" 2>&1 | tail -15 > gpurun_out/nn4.log
timeout 300 python tools/prof_neural.py >> gpurun_out/nn4.log 2>&1
PDP_B200_NN=torch timeout 300 python tools/prof_neural.py >> gpurun_out/nn4.log 2>&1
cat gpurun_out/nn4.log
