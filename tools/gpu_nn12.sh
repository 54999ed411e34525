#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "tensor_core or neural or npdnp or reinforce" 2>&1 | tail -5 > gpurun_out/nn12.log
timeout 300 python tools/prof_neural.py >> gpurun_out/nn12.log 2>&1
timeout 100 python __graft_entry__.py smoke >> gpurun_out/nn12.log 2>&1
cat gpurun_out/nn12.log
