#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 180 python tools/test_edge_nn.py > gpurun_out/nn3.log 2>&1
timeout 120 python tools/prof_edge_nn.py >> gpurun_out/nn3.log 2>&1
cat gpurun_out/nn3.log
