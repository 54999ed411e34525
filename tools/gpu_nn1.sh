#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 180 python tools/test_edge_nn.py > gpurun_out/nn1.log 2>&1
echo "rc=$?" >> gpurun_out/nn1.log
cat gpurun_out/nn1.log | tail -30
