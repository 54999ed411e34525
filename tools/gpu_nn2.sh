#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/prof_edge_nn.py > gpurun_out/nn2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edge_nn -c 1 -f -o gpurun_out/nn2_gru python tools/prof_edge_nn.py 300000 > gpurun_out/nn2_ncu.log 2>&1
cat gpurun_out/nn2.log
