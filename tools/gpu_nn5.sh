#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 180 python tools/test_edge_nn.py > gpurun_out/nn5.log 2>&1
timeout 120 python tools/prof_edge_nn.py >> gpurun_out/nn5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edge_nn -c 1 -f -o gpurun_out/nn5_gru python tools/prof_edge_nn.py 300000 > gpurun_out/nn5_ncu.log 2>&1
tail -8 gpurun_out/nn5.log
