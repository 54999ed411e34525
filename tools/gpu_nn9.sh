#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/nn9.log
timeout 180 python tools/test_edge_nn.py 2>&1 | tail -12 >> gpurun_out/nn9.log
timeout 120 python tools/prof_edge_nn.py 2>&1 | tail -4 >> gpurun_out/nn9.log
echo "chunk 16:" >> gpurun_out/nn9.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_c16.so timeout 120 python tools/prof_edge_nn.py 2>&1 | tail -4 >> gpurun_out/nn9.log
cat gpurun_out/nn9.log
