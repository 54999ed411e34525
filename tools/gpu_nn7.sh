#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 180 python tools/test_edge_nn.py 2>&1 | tail -4 > gpurun_out/nn7.log
timeout 120 python tools/prof_edge_nn.py 2>&1 | tail -4 >> gpurun_out/nn7.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_nnt.so timeout 120 python tools/prof_edge_wait.py >> gpurun_out/nn7.log 2>&1
cat gpurun_out/nn7.log
