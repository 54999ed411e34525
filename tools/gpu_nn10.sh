#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/nn10.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_nofull.so timeout 60 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -2 >> gpurun_out/nn10.log
cat gpurun_out/nn10.log
