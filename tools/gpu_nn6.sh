#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_nnt.so timeout 120 python tools/prof_edge_wait.py > gpurun_out/nn6.log 2>&1
cat gpurun_out/nn6.log
