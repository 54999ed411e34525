#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
for st in 2 4; do echo "stages $st" >> gpurun_out/tmp.log; PDP_B200_NN_STAGES=$st timeout 120 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -1 >> gpurun_out/tmp.log; done
for v in rd8 rd4; do echo "$v" >> gpurun_out/tmp.log; PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so timeout 120 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -1 >> gpurun_out/tmp.log; done
cat gpurun_out/tmp.log
