#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/nn8.log
for v in t1 t2; do
  echo "terms $v" >> gpurun_out/nn8.log
  PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so timeout 120 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -1 >> gpurun_out/nn8.log
done
cat gpurun_out/nn8.log
