#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/nn11.log
for v in "" nosw notm "" nosw notm; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/nn11.log
  timeout 60 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -1 >> gpurun_out/nn11.log
done
cat gpurun_out/nn11.log
