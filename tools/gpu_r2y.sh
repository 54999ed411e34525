#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2y_*
for v in "" p2 p8; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/r2y_sweep.log
  timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2y_sweep.log
done
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_pt.so PDP_PHASE_TIMING=2 timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 1 2>&1 | grep -v "^layout" >> gpurun_out/r2y_sweep.log
cat gpurun_out/r2y_sweep.log
