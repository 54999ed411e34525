#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r3b.log
for v in "" ni; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/r3b.log
  timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r3b.log
done
cat gpurun_out/r3b.log
