#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2u_*
for v in "" s1 s2; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/r2u_sweep.log
  timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2u_sweep.log
done
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_pt.so PDP_PHASE_TIMING=2 timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 1 2>&1 | grep -v "^layout" >> gpurun_out/r2u_sweep.log
unset PDP_B200_LIB
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 3 2>&1 | grep "^E=" >> gpurun_out/r2u_sweep.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2u_sweep.log
cat gpurun_out/r2u_sweep.log
