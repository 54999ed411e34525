#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2p_*
for v in "" np; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/r2p_sweep.log
  timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2p_sweep.log
done
unset PDP_B200_LIB
timeout 600 python -m pytest tests -m gpu -q -x -k "blocked or golden or trajectory or full_size or oracle" 2>&1 | tail -5 >> gpurun_out/r2p_sweep.log
cat gpurun_out/r2p_sweep.log
