#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2w_*
timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2w_sweep.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_pt.so PDP_PHASE_TIMING=2 timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 1 2>&1 | grep -v "^layout" >> gpurun_out/r2w_sweep.log
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 3 2>&1 | grep "^E=" >> gpurun_out/r2w_sweep.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2w_sweep.log
cat gpurun_out/r2w_sweep.log
