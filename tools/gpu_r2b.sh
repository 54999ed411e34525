#!/bin/bash
# round 2, GPU call B: padded transposed variable rows; product vs accurate-log build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_slow.so timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_tests_slow.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests_slow.log
rm -f gpurun_out/r2b_sweep.log
for v in "" slow; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/r2b_sweep.log
  timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 >> gpurun_out/r2b_sweep.log 2>&1
done
export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_pt.so
echo "=== phase timing" >> gpurun_out/r2b_sweep.log
PDP_PHASE_TIMING=2 timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 2 >> gpurun_out/r2b_sweep.log 2>&1
unset PDP_B200_LIB
echo "=== config0-like: 5000 x n=100" >> gpurun_out/r2b_sweep.log
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 3 >> gpurun_out/r2b_sweep.log 2>&1
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2b_tests.log | tail -15; echo ---- slow; grep -E "passed|failed|FAILED|rc=" gpurun_out/r2b_tests_slow.log | tail -15; grep -v "^layout" gpurun_out/r2b_sweep.log
