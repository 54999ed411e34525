#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2c_*.log
for v in vb vc vd ve; do
  echo "=== math variant $v" >> gpurun_out/r2c_tests.log
  PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so timeout 600 python -m pytest tests -m gpu -q -k "test_forward_vs_reference or test_operators_vs_reference or test_cli" 2>&1 | grep -E "passed|failed|FAILED" >> gpurun_out/r2c_tests.log
done
for v in "" vb vc vd ve wo8 wo12 vl3 cl6; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/r2c_sweep.log
  timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 2 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2c_sweep.log
done
cat gpurun_out/r2c_tests.log gpurun_out/r2c_sweep.log
