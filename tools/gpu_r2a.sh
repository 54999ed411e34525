#!/bin/bash
# round 2, GPU call A: correctness of the re-laid-out blocked passes + first timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_env.log 2>&1
tools/probe/probe_math > gpurun_out/r2a_probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
for v in "" dc slow; do
  if [ -n "$v" ]; then export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_$v.so; else unset PDP_B200_LIB; fi
  echo "=== variant '$v'" >> gpurun_out/r2a_sweep.log
  timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 >> gpurun_out/r2a_sweep.log 2>&1
done
export PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_pt.so
echo "=== phase timing" >> gpurun_out/r2a_sweep.log
PDP_PHASE_TIMING=2 timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 2 >> gpurun_out/r2a_sweep.log 2>&1
unset PDP_B200_LIB
echo "=== tests with the FAST_DC variant" >> gpurun_out/r2a_tests_dc.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_dc.so timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2a_tests_dc.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests_dc.log
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_sweep.log | grep -v "^layout"; cat gpurun_out/r2a_probe.log; tail -5 gpurun_out/r2a_tests_dc.log
