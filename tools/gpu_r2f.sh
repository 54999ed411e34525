#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2f_*.log
for sv in "0 0" "10000 10000" "20000 20000" "30000 30000" "40000 40000" "30000 0" "0 30000" "50000 50000"; do
  set -- $sv
  echo "=== stagger C=$1 V=$2" >> gpurun_out/r2f_sweep.log
  PDP_B200_STAGGER_C=$1 PDP_B200_STAGGER_V=$2 timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 2 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2f_sweep.log
done
cat gpurun_out/r2f_sweep.log
