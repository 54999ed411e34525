#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
for cfg in "0 0" "0 12000" "0 24000" "8000 24000" "8000 0"; do
set -- $cfg
echo "stagger c=$1 v=$2" >> gpurun_out/tmp.log
PDP_B200_STAGGER_C=$1 PDP_B200_STAGGER_V=$2 timeout 120 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/tmp.log
done
cat gpurun_out/tmp.log
