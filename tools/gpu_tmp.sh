#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
for v in 2 3 4; do
  echo "=== stages $v" >> gpurun_out/tmp.log
  PDP_B200_NN_STAGES=$v timeout 200 python tools/prof_neural.py >> gpurun_out/tmp.log 2>&1
done
cat gpurun_out/tmp.log
