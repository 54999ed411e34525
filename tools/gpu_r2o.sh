#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2o_*
timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1 >> gpurun_out/r2o_sweep.log
ncu --set full --clock-control none --import-source on -k regex:k_sp_run -c 1 -f -o gpurun_out/r2o_sp_run \
    python tools/prof_sweep.py --problems 8 --iterations 12 > gpurun_out/r2o_sp_run.log 2>&1
cat gpurun_out/r2o_sweep.log
