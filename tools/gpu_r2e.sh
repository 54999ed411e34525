#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_sp_run -c 1 -f -o gpurun_out/r2e_sp_run \
    python tools/prof_sweep.py --problems 8 --iterations 12 > gpurun_out/r2e_sp_run.log 2>&1
tail -2 gpurun_out/r2e_sp_run.log
ls -la gpurun_out/r2e_sp_run.ncu-rep
