#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/neural_launches.csv python tools/prof_neural.py --iterations 4 > gpurun_out/tmp.log 2>&1
tail -3 gpurun_out/tmp.log
