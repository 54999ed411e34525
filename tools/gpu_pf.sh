#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/prof_forward.py > gpurun_out/pf2.log 2>&1
timeout 200 python tools/prof_create.py >> gpurun_out/pf2.log 2>&1
tail -60 gpurun_out/pf2.log
