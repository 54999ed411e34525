#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2z_tests.log
cat gpurun_out/r2z_tests.log
timeout 2400 tools/make_profiles.sh r2 > gpurun_out/r2_make_profiles.log 2>&1
tail -40 gpurun_out/r2_make_profiles.log
