#!/bin/bash
# usage: tools/gpu_run_tests.sh <tag> [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
if [ -n "$1" ]; then
  timeout 1700 python -m pytest tests -m gpu -q -x -k "$1" 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
else
  timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
fi
cat gpurun_out/${tag}_tests.log
