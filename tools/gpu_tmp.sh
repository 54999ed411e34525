#!/bin/bash
# scratch script for one-off gpurun calls
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2z_tests.log
cat gpurun_out/r2z_tests.log
timeout 120 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep "^E=" | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
