#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/r2d_*.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_tests.log
timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 3 2>&1 | grep -v "^layout" >> gpurun_out/r2d_sweep.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_pt.so PDP_PHASE_TIMING=2 timeout 300 python tools/prof_sweep.py --problems 8 --iterations 20 --repeat 2 2>&1 | grep -v "^layout" >> gpurun_out/r2d_sweep.log
timeout 300 python tools/prof_sweep.py --problems 5000 --n 100 --iterations 50 --repeat 3 2>&1 | grep -v "^layout" >> gpurun_out/r2d_sweep.log
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
cat gpurun_out/r2d_tests.log gpurun_out/r2d_sweep.log; cat gpurun_out/r2d_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','phase_ms_per_step_rank0','roofline']}); print(d['e2e'])"; tail -3 gpurun_out/r2d_bench.err
