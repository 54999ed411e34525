#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --problems 2 --n 200000 --steps 2 --warmup 1 --strong-problems 4 --no-cpu-baseline --no-config0 > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err
tail -c 1500 gpurun_out/tmp_bench.json; tail -5 gpurun_out/tmp_bench.err
