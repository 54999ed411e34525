#!/bin/bash
# Runs on the GPU box (under gpurun): the evidence bench.py's numbers are read against.
#   tools/make_profiles.sh <tag>     -> gpurun_out/<tag>_*
set -x
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 700 gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
for w in config0 config1 config2 config4; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${tag}_bench_$w.json 2>> gpurun_out/${tag}_bench.err
  tail -c 400 gpurun_out/${tag}_bench_$w.json
done
# every launch of two steps with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --problems 2 --no-cpu-baseline --no-config0 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
# the top kernel, full sections, same workload shape as the bench (8 x n=1M), fewer iterations
ncu --set full --clock-control none --import-source on -k regex:k_sp_run -c 1 -f -o gpurun_out/${tag}_sp_run \
    python tools/prof_sweep.py --problems 8 --iterations 12 > gpurun_out/${tag}_sp_run.log 2>&1
tail -2 gpurun_out/${tag}_sp_run.log
# the tensor-core GRU kernel of the neural model types
ncu --set full --clock-control none --import-source on -k regex:k_edge_nn -c 1 -f -o gpurun_out/${tag}_edge_gru \
    python tools/prof_edge_nn.py 300000 > gpurun_out/${tag}_edge_gru.log 2>&1
python tools/prof_edge_nn.py > gpurun_out/${tag}_edge_nn_timing.log 2>&1
tail -5 gpurun_out/${tag}_edge_nn_timing.log
