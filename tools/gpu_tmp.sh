#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; tag=r2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2z_tests.log
cat gpurun_out/r2z_tests.log
for w in config1 config2; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${tag}_bench_$w.json 2>> gpurun_out/${tag}_bench.err
  tail -c 200 gpurun_out/${tag}_bench_$w.json
done
ncu --set full --clock-control none --import-source on -k regex:k_edge_nn -c 1 -f -o gpurun_out/${tag}_edge_gru \
    python tools/prof_edge_nn.py 300000 > gpurun_out/${tag}_edge_gru.log 2>&1
python tools/prof_edge_nn.py > gpurun_out/${tag}_edge_nn_timing.log 2>&1
tail -4 gpurun_out/${tag}_edge_nn_timing.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
