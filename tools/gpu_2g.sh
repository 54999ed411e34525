#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/2g.log 2>&1
nvidia-smi topo -m >> gpurun_out/2g.log 2>&1
lscpu | head -20 >> gpurun_out/2g.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/probe_h2d.py > gpurun_out/2g_h2d.json 2>> gpurun_out/2g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 --strong-problems 0 > gpurun_out/2g_bench.json 2>> gpurun_out/2g.log
tail -3 gpurun_out/2g.log; cat gpurun_out/2g_h2d.json; tail -c 1500 gpurun_out/2g_bench.json
