#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/2g.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 --strong-problems 0 > gpurun_out/2g_bench.json 2>> gpurun_out/2g.log
tail -2 gpurun_out/2g.log; tail -c 1200 gpurun_out/2g_bench.json
