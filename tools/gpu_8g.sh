#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus $N" > gpurun_out/8g.log
lscpu | grep -i "^CPU(s)\|Model name\|Socket\|NUMA" >> gpurun_out/8g.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/8g_bench.json 2>> gpurun_out/8g.log
echo "rc=$?" >> gpurun_out/8g.log
tail -c 1500 gpurun_out/8g_bench.json
