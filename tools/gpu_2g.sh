#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/2g.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -k "cli or reinforce" 2>&1 | tail -6 >> gpurun_out/2g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/2g_bench.json 2>> gpurun_out/2g.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/2g_bench_reference.json 2>> gpurun_out/2g.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload config0 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/2g_bench_config0.json 2>> gpurun_out/2g.log
cat gpurun_out/2g.log | tail -12; cut -c1-400 gpurun_out/2g_bench.json; cut -c1-300 gpurun_out/2g_bench_config0.json; cut -c1-300 gpurun_out/2g_bench_reference.json
