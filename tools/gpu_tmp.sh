#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/tmp.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or neural or npdnp" 2>&1 | tail -15 >> gpurun_out/tmp.log
timeout 120 python tools/prof_edge_nn.py 2>&1 | tail -4 >> gpurun_out/tmp.log
PDP_B200_NN_STAGES=3 timeout 120 python tools/prof_edge_nn.py 2>&1 | grep "^E=" | tail -1 >> gpurun_out/tmp.log
PDP_B200_LIB=$PWD/pdp_solver_b200/csrc/libpdp_b200_alt_nnt.so timeout 120 python tools/prof_edge_wait.py 1200000 2>&1 | grep -v "^  chunk [0-9]*:  " >> gpurun_out/tmp.log
cat gpurun_out/tmp.log
