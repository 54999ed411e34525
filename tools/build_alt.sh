#!/bin/bash
# builds an A/B variant of the product library: tools/build_alt.sh <suffix> <extra nvcc flags...>
set -e
cd "$(dirname "$0")/../pdp_solver_b200/csrc"
sfx=$1; shift
d=/tmp/alt_$sfx; mkdir -p $d
for f in pdp_graph pdp_layout pdp_ops pdp_loop pdp_walksat pdp_host pdp_edge_nn; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f.cu -o $d/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libpdp_b200_alt_$sfx.so $d/*.o -lcudart_static -lpthread -ldl -lrt
echo built libpdp_b200_alt_$sfx.so
