#!/bin/bash
# Builds the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU):
#   libpdp_b200.so         the product
#   libpdp_b200_strict.so  TEST build with correctly rounded fp32 log/exp (PDP_STRICT_MATH), used by the
#                          parity tests to compare whole trajectories with the C oracle bit for bit
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -Wno-deprecated-declarations --expt-relaxed-constexpr"
SRCS="pdp_graph pdp_layout pdp_ops pdp_loop pdp_walksat pdp_host pdp_edge_nn"
stale() {  # $1 = object, $2 = source
  [ ! -f "$1" ] || [ "$2" -nt "$1" ] || [ pdp_common.cuh -nt "$1" ] || [ pdp_device.cuh -nt "$1" ] || [ pdp_sweep.cuh -nt "$1" ] || [ ../../include/pdp_b200.h -nt "$1" ] || [ build.sh -nt "$1" ]
}
for f in $SRCS; do
  if stale $f.o $f.cu; then $NVCC $FLAGS ${PDP_NVCC_EXTRA} -c $f.cu -o $f.o & fi
  if stale ${f}_strict.o $f.cu; then $NVCC $FLAGS -DPDP_STRICT_MATH=1 -c $f.cu -o ${f}_strict.o & fi
done
wait
OBJS=""; SOBJS=""
for f in $SRCS; do OBJS="$OBJS $f.o"; SOBJS="$SOBJS ${f}_strict.o"; done
$NVCC $ARCH -shared -o libpdp_b200.so $OBJS -lcudart_static -lpthread -ldl -lrt
$NVCC $ARCH -shared -o libpdp_b200_strict.so $SOBJS -lcudart_static -lpthread -ldl -lrt
echo built $(pwd)/libpdp_b200.so and libpdp_b200_strict.so
