#!/bin/bash
# Builds libcusumtools_b200.so for sm_100a in-tree (the .so travels to the GPU box).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include ${CT_EXTRA_NVCC_FLAGS}"
OBJS=""
for f in ct_api ct_filter ct_filter_seq ct_detect ct_cusum ct_welch ct_loader; do
  [ -f $f.cu ] || continue
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ ct_common.cuh -nt $f.o ] || [ ../../include/cusumtools_b200.h -nt $f.o ]; then
    $NVCC $FLAGS -c $f.cu -o $f.o &
  fi
  OBJS="$OBJS $f.o"
done
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../libcusumtools_b200.so $OBJS -lcudart
echo "built $(cd .. && pwd)/libcusumtools_b200.so"
