#!/bin/bash
# Builds mocogan_chainer_b200/libmcg.so (sm_100a only). Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")/../mocogan_chainer_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p ../../build
pids=()
for f in elementwise small simt_conv tc_conv; do
  ( $NVCC $FLAGS -c $f.cu -o ../../build/$f.o > ../../build/$f.log 2>&1 || { cat ../../build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o ../libmcg.so ../../build/elementwise.o ../../build/small.o ../../build/simt_conv.o ../../build/tc_conv.o -lcudart
echo built $(realpath ../libmcg.so)
