#!/bin/bash
# Builds mocogan_chainer_b200/libmcg.so (sm_100a only). Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")/../mocogan_chainer_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${MCG_LIB_OUT:-../libmcg.so}   # MCG_LIB_OUT / MCG_EXTRA_FLAGS: A/B builds (e.g. -DMCG_PDL_EW_TRIGGER=1), loaded with MCG_LIB
OBJ=../../build/${MCG_OBJ_TAG:-obj}
FLAGS="$MCG_EXTRA_FLAGS -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p $OBJ
pids=()
for f in elementwise small simt_conv tc_conv; do
  ( $NVCC $FLAGS -c $f.cu -o $OBJ/$f.o > $OBJ/$f.log 2>&1 || { cat $OBJ/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT $OBJ/elementwise.o $OBJ/small.o $OBJ/simt_conv.o $OBJ/tc_conv.o -lcudart
echo built $(realpath $OUT)
