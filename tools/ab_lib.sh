#!/bin/bash
# A/B of builds of libmcg.so (same ABI, see tools/build_lib.sh MCG_LIB_OUT / MCG_EXTRA_FLAGS) on one GPU.
# usage: tools/ab_lib.sh libA.so libB.so ...   -> one summary per library: step rate + the HBM-bound kernels timed alone
mkdir -p gpurun_out
B="--steps 60 --warmup 10 --no-cpu-baseline"
for lib in "$@"; do
  MCG_LIB=$PWD/mocogan_chainer_b200/$lib timeout 200 python bench.py $B > gpurun_out/bench_$lib.log 2>&1
  echo "$lib rc=$? $(grep -h '^{' gpurun_out/bench_$lib.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('steps/s %.1f ms %.3f e2e %.1f gen %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gen']['value']))
for k in r['hbm_kernels']['kernels']: print('    %-46s %-10s %.4f ms %6.0f GB/s %.2f' % (k['kernel'], k['layer'], k['ms'], k['gb_per_s'], k['frac_of_hbm_peak']))
" 2>&1)"
done
