#!/bin/bash
# A/B of programmatic dependent launch on one GPU: parity first (MCG_PDL=1 through the whole-step tests), then bench lines.
# libmcg_ew.so = the variant whose streaming kernels also release their successor early; build it first with
#   MCG_LIB_OUT=$PWD/mocogan_chainer_b200/libmcg_ew.so MCG_OBJ_TAG=obj_ew MCG_EXTRA_FLAGS=-DMCG_PDL_EW_TRIGGER=1 tools/build_lib.sh
mkdir -p gpurun_out
B="--steps 60 --warmup 10 --no-cpu-baseline"
MCG_PDL=1 timeout 300 python -m pytest tests/test_step_gpu.py tests/test_surface_gpu.py -m gpu -x -q > gpurun_out/tests_pdl1.log 2>&1; echo "pdl1 tests rc=$? $(tail -1 gpurun_out/tests_pdl1.log)"
for v in "0 libmcg.so" "1 libmcg.so" "1 libmcg_ew.so" "0 libmcg.so" "1 libmcg.so"; do
  set -- $v
  MCG_PDL=$1 MCG_LIB=$PWD/mocogan_chainer_b200/$2 timeout 200 python bench.py $B > gpurun_out/bench_pdl$1_$2.log 2>&1
  echo "MCG_PDL=$1 $2 rc=$? $(grep -h '^{' gpurun_out/bench_pdl$1_$2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('steps/s %.1f ms %.3f e2e %.1f gen %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gen']['value']))" 2>&1 | tail -1)"
done
