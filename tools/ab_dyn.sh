#!/bin/bash
# A/B of the dynamic work distribution of the persistent fprop/dgrad kernels (MCG_TC_DYN) on one GPU: parity, then bench.
mkdir -p gpurun_out
B="--steps 60 --warmup 10 --no-cpu-baseline"
MCG_TC_DYN=1 timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_dyn1.log 2>&1; echo "dyn1 tests rc=$? $(tail -1 gpurun_out/tests_dyn1.log)"
for v in 0 1 0 1; do
  MCG_TC_DYN=$v timeout 200 python bench.py $B > gpurun_out/bench_dyn$v.log 2>&1
  echo "MCG_TC_DYN=$v rc=$? $(grep -h '^{' gpurun_out/bench_dyn$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('steps/s %.1f ms %.3f e2e %.1f gen %.0f | %s %s %.0f TF/s | conv isolated %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gen']['value'], r['kernel'], r['layer'], r['achieved'], r['step']['tc_conv_ms_per_step_isolated']))" 2>&1 | tail -1)"
done
