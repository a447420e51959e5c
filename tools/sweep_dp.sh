#!/bin/bash
# Sweeps the data-parallel overlap knobs (mocogan_chainer_b200/parallel.py: attach) on N GPUs of one box.
# usage: tools/sweep_dp.sh N "buckets,thin_ctas,sm_reserve" ...   -> one bench.py JSON line per configuration
N=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  IFS=, read b t r <<< "$cfg"
  MCG_DP_BUCKETS=$b MCG_DP_THIN_CTAS=$t MCG_DP_SM_RESERVE=$r timeout 150 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 80 \
    --warmup 10 --no-cpu-baseline --no-gen --no-other-model --no-sustained > gpurun_out/sweep_${N}_${b}_${t}_${r}.log 2>&1
  echo "cfg=$cfg rc=$? $(grep -h '^{' gpurun_out/sweep_${N}_${b}_${t}_${r}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('steps/s %.1f ms %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" 2>&1 | tail -1)"
done
