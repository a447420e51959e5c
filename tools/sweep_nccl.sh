#!/bin/bash
# NCCL tuning knobs for the three per-step gradient all-reduces on N GPUs of one box: one bench.py line per setting.
# usage: tools/sweep_nccl.sh N "VAR=val VAR2=val" ...      ("-" = defaults)
N=$1; shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  envs=""; [ "$cfg" != "-" ] && envs="$cfg"
  env $envs timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 60 --warmup 10 --no-cpu-baseline --no-gen \
    --no-other-model --no-sustained > gpurun_out/nccl_${N}_$i.log 2>&1
  echo "[$cfg] rc=$? $(grep -h '^{' gpurun_out/nccl_${N}_$i.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('steps/s %.1f ms %.3f replicas_identical %s' % (d['value'], d['ms_per_step'], d['replicas_identical']))" 2>&1 | tail -1)"
done
