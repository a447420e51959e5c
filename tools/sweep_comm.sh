#!/bin/bash
# bf16 vs fp32 gradient all-reduce (parallel.Bf16GradAllReduce) on N GPUs of one box: one bench.py line per setting.
# usage: tools/sweep_comm.sh N dtype...
N=$1; shift
mkdir -p gpurun_out
for d in "$@"; do
  MCG_DP_GRAD_DTYPE=$d timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 80 --warmup 10 --no-cpu-baseline --no-gen \
    --no-other-model --no-sustained > gpurun_out/comm_${N}_${d}.log 2>&1
  echo "grad_comm=$d rc=$? $(grep -h '^{' gpurun_out/comm_${N}_${d}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('steps/s %.1f ms %.3f e2e %.1f replicas_identical %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['replicas_identical']))" 2>&1 | tail -1)"
done
