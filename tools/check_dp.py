#!/usr/bin/env python
"""Data-parallel invariant on real GPUs (run under torchrun): after every step all ranks hold IDENTICAL weights — with the
bucketed all-reduce (large weight gradients reduced on a communication stream under the rest of backward) in eager and
CUDA-graph mode.  A missed or doubly-reduced slice makes the replicas diverge immediately."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mocogan_chainer_b200 import parallel  # noqa: E402

rank, world = parallel.init_from_env()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
modes = {"eager": (False,), "graph": (True,)}.get(os.environ.get("CHECK_DP_MODE", ""), (False, True))
for graph in modes:
    up, it = bench.build_updater(4, parallel.shard_seed(1234, rank), use_graph=graph)
    x, t = it.x[0][:4].cuda(), it.t[0][:4].cuda()
    for step in range(6):
        up.step_host_inputs(x, t)
        torch.cuda.synchronize()
        if rank == 0:
            print("graph=%s step %d done" % (graph, step), flush=True)
        for name in ("image_gen", "image_dis", "video_dis"):
            w = up.get_optimizer(name).target.arena().data
            ref = w.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(w, ref), "rank %d: %s differs from rank 0 after step %d (graph=%s)" % (rank, name, step, graph)
    used = {n: (up.get_optimizer(n).grad_buckets is not None and len(up.get_optimizer(n).grad_buckets.reduced)) for n in
            ("image_gen", "image_dis", "video_dis")}
    if rank == 0:
        print("graph=%s: replicas identical over 6 steps; early-reduced parameters per pass: %s" % (graph, used), flush=True)
dist.barrier()
torch.cuda.synchronize()
# (no destroy_process_group: tearing the communicator down while captured graphs that hold NCCL kernels are still alive
# hangs at exit with NCCL 2.28; the process simply ends)
sys.stdout.flush()
os._exit(0)
