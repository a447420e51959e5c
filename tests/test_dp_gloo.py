"""Data-parallel host logic on CPU: world_size 2, gloo.  Replicas start from rank 0's weights, the flat gradient
buffer is summed across ranks, 1/world is folded into the optimizer's grad_scale, seeds are sharded by rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Arena(object):
    def __init__(self, rank):
        self.data = torch.full((16,), float(rank + 1))
        self.grad = torch.arange(16, dtype=torch.float32) * (rank + 1)
        self.refreshed = 0

    def refresh_bf16(self):
        self.refreshed += 1


class _Target(object):
    def __init__(self, rank):
        self._a = _Arena(rank)

    def arena(self):
        return self._a


class _Opt(object):
    def __init__(self, rank):
        self.target = _Target(rank)
        self.grad_transform = None
        self.grad_scale = 1.0


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mocogan_chainer_b200 import parallel
    r, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    opts = [_Opt(rank), _Opt(rank)]
    assert parallel.attach(opts) == world
    for o in opts:
        a = o.target.arena()
        assert torch.equal(a.data, torch.full((16,), 1.0)) and a.refreshed == 1   # rank 0's weights everywhere
        o.grad_transform(a.grad)
        assert torch.equal(a.grad, torch.arange(16, dtype=torch.float32) * 3)      # sum over ranks 1x + 2x
        assert o.grad_scale == 0.5                                                 # mean applied inside Adam
    assert parallel.shard_seed(1234, rank) == 1234 + rank
    # replicas_identical: bit-level checksums of every model's parameter arena, all-gathered and compared
    links = [o.target for o in opts]
    assert parallel.replicas_identical(links)                  # every rank holds rank 0's weights after attach()
    if rank == 1:
        links[1].arena().data[5] = torch.nextafter(links[1].arena().data[5], torch.tensor(2.0))   # one ulp on one rank
    assert not parallel.replicas_identical(links)
    if rank == 1:
        a, b = links[0].arena().data[2].clone(), links[0].arena().data[3].clone()
        links[0].arena().data[2], links[0].arena().data[3] = b, a                                  # (equal values: no change)
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_two_rank_gloo_allreduce_and_broadcast():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_single_process_is_a_noop():
    from mocogan_chainer_b200 import parallel
    o = _Opt(0)
    assert parallel.attach([o]) == 1
    assert o.grad_transform is None and o.grad_scale == 1.0
