"""Data parallelism for the MoCoGAN step: one process per GPU, every rank draws its own batch of
clips / latents / noise, BatchNorm statistics stay local, and the three flat gradient buffers (image_dis after pass
A, video_dis after pass B, image_gen after pass C) are each averaged with ONE collective before their Adam step —
the discriminators' all-reduce and update must be complete before pass C's dgrad reads those weights
(updater.py:111-113), so the collectives are ordered on the compute stream.

The reference has no multi-device path (single `--gpu` id, train.py:26,87-91); this layer is additive.
torch.distributed is plumbing: NCCL on GPUs, gloo in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_mean_(flat, group=None):
    """In-place sum over ranks; the 1/world factor is folded into the Adam kernel's grad_scale."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_(flat, src=0, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def attach(optimizers, group=None):
    """Makes each Adam optimizer data-parallel: replicas start from rank 0's weights, gradients are averaged."""
    w = dist.get_world_size(group) if dist.is_initialized() else 1
    for opt in optimizers:
        arena = opt.target.arena()
        broadcast_(arena.data, 0, group)
        arena.refresh_bf16()
        if w > 1:
            opt.grad_transform = lambda g, _grp=group: allreduce_mean_(g, _grp)
            opt.grad_scale = 1.0 / w
    return w


def shard_seed(base_seed, rank):
    """rank r uses seed base+r for its clips, latents and noise."""
    return int(base_seed) + int(rank)
