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


# defaults of the overlap knobs (see attach); chosen from the 2- and 8-GPU sweeps recorded in profiles/
DEFAULT_BUCKETS = 0
DEFAULT_THIN_CTAS = 0


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


class Bf16GradAllReduce(object):
    """The flat fp32 gradient of one model through the collective as bf16: round (one kernel), sum over ranks (NCCL; NVLS
    accumulates bf16 in fp32 inside the switch), widen back into the fp32 buffer Adam reads (one kernel).  Halves the bytes
    on NVLink — at 8 GPUs the three all-reduces are the whole gap to linear scaling and the video discriminator's 44 MB
    sits exposed between pass B and pass C.  What it costs: every rank's gradient is rounded to bf16 once (2^-9
    relative) before the sum; the bf16 compute mode's own gradients differ from float64 by 5-25 % (tests/
    test_step_gpu.py), so this is far inside the mode's tolerance.  Replicas stay bit-identical: every rank widens the same
    reduced values.  Only used in the bf16 compute mode; MCG_DP_GRAD_DTYPE=fp32 switches it off."""

    def __init__(self, arena, group=None):
        self.group = group
        self.buf = torch.empty(arena.grad.numel(), dtype=torch.bfloat16, device=arena.grad.device)

    def __call__(self, flat):
        from . import kernels as K
        K.cast_bf16(flat, self.buf)
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        K.cast_f32(self.buf, flat)
        return flat


def grad_comm_dtype():
    """'bf16' or 'fp32': what the gradient all-reduce carries (see Bf16GradAllReduce)."""
    from . import chainer
    want = os.environ.get("MCG_DP_GRAD_DTYPE", "").strip().lower()
    if want in ("fp32", "bf16"):
        return want
    return "bf16" if chainer.config.compute_dtype == "bf16" else "fp32"


def broadcast_(flat, src=0, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


class GradBuckets(object):
    """All-reduce overlapped with backward (north_star).  The weight gradient of a discriminator's widest
    layer is the first to be complete and most of the payload (Dv.dc4.W: 33.5 of 44.2 MB), so every LARGE parameter is
    all-reduced on a communication stream as soon as the last kernel that accumulates into it has been launched, under
    the rest of the backward pass; what is left (the small parameters, as the contiguous gaps of the flat gradient
    buffer) is reduced when backward is over.  How many kernels write a parameter in one pass (2 for the
    discriminators: real and fake chain; 1 for the generator) is LEARNED in the first pass, which uses the plain flat
    all-reduce — identically on every rank, so all ranks issue the same collectives in the same order."""

    def __init__(self, opt, group=None, min_numel=1 << 20, early_group=None):
        self.group = group                                   # what is left when backward is over (exposed: full width)
        self.early_group = early_group if early_group is not None else group   # under backward (may be a thin one)
        self.arena = opt.target.arena()
        self.big = {}
        for p in self.arena.params:
            off, n = self.arena.offsets[id(p)]
            if n >= min_numel:
                self.big[id(p)] = (off, n)
                p.grad_written_hook = self.written
        self.expected = None
        self.comm = torch.cuda.Stream()
        self.begin()

    def begin(self):
        self.count = {k: 0 for k in self.big}
        self.events = {k: [] for k in self.big}
        self.reduced = []

    def written(self, p):
        k = id(p)
        if k not in self.big:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.events[k].append(ev)
        self.count[k] += 1
        if self.expected is not None and self.count[k] == self.expected.get(k, -1):
            off, n = self.big[k]
            for e in self.events[k]:
                self.comm.wait_event(e)
            with torch.cuda.stream(self.comm):
                dist.all_reduce(self.arena.grad[off:off + n], op=dist.ReduceOp.SUM, group=self.early_group)
            self.reduced.append((off, n))

    def finish(self, flat):
        """The optimizer's grad_transform: everything not reduced early, then wait for the communication stream."""
        if self.expected is None:
            self.expected = dict(self.count)       # first pass: learn the writer counts, reduce the whole buffer at once
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            return flat
        pos = 0
        for off, n in sorted(self.reduced):
            if off > pos:
                dist.all_reduce(flat[pos:off], op=dist.ReduceOp.SUM, group=self.group)
            pos = off + n
        if pos < flat.numel():
            dist.all_reduce(flat[pos:], op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.current_stream().wait_stream(self.comm)
        return flat


def _env_int(name, default):
    v = os.environ.get(name, "")
    return int(v) if v.strip() else default


_thin_groups = {}


def thin_group(max_ctas):
    """A second NCCL communicator over all ranks whose kernels use at most `max_ctas` CTAs: for the all-reduces that run
    UNDER backward.  The persistent tcgen05 kernels own whole SMs, so a wide collective launched beside them either waits
    for SMs or pushes CTAs of the next convolution into a second wave; a thin one on SMs the convolutions leave free
    (kernels.set_tc_sm_limit) does neither, and has the rest of the pass to finish in."""
    if max_ctas not in _thin_groups:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.min_ctas = 1
        opts.config.max_ctas = int(max_ctas)
        _thin_groups[max_ctas] = dist.new_group(backend="nccl", pg_options=opts)
    return _thin_groups[max_ctas]


def attach(optimizers, group=None):
    """Makes each Adam optimizer data-parallel: replicas start from rank 0's weights, gradients are averaged.

    On GPUs the knobs of the overlap (all read once, here; identical on every rank):
      MCG_DP_BUCKETS=1      large weight gradients are all-reduced under the rest of backward (GradBuckets)
      MCG_DP_THIN_CTAS=k    collectives that overlap backward go through a k-CTA communicator (0: the default one)
      MCG_DP_SM_RESERVE=k   SMs the persistent convolution kernels leave free for them (default: MCG_DP_THIN_CTAS)
    """
    w = dist.get_world_size(group) if dist.is_initialized() else 1
    on_gpu = w > 1 and torch.cuda.is_available() and dist.get_backend(group) == "nccl"
    buckets = on_gpu and _env_int("MCG_DP_BUCKETS", DEFAULT_BUCKETS) != 0
    thin_ctas = _env_int("MCG_DP_THIN_CTAS", DEFAULT_THIN_CTAS) if on_gpu else 0
    reserve = _env_int("MCG_DP_SM_RESERVE", thin_ctas) if on_gpu else 0
    early = thin_group(thin_ctas) if thin_ctas > 0 and group is None else group
    if on_gpu and reserve > 0:
        from . import kernels as K
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        K.set_tc_sm_limit(max(sms - reserve, 1))
    for opt in optimizers:
        arena = opt.target.arena()
        broadcast_(arena.data, 0, group)
        arena.refresh_bf16()
        if w > 1:
            # the image discriminator's all-reduce (after pass A) runs under all of pass B: overlapped as a whole
            flat_group = early if type(opt.target).__name__ == "ImageDiscriminator" else group
            if on_gpu and grad_comm_dtype() == "bf16" and arena.grad.numel() % 8 == 0:
                opt.grad_transform = Bf16GradAllReduce(arena, flat_group)
            else:
                opt.grad_transform = lambda g, _grp=flat_group: allreduce_mean_(g, _grp)
            opt.grad_scale = 1.0 / w
            if buckets and hasattr(arena, "offsets") and arena.data.is_cuda and flat_group is group:
                opt.grad_buckets = GradBuckets(opt, group, early_group=early)
                if opt.grad_buckets.big:
                    opt.grad_transform = opt.grad_buckets.finish
                else:
                    opt.grad_buckets = None
    return w


def describe():
    """The data-parallel configuration in force, for bench.py's `config`."""
    buckets = _env_int("MCG_DP_BUCKETS", DEFAULT_BUCKETS)
    # GradBuckets reduces slices of the fp32 gradient buffer in place: with buckets on, only the image discriminator's
    # whole-buffer all-reduce can still go through Bf16GradAllReduce (attach)
    comm = grad_comm_dtype() if not buckets else "fp32 (bucketed); image_dis %s" % grad_comm_dtype()
    return {"grad_comm_dtype": comm, "buckets": buckets, "thin_ctas": _env_int("MCG_DP_THIN_CTAS", DEFAULT_THIN_CTAS),
            "sm_reserve": _env_int("MCG_DP_SM_RESERVE", _env_int("MCG_DP_THIN_CTAS", DEFAULT_THIN_CTAS))}


def shard_seed(base_seed, rank):
    """rank r uses seed base+r for its clips, latents and noise."""
    return int(base_seed) + int(rank)


def replica_checksums(links):
    """Two position-sensitive integer checksums of each model's fp32 parameter arena (bit patterns, not values): equal
    on every rank iff the replicas are bit-identical.  torch ops on the arena — bookkeeping, not the step."""
    sums = []
    for link in links:
        bits = link.arena().data.view(torch.int32).to(torch.int64)
        idx = torch.arange(bits.numel(), device=bits.device, dtype=torch.int64) % 65521 + 1
        sums += [bits.sum(), (bits * idx).sum()]
    return torch.stack(sums)


def replicas_identical(links, group=None):
    """True iff every rank holds bit-identical parameters for all `links` (BatchNorm running statistics are local by
    design, SURVEY.md §8e, and not part of the check).  Collective: call on every rank."""
    mine = replica_checksums(links)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return True
    w = dist.get_world_size(group)
    if dist.get_backend(group) == "gloo":
        mine = mine.cpu()
    gathered = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(gathered, mine, group=group)
    return all(bool(torch.equal(gathered[0], g)) for g in gathered[1:])
