"""Red zones around every device buffer the package allocates — the bounds check the kernels are run under in
tests/test_guard_gpu.py (compute-sanitizer is not available on the B200 pool, so the check is our own).

Inside `guarded_allocations()` the four allocation calls the package uses for kernel outputs and scratch
(`torch.empty`, `torch.zeros`, `torch.empty_like`, `torch.zeros_like`; the kernels themselves never allocate, mcg.h)
hand out the middle of a larger byte buffer whose first and last GUARD bytes hold a fixed pattern.  `check()` then
proves that no kernel wrote outside the tensor it was given: a box that walks past the end of its tensor, an epilogue
whose row predicate is off by one tile, a workspace sized for another dispatch path all land in a red zone.  Reads
outside a tensor are not detected (the parity tests catch those through wrong values).

`poison=True` additionally fills every `empty` / `empty_like` payload with 0xFF bytes (NaN as bf16 and fp32, -1 as
int32) before it is handed out: a kernel that consumes part of a buffer nobody wrote — padded rows or channels, a
scratch slot, a layout copy's slack — turns the step's losses and weights into NaN instead of depending on whatever
the caching allocator left there.
"""
import contextlib

import torch

GUARD = 64 * 1024        # bytes on each side; a multiple of 1024 keeps TMA's 128-byte base alignment
PATTERN = 0xA5


def dense(t):
    """True when t's elements tile one contiguous range exactly once (any dimension order)."""
    expect = 1
    for stride, size in sorted((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1):
        if stride != expect:
            return False
        expect *= size
    return True


class Guards(object):
    def __init__(self, cuda_only=True, poison=False):
        self.buffers = []        # (flat uint8 buffer, payload bytes, description)
        self.cuda_only = cuda_only
        self.poison = poison

    def wrap(self, proto, zero):
        """A tensor with proto's shape, strides and dtype inside a fresh red-zoned buffer."""
        if (self.cuda_only and not proto.is_cuda) or proto.numel() == 0 or not dense(proto):
            return proto
        nbytes = proto.numel() * proto.element_size()
        body = (nbytes + 1023) // 1024 * 1024
        flat = self._empty(GUARD + body + GUARD, dtype=torch.uint8, device=proto.device)
        flat[:GUARD].fill_(PATTERN)
        flat[GUARD + nbytes:].fill_(PATTERN)           # the round-up slack is red zone too
        mid = flat[GUARD:GUARD + nbytes].view(proto.dtype)
        if zero:
            mid.zero_()
        elif self.poison:
            flat[GUARD:GUARD + nbytes].fill_(0xFF)
        self.buffers.append((flat, nbytes, "%s %s" % (tuple(proto.shape), proto.dtype)))
        return mid.as_strided(proto.shape, proto.stride())

    def check(self):
        """Raises AssertionError naming the first buffers whose red zones were written."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        bad = []
        for flat, nbytes, what in self.buffers:
            lo = int((flat[:GUARD] != PATTERN).sum())
            hi = int((flat[GUARD + nbytes:] != PATTERN).sum())
            if lo or hi:
                bad.append("%s: %d bytes below, %d bytes above" % (what, lo, hi))
        n = len(self.buffers)
        self.buffers = []
        assert not bad, "writes outside %d of %d buffers: %s" % (len(bad), n, "; ".join(bad[:8]))
        return n


@contextlib.contextmanager
def guarded_allocations(cuda_only=True, poison=False):
    g = Guards(cuda_only, poison)
    orig = {name: getattr(torch, name) for name in ("empty", "zeros", "empty_like", "zeros_like")}
    g._empty = orig["empty"]

    def patched(name, zero):
        def alloc(*args, **kwargs):
            if kwargs.get("out") is not None or kwargs.get("pin_memory"):
                return orig[name](*args, **kwargs)
            return g.wrap(orig[name](*args, **kwargs), zero)
        return alloc

    try:
        for name in orig:
            setattr(torch, name, patched(name, name.startswith("zeros")))
        yield g
    finally:
        for name, fn in orig.items():
            setattr(torch, name, fn)
