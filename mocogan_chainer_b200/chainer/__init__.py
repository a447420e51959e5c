"""A minimal Chainer-v3-shaped host surface (Variable / FunctionNode / Link / Chain / Optimizer) for the MoCoGAN
hot path.  Chainer itself cannot be imported in this environment, so this package provides the
classes `model/net.py` and `model/updater.py` of the reference touch — same names, same call signatures — with
every array operation routed to libmcg.so.  Arrays are torch CUDA tensors (device memory / streams only).

Reference semantics restated here come from Chainer v3.1.0 (requirements.txt:1); see SURVEY.md App. A.
"""
import contextlib
import heapq
import weakref

import numpy as np
import torch

from .. import kernels as K

__version__ = "3.1.0-mcg"


# ------------------------------------------------------------------------------------------------ configuration
class _Config(object):
    train = True
    enable_backprop = True
    # 'bf16': activations bf16, tcgen05 kernels where the shape allows; 'fp32': strict fp32 CUDA-core path
    compute_dtype = "bf16"
    # True: every node remembers the CUDA stream its forward ran on and Variable.backward() replays its backward on the
    # same stream, with events carrying the gradients across streams — independent branches of the step (real / fake
    # clips, image / video discriminator, generator) then overlap on the device.  Set by the Updater.
    branch_streams = False


config = _Config()
global_config = config


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)


def no_backprop_mode():
    return using_config("enable_backprop", False)


def act_dtype():
    return torch.bfloat16 if config.compute_dtype == "bf16" else torch.float32


# ------------------------------------------------------------------------------------------------ video gradients
class VideoGrad(object):
    """Lazy gradient of a video Variable: `gv` is d/d(video) in channels-last (N,T,H,W,C) storage, `gi` the
    gradient of the single frame *frame_ptr (N,1,H,W,C).  Kept lazy so the frame scatter, the (N,C,T,H,W) ->
    (T,N,C,H,W) transpose and tanh' fuse into one kernel (updater.py:102,107 + net.py:114-115 backward)."""

    def __init__(self, gv=None, gi=None, frame_ptr=None, ops=()):
        self.gv, self.gi, self.frame_ptr, self.ops = gv, gi, frame_ptr, tuple(ops)

    def with_op(self, op):
        return VideoGrad(self.gv, self.gi, self.frame_ptr, self.ops + (op,))

    def __add__(self, other):
        if not isinstance(other, VideoGrad) or other.ops != self.ops:
            raise TypeError("cannot accumulate VideoGrad with %r" % type(other))
        if (self.gv is not None and other.gv is not None) or (self.gi is not None and other.gi is not None):
            raise NotImplementedError("VideoGrad: two gradients of the same kind")
        return VideoGrad(self.gv if self.gv is not None else other.gv, self.gi if self.gi is not None else other.gi,
                         self.frame_ptr if self.frame_ptr is not None else other.frame_ptr, self.ops)

    __radd__ = __add__


_ONES = {}


def _accumulate(a, b):
    if a is None:
        return b
    if isinstance(a, VideoGrad) or isinstance(b, VideoGrad):
        return a + b
    return a + b  # dense torch tensors (rare: parameters accumulate inside kernels instead)


# ------------------------------------------------------------------------------------------------ Variable
class Variable(object):
    def __init__(self, data=None, name=None, grad=None, requires_grad=True):
        self.data = data
        self.name = name
        self._grad = grad
        self.requires_grad = requires_grad
        self.creator_node = None
        self.rank = 0
        self.stop = False  # backward() does not propagate into or past this variable while True

    array = property(lambda self: self.data)
    creator = property(lambda self: self.creator_node)

    @property
    def grad(self):
        return self._grad

    @grad.setter
    def grad(self, g):
        self._grad = g

    shape = property(lambda self: tuple(self.data.shape))
    ndim = property(lambda self: self.data.dim())
    dtype = property(lambda self: self.data.dtype)
    size = property(lambda self: self.data.numel())

    def __len__(self):
        return self.data.shape[0]

    def cleargrad(self):
        self._grad = None

    def set_creator_node(self, node):
        self.creator_node = node
        self.rank = node.rank + 1

    def unchain(self):
        self.creator_node = None

    def transpose(self, *axes):
        from . import functions as F
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        return F.transpose(self, axes)

    def reshape(self, *shape):
        from . import functions as F
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return F.reshape(self, shape)

    def backward(self, retain_grad=False):
        """Reverse-mode sweep in creator-rank order (Chainer's Variable.backward).  A scalar loss seeds 1."""
        if self.creator_node is None:
            return
        if self._grad is None:
            # a scalar loss seeds 1; the constant is made once per device (no fill kernel per backward pass)
            if self.data.dim() == 0 and self.data.dtype == torch.float32:
                key = str(self.data.device)
                if key not in _ONES:
                    _ONES[key] = torch.ones((), device=self.data.device)
                self._grad = _ONES[key]
            else:
                self._grad = torch.ones_like(self.data)
        grads = {id(self): self._grad}
        keep = {id(self): self}
        heap, seen = [], set()

        def push(node):
            if id(node) not in seen:
                seen.add(id(node))
                heapq.heappush(heap, (-node.rank, len(seen), node))

        memo = {}

        def needs(v):
            """True iff some un-stopped, grad-requiring leaf is reachable below v (dead dgrads are never launched)."""
            k = id(v)
            if k not in memo:
                if not v.requires_grad or v.stop:
                    memo[k] = False
                elif v.creator_node is None:
                    memo[k] = True
                else:
                    memo[k] = False  # cycle guard; graphs are DAGs
                    memo[k] = any(needs(i) for i in v.creator_node.inputs)
            return memo[k]

        # Branch concurrency: a node whose forward ran on stream s runs its backward on s; the gradient of a variable
        # carries the (stream, event) pairs of the kernels that produced it, and a consumer on another stream waits
        # for them.  Parameter gradients accumulate inside kernels (red.add / atomics), so only the caller has to wait
        # for every stream used — done once at the end.
        multi = config.branch_streams and torch.cuda.is_available()
        home = torch.cuda.current_stream() if multi else None
        ready = {}       # id(variable) -> [(stream, event)] of its gradient's producers
        used = {}        # streams other than `home` that ran backward work
        carry = []       # gradients that crossed streams: kept alive until every stream has been joined

        push(self.creator_node)
        while heap:
            _, _, node = heapq.heappop(heap)
            outs = [o() for o in node.outputs]
            gys = tuple(None if o is None else grads.get(id(o)) for o in outs)
            if all(g is None for g in gys):
                continue
            idx = tuple(i for i, x in enumerate(node.inputs) if needs(x))
            if not idx:
                continue
            if multi:
                s = node.stream if node.stream is not None else home
                for o, g in zip(outs, gys):
                    if o is None or g is None:
                        continue
                    for (ps, ev) in ready.get(id(o), ()):
                        if ps != s:
                            s.wait_event(ev)
                            carry.append(g)
                    if id(o) not in ready and s != home:
                        s.wait_stream(home)      # the seed gradient lives on the caller's stream
                if s != home:
                    used[s.cuda_stream] = s
                with torch.cuda.stream(s):
                    gxs = node.backward(idx, gys)
                    ev = torch.cuda.Event()
                    ev.record(s)
            else:
                gxs = node.backward(idx, gys)
            for i, gx in zip(idx, gxs):
                if gx is None:
                    continue
                x = node.inputs[i]
                if isinstance(x, Parameter) and x.accumulates_in_kernel:
                    pass  # the kernel added straight into the parameter's .grad storage
                else:
                    prev = grads.get(id(x))
                    if multi and torch.is_tensor(prev) and torch.is_tensor(gx):
                        # two dense gradients of one variable, possibly produced on different streams: the sum is a
                        # kernel too, so it runs on this node's stream AFTER the events of everything added so far,
                        # and the sum's own event replaces them (none of the models reaches this: their only fan-out
                        # is the lazy VideoGrad, which is a tuple of references, not arithmetic)
                        with torch.cuda.stream(s):
                            for (ps, pev) in ready.get(id(x), ()):
                                if ps != s:
                                    s.wait_event(pev)
                            total = prev + gx
                            ev_sum = torch.cuda.Event()
                            ev_sum.record(s)
                        prev.record_stream(s)
                        carry.extend((prev, gx))
                        grads[id(x)] = total
                        ready[id(x)] = [(s, ev_sum)]
                    else:
                        grads[id(x)] = _accumulate(prev, gx)
                        if multi:
                            ready.setdefault(id(x), []).append((s, ev))
                    keep[id(x)] = x
                    if x.creator_node is None or retain_grad:
                        x._grad = grads[id(x)]
                if x.creator_node is not None:
                    push(x.creator_node)
            if not retain_grad:
                for o in outs:
                    if o is not None and o is not self:
                        grads.pop(id(o), None)
                        ready.pop(id(o), None)
        if multi:
            for s in used.values():
                home.wait_stream(s)
            del carry


class Parameter(Variable):
    """A trainable array.  After its Link tree has been packed into an arena (Link.arena()), `.data` / `.grad` are
    views in the reference's logical layout over flat fp32 storage kept in the kernels' layout
    ((Cout, kT, kH, kW, Cin) for convolution weights)."""
    accumulates_in_kernel = True

    def __init__(self, array, channels_last_weight=False):
        super(Parameter, self).__init__(None, requires_grad=True)
        self._init_array = np.asarray(array, dtype=np.float32)
        self.logical_shape = tuple(self._init_array.shape)
        self.cl = bool(channels_last_weight and self._init_array.ndim >= 3)
        self._store = self._gstore = self._bstore = None
        self.update_rule = None
        self.grad_written_hook = None   # data-parallel layer: called after a kernel accumulated into this .grad

    @property
    def internal_shape(self):
        s = self.logical_shape
        return (s[0],) + s[2:] + (s[1],) if self.cl else s

    def _to_logical(self, t):
        if not self.cl:
            return t
        nd = t.dim()
        return t.permute(0, nd - 1, *range(1, nd - 1))

    def _internal_init(self):
        a = self._init_array
        return np.ascontiguousarray(np.moveaxis(a, 1, -1)) if self.cl else a

    def _bind(self, store, gstore, bstore):
        self._store, self._gstore, self._bstore = store, gstore, bstore

    @property
    def data(self):
        if self._store is None:
            return None
        return self._to_logical(self._store)

    @data.setter
    def data(self, value):
        if value is None:
            return
        if self._store is None:
            self._init_array = np.asarray(value.detach().cpu().numpy() if torch.is_tensor(value) else value, np.float32)
            return
        src = value if torch.is_tensor(value) else torch.from_numpy(np.asarray(value, np.float32))
        self._to_logical(self._store).copy_(src.to(self._store.device))
        if self._bstore is not None:
            K.cast_bf16(self._store, self._bstore)

    @property
    def grad(self):
        return None if self._gstore is None else self._to_logical(self._gstore)

    @grad.setter
    def grad(self, g):
        if g is None or self._gstore is None:
            return
        self._to_logical(self._gstore).copy_(g)

    # kernel-facing raw views
    store = property(lambda self: self._store)      # fp32 master, internal layout
    gstore = property(lambda self: self._gstore)    # fp32 gradient, internal layout
    bstore = property(lambda self: self._bstore)    # bf16 copy of the master, internal layout

    shape = property(lambda self: self.logical_shape)
    size = property(lambda self: int(np.prod(self.logical_shape)))
    ndim = property(lambda self: len(self.logical_shape))
    dtype = property(lambda self: torch.float32)

    def cleargrad(self):
        if self._gstore is not None:
            self._gstore.zero_()


# ------------------------------------------------------------------------------------------------ FunctionNode
class FunctionNode(object):
    """Chainer v3 FunctionNode surface: apply / forward / backward / retain_inputs / retain_outputs."""

    def __init__(self):
        self.inputs = ()
        self.outputs = ()
        self.stream = None
        self.rank = 0
        self._retain_in = None
        self._retain_out = ()
        self._retained_out = ()

    def apply(self, inputs):
        inputs = tuple(x if isinstance(x, Variable) else Variable(x, requires_grad=False) for x in inputs)
        in_data = tuple(x.data if not isinstance(x, Parameter) else x for x in inputs)
        self.check_type_forward(in_data)
        outs = self.forward(in_data)
        if not isinstance(outs, tuple):
            outs = (outs,)
        need = config.enable_backprop and any(x.requires_grad for x in inputs)
        ret = tuple(Variable(o, requires_grad=need) for o in outs)
        if need:
            self.inputs = inputs
            # the stream this node's forward was enqueued on; backward() runs it there again (branch concurrency)
            self.stream = torch.cuda.current_stream() if (config.branch_streams and torch.cuda.is_available()) else None
            self.rank = max([x.rank for x in inputs] + [0])
            for v in ret:
                v.set_creator_node(self)
            self.outputs = tuple(weakref.ref(v) for v in ret)
            self._retained_out = tuple(ret[i].data for i in self._retain_out)
        return ret

    def check_type_forward(self, in_data):
        pass

    def forward(self, inputs):
        raise NotImplementedError

    def backward(self, target_input_indexes, grad_outputs):
        raise NotImplementedError

    def retain_inputs(self, indexes):
        self._retain_in = tuple(indexes)

    def retain_outputs(self, indexes):
        self._retain_out = tuple(indexes)

    def get_retained_inputs(self):
        idx = range(len(self.inputs)) if self._retain_in is None else self._retain_in
        return tuple(self.inputs[i] for i in idx)

    def get_retained_outputs(self):
        return self._retained_out


# ------------------------------------------------------------------------------------------------ Link / Chain
class ParamArena(object):
    """All parameters of one model in three flat buffers (fp32 master, fp32 grad, bf16 copy) so that cleargrads,
    the gradient all-reduce and Adam+WeightDecay are each ONE operation per model."""

    def __init__(self, named_params, device):
        self.names = [n for n, _ in named_params]
        self.params = [p for _, p in named_params]
        sizes = [p.size for p in self.params]
        pad = lambda n: (n + 7) // 8 * 8  # keep every slice 32-byte aligned (vector loads, TMA base alignment 16 B)
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += pad(n)
        self.n = tot
        self.offsets = {id(p): (o, n) for p, o, n in zip(self.params, offs, sizes)}   # slice of each parameter
        # 128-byte aligned tails for TMA: pad whole arena
        self.data = torch.zeros(tot, dtype=torch.float32, device=device)
        self.grad = torch.zeros(tot, dtype=torch.float32, device=device)
        self.bf16 = torch.zeros(tot, dtype=torch.bfloat16, device=device)
        host = np.zeros(tot, dtype=np.float32)
        for p, o, n in zip(self.params, offs, sizes):
            host[o:o + n] = p._internal_init().reshape(-1)
        self.data.copy_(torch.from_numpy(host))
        K.cast_bf16(self.data, self.bf16)
        for p, o, n in zip(self.params, offs, sizes):
            ish = p.internal_shape
            p._bind(self.data[o:o + n].view(ish), self.grad[o:o + n].view(ish), self.bf16[o:o + n].view(ish))

    def refresh_bf16(self):
        K.cast_bf16(self.data, self.bf16)


class Link(object):
    def __init__(self):
        self._params = []
        self._persistent = []
        self._children = []
        self._within_init_scope = False
        self._arena = None
        self.name = None

    @contextlib.contextmanager
    def init_scope(self):
        old = self._within_init_scope
        self._within_init_scope = True
        try:
            yield
        finally:
            self._within_init_scope = old

    def __setattr__(self, name, value):
        if getattr(self, "_within_init_scope", False):
            if isinstance(value, Parameter):
                self._params.append(name)
            elif isinstance(value, Link):
                self._children.append(name)
                if value.name is None:
                    object.__setattr__(value, "name", name)
        object.__setattr__(self, name, value)

    def add_persistent(self, name, value):
        self._persistent.append(name)
        object.__setattr__(self, name, value)

    register_persistent = lambda self, name: self._persistent.append(name)

    def children(self):
        for n in self._children:
            yield getattr(self, n)

    def links(self, skipself=False):
        if not skipself:
            yield self
        for c in self.children():
            for l in c.links():
                yield l

    def namedparams(self, include_uninit=True):
        for n in sorted(self._params):
            yield "/" + n, getattr(self, n)
        for cn in sorted(self._children):
            for path, p in getattr(self, cn).namedparams():
                yield "/" + cn + path, p

    def params(self, include_uninit=True):
        for _, p in self.namedparams():
            yield p

    def namedpersistents(self):
        for n in sorted(self._persistent):
            yield "/" + n, self, n
        for cn in sorted(self._children):
            for path, link, n in getattr(self, cn).namedpersistents():
                yield "/" + cn + path, link, n

    def arena(self, device="cuda"):
        if self._arena is None:
            if not torch.cuda.is_available():
                raise K._lib.McgError("mocogan_chainer_b200 needs a CUDA device: there is no CPU fallback")
            self._arena = ParamArena(list(self.namedparams()), device)
            for _, link, n in self.namedpersistents():
                v = getattr(link, n)
                if isinstance(v, np.ndarray):
                    object.__setattr__(link, n, torch.from_numpy(v.astype(np.float32)).to(device))
        return self._arena

    def to_gpu(self, device=None):
        self.arena()
        return self

    def to_cpu(self):
        return self  # parameters live on the device; serializers copy to the host when saving

    def cleargrads(self):
        """Chainer sets grads to None and lets backward allocate; here the gradient arena is zero-filled (one
        memset) because the wgrad kernels accumulate in place."""
        K.fill_zero(self.arena().grad)

    zerograds = cleargrads

    def count_params(self):
        return sum(p.size for p in self.params())

    def serialize(self, serializer):
        for path, p in self.namedparams():
            serializer(path.lstrip("/"), p)
        for path, link, n in self.namedpersistents():
            serializer(path.lstrip("/"), (link, n))


class Chain(Link):
    def __getitem__(self, name):
        return getattr(self, name)


class ChainList(Link):
    pass


# ------------------------------------------------------------------------------------------------ reporter
_report_sink = {}


def report(values, observer=None):
    """chainer.report: records '<observer name>/<key>' -> value (updater.py:40,59)."""
    prefix = ""
    if observer is not None:
        prefix = getattr(observer, "report_name", None) or getattr(observer, "name", observer.__class__.__name__)
        prefix += "/"
    for k, v in values.items():
        _report_sink[prefix + k] = v


def get_report():
    return _report_sink


from . import cuda  # noqa: E402,F401
from . import dataset  # noqa: E402,F401
from . import functions  # noqa: E402,F401
from . import initializers  # noqa: E402,F401
from . import iterators  # noqa: E402,F401
from . import links  # noqa: E402,F401
from . import optimizer  # noqa: E402,F401
from . import optimizers  # noqa: E402,F401
from . import serializers  # noqa: E402,F401
from . import training  # noqa: E402,F401
