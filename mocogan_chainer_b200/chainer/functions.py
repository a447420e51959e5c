"""FunctionNode subclasses wrapping the libmcg.so C ABI — the operator layer of the MoCoGAN hot path.

Each node cites the Chainer operator / reference call site it stands in for.  Activations travel between nodes
as *logical* (N,C,H,W) / (N,C,T,H,W) torch views over channels-last (N,T,H,W,C) storage, so shapes and axis
orders seen by model code equal the reference's while kernels read their preferred layout with no copies.
"""
import numpy as np
import torch

from .. import kernels as K
from .. import random as mrandom
from . import FunctionNode, Parameter, Variable, VideoGrad, act_dtype, config

ACT_CODES = {None: K.ACT_NONE, "none": K.ACT_NONE, "relu": K.ACT_RELU, "leaky_relu": K.ACT_LRELU, "tanh": K.ACT_TANH}


# ---------------------------------------------------------------------------------------------- layout helpers
def logical_view(p5, nd):
    """physical (N,T,H,W,C) -> logical (N,C,T,H,W) (nd=3) or (N,C,H,W) (nd=2, T must be 1)."""
    v = p5.permute(0, 4, 1, 2, 3)
    return v[:, :, 0] if nd == 2 else v


def physical_view(x):
    """logical (N,C,H,W)/(N,C,T,H,W)/(N,C) tensor -> (N,T,H,W,C) view (not necessarily contiguous)."""
    if x.dim() == 2:
        x = x[:, :, None, None, None]
    elif x.dim() == 4:
        x = x.unsqueeze(2)
    return x.permute(0, 2, 3, 4, 1)


def as_physical(x, dtype=None):
    """Channels-last contiguous storage of a logical tensor in `dtype`; zero-copy when it already is."""
    dtype = dtype or x.dtype
    p = physical_view(x)
    if p.is_contiguous() and p.dtype == dtype:
        return p
    N, T, H, W, Cc = p.shape
    out = torch.empty((N, T, H, W, Cc), dtype=dtype, device=x.device)
    sn, st, sh, sw, sc = p.stride()
    K.pack_video(p, N, Cc, T, H, W, (sn, sc, st, sh, sw), None, 0.0, None, None, None, 0, out)
    return out


def _nd_of(x):
    return 2 if x.dim() in (2, 4) else 3


def _noise_kwargs(noise, out_shape_logical):
    """noise spec -> (sigma, tensor, strides, rng_state, call_id)."""
    if noise is None:
        return 0.0, None, None, None, 0
    kind = noise[0]
    if kind == "tensor":
        _, sigma, t = noise
        t = t.contiguous().float()
        if tuple(t.shape) != tuple(out_shape_logical):
            raise ValueError("injected noise shape %s != activation shape %s" % (tuple(t.shape), tuple(out_shape_logical)))
        Cc = t.shape[1]
        P = int(np.prod(t.shape[2:])) if t.dim() > 2 else 1
        return float(sigma), t, (Cc * P, P, 1), None, 0
    if kind == "philox":
        _, sigma, state, call_id = noise
        return float(sigma), None, None, state, int(call_id)
    raise ValueError("unknown noise spec %r" % (kind,))


# ---------------------------------------------------------------------------------------------- PackVideo
class PackVideo(FunctionNode):
    """x[:, :, t] (updater.py:97,107) / whole clip (updater.py:98,108) + add_noise on the network input
    (net.py:148,189) + cast to the compute dtype, in one gather pass over arbitrary strides — which also absorbs
    the (T,N,C,H,W)->(N,C,T,H,W) transpose of updater.py:102."""

    def __init__(self, frame=None, noise=None):
        super(PackVideo, self).__init__()
        self.frame, self.noise = frame, noise

    def forward(self, inputs):
        x, = inputs
        nd = _nd_of(x)
        p = physical_view(x)
        N, T, H, W, Cc = p.shape
        sn, st, sh, sw, sc = p.stride()
        frame_ptr = None
        if self.frame is not None:
            frame_ptr = self.frame if torch.is_tensor(self.frame) else torch.tensor([int(self.frame)], dtype=torch.int32,
                                                                                    device=x.device)
            self.frame_ptr = frame_ptr
        Tout = 1 if frame_ptr is not None else T
        out = torch.empty((N, Tout, H, W, Cc), dtype=act_dtype(), device=x.device)
        out_nd = 2 if (frame_ptr is not None or nd == 2) else 3
        logical = logical_view(out, out_nd)
        sigma, nt, nstr, state, cid = _noise_kwargs(self.noise, logical.shape)
        K.pack_video(p, N, Cc, T, H, W, (sn, sc, st, sh, sw), frame_ptr, sigma, nt, nstr, state, cid, out)
        return logical,

    def backward(self, idx, gys):
        g = as_physical(gys[0])
        if self.frame is not None:
            return VideoGrad(gi=g, frame_ptr=self.frame_ptr),
        return VideoGrad(gv=g),


def pack_video(x, frame=None, noise=None):
    return PackVideo(frame, noise).apply((x,))[0]


class ConcatLabelVideo(FunctionNode):
    """Updater.concat_label_video (updater.py:65-76, cgan): `dim_zl` planes of -1 with +1 at each clip's label plane,
    F.concat'ed to the clip on the channel axis.  Output storage is channels-last (N,T,H,W,C+dim_zl), so the
    discriminators' input pass reads it like any clip; backward hands the clip channels' slice of the (lazy) video
    gradient on to the generator (F.concat's backward, updater.py:104-106) — the label planes are constants.
    Layout plumbing only (torch copies, no arithmetic): cgan is not one of BASELINE.json's measured configurations."""

    def __init__(self, label, dim_zl):
        super(ConcatLabelVideo, self).__init__()
        self.label, self.dim_zl = label, int(dim_zl)

    def forward(self, inputs):
        x, = inputs
        if x.dim() != 5:
            raise ValueError("concat_label_video expects (N,C,T,H,W), got %s" % (tuple(x.shape),))
        p = physical_view(x)
        N, T, H, W, Cc = p.shape
        self.C = Cc
        odt = torch.float32 if x.dtype == torch.uint8 else x.dtype
        out = torch.empty((N, T, H, W, Cc + self.dim_zl), dtype=odt, device=x.device)
        if x.dtype == torch.uint8:   # clips from the uint8 cache: datasets.py:91's (v - 128) / 128
            out[..., :Cc].copy_((p.float() - 128.0) / 128.0)
        else:
            out[..., :Cc].copy_(p)
        lab = self.label.to(device=x.device, dtype=torch.int64)
        planes = torch.where(lab[:, None] == torch.arange(self.dim_zl, device=x.device)[None, :], 1.0, -1.0).to(odt)
        out[..., Cc:] = planes[:, None, None, None, :]
        return logical_view(out, 3),

    def backward(self, idx, gys):
        g = gys[0]
        Cc = self.C
        if isinstance(g, VideoGrad):
            cut = lambda a: None if a is None else a[..., :Cc].contiguous()
            return VideoGrad(cut(g.gv), cut(g.gi), g.frame_ptr, g.ops),
        return g[:, :Cc],


def concat_label_video(video, label, dim_zl):
    lab = label.data if isinstance(label, Variable) else label
    if not isinstance(video, Variable):
        video = Variable(video, requires_grad=False)
    return ConcatLabelVideo(lab, dim_zl).apply((video,))[0]


# ---------------------------------------------------------------------------------------------- convolutions
class ConvolutionND(FunctionNode):
    """F.convolution_2d / F.convolution_nd (net.py:149-156,190-197) when deconv=False;
    F.deconvolution_nd (net.py:110-114) when deconv=True.  Inputs (x, W, b); W, b are Parameters whose storage is
    (Cout_conv, kT, kH, kW, Cin_conv) — for a deconvolution that is exactly Chainer's (in, out, kh, kw) moved to
    channels-last, because a deconvolution's forward is the dgrad of the convolution with the same weight."""

    def __init__(self, stride, pad, deconv=False, out_dtype=None, bias_grad=True):
        super(ConvolutionND, self).__init__()
        self.stride, self.pad, self.deconv = tuple(stride), tuple(pad), deconv
        self.out_dtype, self.bias_grad = out_dtype, bias_grad

    @staticmethod
    def _tri(v, fill):
        v = tuple(v)
        return (fill,) * (3 - len(v)) + v

    def forward(self, inputs):
        x, W, b = inputs
        nd = _nd_of(x)
        self.nd = nd
        # bf16 mode: every convolution reads bf16 activations (the generator's fp32 latent z included), so gradients
        # flowing back never need a dtype conversion pass
        xdt = act_dtype() if config.compute_dtype == "bf16" else (x.dtype if x.dtype in (torch.float32, torch.bfloat16) else act_dtype())
        xp = as_physical(x, xdt)
        N, T, H, Wd, Cx = xp.shape
        ish = W.internal_shape  # (Cout_c, *k, Cin_c)
        k = self._tri(ish[1:-1], 1)
        s, p = self._tri(self.stride, 1), self._tri(self.pad, 0)
        if not self.deconv:
            if Cx != ish[-1]:
                raise ValueError("convolution: input has %d channels, weight expects %d" % (Cx, ish[-1]))
            g = K.make_geom(N, ish[-1], ish[0], (T, H, Wd), k, s, p)
            out_sp, Cout = (g.To, g.Ho, g.Wo), ish[0]
        else:
            if Cx != ish[0]:
                raise ValueError("deconvolution: input has %d channels, weight expects %d" % (Cx, ish[0]))
            in_sp = tuple(ss * (i - 1) + kk - 2 * pp for i, kk, ss, pp in zip((T, H, Wd), k, s, p))
            g = K.make_geom(N, ish[-1], ish[0], in_sp, k, s, p)  # the conv whose dgrad this deconv is
            out_sp, Cout = in_sp, ish[-1]
        self.w_rows = 0
        if self.deconv and config.compute_dtype == "bf16" and xp.dtype == torch.bfloat16 and Cx % 64 and not K.tc_ok(g):
            # the generator's first layer: 60 = dim_zc + dim_zm input channels (net.py:31,44).  The latent is zero-padded
            # to 64 channels so the layer runs on tcgen05; the weight keeps its 60 rows (MCG_W_ROWS: the missing rows are
            # TMA zero fill, the matching dw columns are never written).
            Cp = (Cx + 63) // 64 * 64
            gp = K.make_geom(N, ish[-1], Cp, in_sp, k, s, p)
            if K.tc_ok(gp):
                xpad = torch.empty((N, T, H, Wd, Cp), dtype=xp.dtype, device=xp.device)
                K.pad_channels(xp.contiguous(), xpad)
                xp, g, self.w_rows = xpad, gp, Cx
        self.g = g
        self.impl = K.IMPL_TC if (config.compute_dtype == "bf16" and xp.dtype == torch.bfloat16 and K.tc_ok(g)) else K.IMPL_SIMT
        w = W.bstore if self.impl == K.IMPL_TC else W.store
        odt = self.out_dtype or xp.dtype
        y = torch.empty((N,) + out_sp + (Cout,), dtype=odt, device=xp.device)
        bias = None if b is None else b.store
        self.cols_ws = None
        if not self.deconv:
            ws = K.conv_fprop(g, xp, w, bias, y, self.impl)
            if config.enable_backprop:
                self.cols_ws = ws   # small-Cin layers: keep x's im2col matrix for wgrad
        else:
            K.conv_dgrad(g, xp, w, bias, y, self.impl | K.w_rows(self.w_rows))
        self.xp = xp
        return logical_view(y, nd),

    def backward(self, idx, gys):
        x, W, b = self.inputs
        g = self.g
        gyp = as_physical(gys[0], self.xp.dtype)
        w = W.bstore if self.impl == K.IMPL_TC else W.store
        gx = None
        ws = None
        wr = K.w_rows(self.w_rows)
        if 0 in idx:
            dx = torch.empty_like(self.xp)
            if not self.deconv:
                K.conv_dgrad(g, gyp, w, None, dx, self.impl)
            else:
                ws = K.conv_fprop(g, gyp, w, None, dx, self.impl | wr)
            if self.w_rows:
                dx = dx[..., :self.w_rows]      # the padded channels carry no gradient
            gx = logical_view(dx, self.nd)
        if 1 in idx:
            if not self.deconv:
                K.conv_wgrad(g, self.xp, gyp, W.gstore, self.impl, ws=self.cols_ws, cols_valid=self.cols_ws is not None)
            else:
                K.conv_wgrad(g, gyp, self.xp, W.gstore, self.impl | wr, ws=ws, cols_valid=ws is not None)
            if W.grad_written_hook is not None:
                W.grad_written_hook(W)
        self.cols_ws = None
        if 2 in idx and b is not None and self.bias_grad:
            gb_src = as_physical(gys[0])  # original precision: the fp32 loss gradient of the last layer cancels heavily
            M = gb_src.numel() // gb_src.shape[-1]
            K.colsum(gb_src, M, gb_src.shape[-1], b.gstore, True)
        out = {0: gx, 1: True if 1 in idx else None, 2: True if 2 in idx else None}
        return tuple(out[i] for i in idx)


# ---------------------------------------------------------------------------------------------- BN + act + noise
class BNActNoise(FunctionNode):
    """L.BatchNormalization (train mode: batch statistics, running-stat update; eval mode: fixed statistics) +
    F.relu / F.leaky_relu(0.2) / F.tanh + add_noise of the NEXT layer (net.py:10-15) fused into
    statistics -> one elementwise pass.  Inputs (y,) or (y, gamma, beta)."""

    def __init__(self, bn=None, act=None, slope=0.2, noise=None, out_dtype=None):
        super(BNActNoise, self).__init__()
        self.bn, self.act, self.slope, self.noise, self.out_dtype = bn, ACT_CODES[act], float(slope), noise, out_dtype

    def forward(self, inputs):
        y = inputs[0]
        nd = _nd_of(y)
        self.nd = nd
        self.in_logical_shape = tuple(y.shape)
        yp = as_physical(y)
        Cc = yp.shape[-1]
        M = yp.numel() // Cc
        P = M // yp.shape[0]
        self.yp, self.M, self.C = yp, M, Cc
        dev = yp.device
        scale = shift = None
        self.mean = None
        if self.bn is not None:
            gamma, beta = inputs[1], inputs[2]
            bn = self.bn
            scale, shift = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
            if config.train:
                self.mean, self.invstd = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
                K.bn_stats(yp, M, Cc, gamma.store, beta.store, bn.eps, bn.decay, self.mean, self.invstd, scale, shift,
                           bn.avg_mean, bn.avg_var)
                # (the persistent N counts finetune-mode calls only in Chainer v3; the reference never finetunes, so
                # its checkpoints hold N = 0 — and so do these)
            else:  # F.fixed_batch_normalization with the running statistics (util.py:92 is the only caller)
                inv = torch.rsqrt(bn.avg_var + bn.eps)
                scale = gamma.store * inv
                shift = beta.store - bn.avg_mean * scale
                self.mean, self.invstd = bn.avg_mean.clone(), inv
        self.scale, self.shift = scale, shift
        odt = self.out_dtype or yp.dtype
        out = torch.empty(yp.shape, dtype=odt, device=dev)
        logical = logical_view(out, nd) if len(self.in_logical_shape) > 2 else out.reshape(self.in_logical_shape)
        sigma, nt, nstr, state, cid = _noise_kwargs(self.noise, self.in_logical_shape)
        K.affine_act_noise(yp, M, Cc, P, scale, shift, self.act, self.slope, sigma, nt, nstr, state, cid, out)
        if self.act == K.ACT_TANH:
            self.out_saved = out
        return logical,

    def backward(self, idx, gys):
        g = gys[0]
        yp, M, Cc = self.yp, self.M, self.C
        gy = torch.empty_like(yp)
        if isinstance(g, VideoGrad):
            return self._backward_video(g, gy),
        gp = as_physical(g, yp.dtype) if g.dim() > 2 else g.reshape(yp.shape).to(yp.dtype)
        if self.bn is not None:
            gamma, beta = self.inputs[1], self.inputs[2]
            dgam, dbet = torch.empty(Cc, device=yp.device), torch.empty(Cc, device=yp.device)
            want_param = 1 in idx or 2 in idx
            K.act_bn_bwd_reduce(gp, yp, M, Cc, self.mean, self.invstd, self.scale, self.shift, self.act, self.slope, dgam,
                                dbet, gamma.gstore if want_param else None, beta.gstore if want_param else None)
            if 0 in idx:
                K.act_bn_bwd_apply(gp, yp, M, Cc, self.mean, self.invstd, gamma.store, self.scale, self.shift, self.act,
                                   self.slope, 0, dgam, dbet, gy)
        elif 0 in idx:
            if self.act == K.ACT_TANH and self.out_saved.dtype == yp.dtype:
                K.act_bn_bwd_apply(gp, self.out_saved, M, Cc, None, None, None, None, None, self.act, self.slope, 1, None,
                                   None, gy)
            else:
                K.act_bn_bwd_apply(gp, yp, M, Cc, None, None, None, None, None, self.act, self.slope, 0, None, None, gy)
        gx = (logical_view(gy, self.nd) if len(self.in_logical_shape) > 2 else gy.reshape(self.in_logical_shape)) \
            if 0 in idx else None
        out = {0: gx, 1: True, 2: True}
        return tuple(out[i] for i in idx)

    def _backward_video(self, vg, gy):
        """gy = (gv + [t == frame] gi) * tanh'(out), rows re-ordered (n,t) -> (t,n): the generator's last layer."""
        if self.act != K.ACT_TANH or self.bn is not None:
            raise NotImplementedError("VideoGrad reaches a node other than the generator's tanh output")
        ops = vg.ops
        if not (len(ops) == 2 and ops[0] == ("transpose", (1, 2, 0, 3, 4)) and ops[1][0] == "reshape"):
            raise NotImplementedError("unexpected view chain between the generator output and the discriminators: %r" % (ops,))
        ref = vg.gv if vg.gv is not None else vg.gi
        N = ref.shape[0]
        B, _, H, W, Cc = self.yp.shape
        T = B // N
        K.tanh_bwd_video(vg.gv, vg.gi, self.out_saved, N, T, H * W, Cc, vg.frame_ptr, gy)
        return logical_view(gy, self.nd)


def bn_act_noise(y, bn=None, act=None, slope=0.2, noise=None, out_dtype=None):
    node = BNActNoise(bn, act, slope, noise, out_dtype)
    if bn is not None:
        return node.apply((y, bn.gamma, bn.beta))[0]
    return node.apply((y,))[0]


def relu(x):
    return bn_act_noise(x, act="relu")


def leaky_relu(x, slope=0.2):
    return bn_act_noise(x, act="leaky_relu", slope=slope)


def tanh(x):
    return bn_act_noise(x, act="tanh")


def cast(x, dtype):
    return bn_act_noise(x, out_dtype=dtype)


# ---------------------------------------------------------------------------------------------- views
class Transpose(FunctionNode):
    def __init__(self, axes):
        super(Transpose, self).__init__()
        self.axes = tuple(axes)

    def forward(self, inputs):
        return inputs[0].permute(*self.axes),

    def backward(self, idx, gys):
        g = gys[0]
        if isinstance(g, VideoGrad):
            return g.with_op(("transpose", self.axes)),
        inv = np.argsort(self.axes).tolist()
        return g.permute(*inv),


class Reshape(FunctionNode):
    def __init__(self, shape):
        super(Reshape, self).__init__()
        self.shape = tuple(shape)

    def forward(self, inputs):
        self.in_shape = tuple(inputs[0].shape)
        return inputs[0].reshape(self.shape),

    def backward(self, idx, gys):
        g = gys[0]
        if isinstance(g, VideoGrad):
            return g.with_op(("reshape", self.in_shape)),
        return g.reshape(self.in_shape),


def transpose(x, axes):
    return Transpose(axes).apply((x,))[0]


def reshape(x, shape):
    return Reshape(shape).apply((x,))[0]


# ---------------------------------------------------------------------------------------------- GRU
class GRUSequence(FunctionNode):
    """make_zm + zc tiling + concat (net.py:61-81,102-107): T steps of L.StatelessGRU in one persistent kernel.
    Inputs: the 12 GRU parameters (W_r.W, W_r.b, U_r.W, ... U.b).  Output z: (T*N, dim_zc + dim_zm) float32."""

    def __init__(self, labels, n_labels, h0, eps, zc):
        super(GRUSequence, self).__init__()
        self.labels, self.L, self.h0, self.eps, self.zc = labels, n_labels, h0, eps, zc

    def forward(self, inputs):
        T, N, H = self.eps.shape
        Zc = self.zc.shape[1]
        dev = self.eps.device
        z = torch.empty((T * N, Zc + H), device=dev)
        self.cache = torch.empty((T, N, 4, H), device=dev)
        K.gru_forward([p.store for p in inputs], self.labels, self.L, self.h0, self.eps, self.zc, T, N, H, Zc, z,
                      self.cache)
        self.dims = (T, N, H, Zc)
        return z,

    def backward(self, idx, gys):
        T, N, H, Zc = self.dims
        # the gradient arrives as a 60-channel slice of dc1's zero-padded bf16 storage: one gather pass makes it the
        # contiguous fp32 matrix the GRU kernel reads
        gz = as_physical(gys[0].reshape(T * N, Zc + H), torch.float32).reshape(T * N, Zc + H)
        K.gru_backward([p.store for p in self.inputs], [p.gstore for p in self.inputs], self.labels, self.L, self.eps,
                       self.cache, gz, T, N, H, Zc)
        return tuple(True for _ in idx)


# ---------------------------------------------------------------------------------------------- losses
class LossDis(FunctionNode):
    """Updater.loss_dis (updater.py:21-44), including the `[:1]` row slice: GAN terms read sample 0 only."""

    def __init__(self, t_real, t_fake, use_ce):
        super(LossDis, self).__init__()
        self.t_real, self.t_fake, self.use_ce = t_real, t_fake, use_ce

    def forward(self, inputs):
        yr, yf = inputs
        N, Cc = yr.shape[0], yr.shape[1]
        self.shapes = (tuple(yr.shape), tuple(yf.shape))
        yr2, yf2 = yr.reshape(N, Cc).float().contiguous(), yf.reshape(N, Cc).float().contiguous()
        loss = torch.empty((), device=yr.device)
        self.gr, self.gf = torch.empty_like(yr2), torch.empty_like(yf2)
        K.loss_dis(yr2, yf2, self.t_real, self.t_fake, N, Cc, self.use_ce, loss, self.gr, self.gf)
        return loss,

    def backward(self, idx, gys):
        out = {0: self.gr.reshape(self.shapes[0]), 1: self.gf.reshape(self.shapes[1])}
        return tuple(out[i] for i in idx)


class LossGen(FunctionNode):
    """Updater.loss_gen (updater.py:46-63)."""

    def __init__(self, t_fake, use_ce):
        super(LossGen, self).__init__()
        self.t_fake, self.use_ce = t_fake, use_ce

    def forward(self, inputs):
        yi, yv = inputs
        N, Cc = yi.shape[0], yi.shape[1]
        self.shapes = (tuple(yi.shape), tuple(yv.shape))
        yi2, yv2 = yi.reshape(N, Cc).float().contiguous(), yv.reshape(N, Cc).float().contiguous()
        loss = torch.empty((), device=yi.device)
        self.gi, self.gv = torch.empty_like(yi2), torch.empty_like(yv2)
        K.loss_gen(yi2, yv2, self.t_fake, N, Cc, self.use_ce, loss, self.gi, self.gv)
        return loss,

    def backward(self, idx, gys):
        out = {0: self.gi.reshape(self.shapes[0]), 1: self.gv.reshape(self.shapes[1])}
        return tuple(out[i] for i in idx)


def gan_loss_dis(y_real, y_fake, t_real=None, t_fake=None, use_ce=False):
    return LossDis(t_real, t_fake, use_ce).apply((y_real, y_fake))[0]


def gan_loss_gen(y_fake_i, y_fake_v, t_fake=None, use_ce=False):
    return LossGen(t_fake, use_ce).apply((y_fake_i, y_fake_v))[0]


def add_noise_spec(use_noise, sigma):
    """The noise spec a fused node needs to reproduce net.py:10-15 `add_noise` for the tensor it produces."""
    if not (config.train and use_noise):
        return None
    return mrandom.get_source().noise(sigma)
