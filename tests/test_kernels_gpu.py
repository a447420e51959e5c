"""Per-kernel parity: every libmcg.so entry point against the NumPy oracle (float64) on seeded inputs.

Tolerances (north_star): fp32 path <= 1e-5, bf16/tcgen05 path <= 2e-2, both as max|a-b| / max|b| per tensor.
For the bf16 path the inputs handed to the oracle are first rounded to bf16, so what is measured is the kernel
(accumulation + output rounding), not the input quantisation.
"""
import numpy as np
import pytest
import torch

from oracle import chainer_ops as ops
from oracle import mocogan_ref as ref

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 2e-2


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mocogan_chainer_b200 import kernels
    kernels.lib()
    return kernels


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def to_cl(a, dtype):  # (N,C,*sp) numpy -> channels-last (N,T,H,W,C) cuda tensor
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if t.dim() == 4:
        t = t.unsqueeze(2)
    return t.permute(0, 2, 3, 4, 1).contiguous().to(dtype)


def from_cl(t, nd):  # channels-last tensor -> (N,C,*sp) numpy float64
    a = t.float().permute(0, 4, 1, 2, 3).cpu().numpy().astype(np.float64)
    return a[:, :, 0] if nd == 2 else a


def w_to_internal(W, dtype):  # (O,I,*k) -> (O,*k,I)
    t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(W, 1, -1))).cuda()
    return t.contiguous().to(dtype)


def w_from_internal(t):
    return np.moveaxis(t.float().cpu().numpy().astype(np.float64), -1, 1)


def bf16_round(a):
    return torch.from_numpy(np.asarray(a, np.float32)).bfloat16().float().numpy().astype(np.float64)


def tri(v, nd):
    return (1,) * (3 - nd) + tuple(v) if nd == 2 else tuple(v)


CONV_CASES = [
    # name, nd, N, Cin, Cout, in_sp, k, stride, pad
    ("di_dc1", 2, 3, 3, 16, (16, 16), (4, 4), (2, 2), (1, 1)),
    ("di_dc2", 2, 3, 64, 128, (16, 16), (4, 4), (2, 2), (1, 1)),
    ("di_dc5", 2, 5, 64, 7, (4, 4), (4, 4), (1, 1), (0, 0)),
    ("dv_dc1", 3, 2, 3, 8, (7, 8, 8), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("dv_dc2", 3, 2, 64, 64, (7, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("dv_dc5", 3, 3, 64, 1, (4, 4, 4), (4, 4, 4), (1, 3, 3), (0, 0, 0)),
    ("g_dc1", 2, 6, 16, 60, (4, 4), (4, 4), (1, 1), (0, 0)),  # deconv 1x1->4x4 seen as its conv
    ("g_dc1_rows", 2, 300, 32, 60, (4, 4), (4, 4), (1, 1), (0, 0)),  # many rows: the 8-row full-window kernels
    ("odd", 2, 2, 5, 9, (9, 7), (3, 3), (2, 1), (1, 0)),
]


def run_conv_case(K, case, impl, dtype, tol):
    name, nd, N, Cin, Cout, in_sp, k, s, p = case
    rng = np.random.default_rng(abs(hash(name)) % 2 ** 31)
    x = rng.standard_normal((N, Cin) + in_sp)
    W = rng.standard_normal((Cout, Cin) + k) * 0.1
    b = rng.standard_normal(Cout)
    if dtype == torch.bfloat16:
        x, W = bf16_round(x), bf16_round(W)
    y_ref = ops.conv_nd_fwd(x, W, b, s, p)
    gy = rng.standard_normal(y_ref.shape)
    if dtype == torch.bfloat16:
        gy = bf16_round(gy)
    gx_ref, gW_ref, _ = ops.conv_nd_bwd(x, W, gy, s, p)
    g = K.make_geom(N, Cin, Cout, tri(in_sp, nd), tri(k, nd), tri(s, nd) if nd == 3 else (1,) + tuple(s),
                    tri(p, nd) if nd == 3 else (0,) + tuple(p))
    xd, gyd = to_cl(x, dtype), to_cl(gy, dtype)
    wdt = torch.bfloat16 if impl == K.IMPL_TC else torch.float32
    wd = w_to_internal(W, wdt)
    bd = torch.from_numpy(b).float().cuda()
    yd = torch.empty((N, g.To, g.Ho, g.Wo, Cout), dtype=dtype, device="cuda")
    K.conv_fprop(g, xd, wd, bd, yd, impl)
    dxd = torch.empty_like(xd)
    K.conv_dgrad(g, gyd, wd, None, dxd, impl)
    # a deconvolution's forward is this data gradient plus its bias (net.py:110-114)
    b_in = rng.standard_normal(Cin)
    dxb = torch.empty_like(xd)
    K.conv_dgrad(g, gyd, wd, torch.from_numpy(b_in).float().cuda(), dxb, impl)
    dwd = torch.zeros(wd.shape, dtype=torch.float32, device="cuda")
    K.conv_wgrad(g, xd, gyd, dwd, impl)
    K.conv_wgrad(g, xd, gyd, dwd, impl)  # accumulates: expect exactly 2x
    torch.cuda.synchronize()
    assert K.tc_error_flag() == 0
    assert relerr(from_cl(yd, nd), y_ref) < tol, name
    assert relerr(from_cl(dxd, nd), gx_ref) < tol, name
    assert relerr(from_cl(dxb, nd), gx_ref + b_in.reshape((1, Cin) + (1,) * len(in_sp))) < tol, name
    assert relerr(w_from_internal(dwd), 2 * gW_ref) < tol, name


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_simt_fp32(K, case):
    run_conv_case(K, case, K.IMPL_SIMT, torch.float32, TOL_F32)


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[0] in ("di_dc1", "dv_dc1", "g_dc1", "g_dc1_rows")],
                         ids=["di_dc1", "dv_dc1", "g_dc1", "g_dc1_rows"])
def test_conv_simt_bf16_activations(K, case):
    run_conv_case(K, case, K.IMPL_SIMT, torch.bfloat16, TOL_BF16)


TC_CASES = [
    ("tc2d_s2", 2, 3, 64, 128, (16, 16), (4, 4), (2, 2), (1, 1)),
    ("tc2d_s2_256", 2, 5, 128, 256, (8, 8), (4, 4), (2, 2), (1, 1)),
    ("tc2d_8to4", 2, 9, 64, 64, (8, 8), (4, 4), (2, 2), (1, 1)),
    ("tc3d", 3, 2, 64, 64, (7, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("tc3d_last", 3, 3, 64, 128, (7, 8, 8), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("tc2d_s1", 2, 2, 128, 64, (9, 9), (3, 3), (1, 1), (1, 1)),
    # 3-channel image layers: im2col GEMM (fprop/wgrad) + narrow-N dgrad
    ("tc_di_dc1", 2, 3, 3, 64, (16, 16), (4, 4), (2, 2), (1, 1)),
    ("tc_dv_dc1", 3, 2, 3, 64, (7, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("tc_c1", 2, 3, 1, 64, (16, 16), (4, 4), (2, 2), (1, 1)),
    ("tc_di_odd", 2, 3, 3, 64, (10, 10), (4, 4), (2, 2), (1, 1)),     # W*C % 8 != 0: the scalar row-interleave pass
    # class-fused data gradient (Cin = 64, k4 s2 p1, >= 2 x 148 pixel blocks): 315 blocks = pairs + 19 blocks dealt out as
    # single-class units, temporal taps skipped at both ends of the clip; 460 blocks = odd per-CTA count (pair + single)
    ("tc3d_fused", 3, 10, 64, 64, (7, 48, 48), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("tc2d_fused_odd", 2, 230, 64, 128, (32, 32), (4, 4), (2, 2), (1, 1)),
]


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv_tcgen05_bf16(K, case):
    run_conv_case(K, case, K.IMPL_TC, torch.bfloat16, TOL_BF16)


TILE_SHAPES = [(1, 64), (2, 64), (4, 64), (1, 128), (2, 128), (1, 256), (2, 256)]


@pytest.mark.parametrize("mt,bn", TILE_SHAPES, ids=["%dx%d" % (m * 128, b) for m, b in TILE_SHAPES])
def test_conv_tcgen05_every_tile_shape(K, mt, bn, monkeypatch):
    """The persistent kernel picks its tile (MT x 128 rows, BN columns) from a cost model; here every shape is forced
    (MCG_TC_MT / MCG_TC_BN / MCG_TC_WMT / MCG_TC_WBN are read per call) on layers whose box counts leave remainders:
    partial last steps (nlive < MT), rows cut between CTAs, stream-K cuts inside a wgrad tile, dummy wgrad slabs."""
    monkeypatch.setenv("MCG_TC_MT", str(mt))
    monkeypatch.setenv("MCG_TC_BN", str(bn))
    monkeypatch.setenv("MCG_TC_WMT", str(min(mt, 2)))
    monkeypatch.setenv("MCG_TC_WBN", str(bn))
    for case in (("shape3d", 3, 3, 64, 256, (7, 10, 10), (4, 4, 4), (1, 2, 2), (0, 1, 1)),     # 4*5*5*3 = 300 px: 3 boxes
                 ("shape2d", 2, 11, 128, 256, (12, 12), (4, 4), (2, 2), (1, 1)),               # odd box counts per class
                 ("shape_k3", 2, 5, 64, 256, (9, 9), (3, 3), (1, 1), (1, 1))):                 # 9 taps: odd slab count
        run_conv_case(K, case, K.IMPL_TC, torch.bfloat16, TOL_BF16)


@pytest.mark.parametrize("r", [16, 32, 64])
def test_conv_tcgen05_temporal_tap_reuse(K, r, monkeypatch):
    """The opt-in temporal re-use kernel (MCG_TC_TR=1: one TMA box of BT + kT - 1 frames serves the kT temporal taps):
    every frame-row count R on 3-D layers with remainders in T, H and W.  (Off by default: see tc_conv.cu / DESIGN.md.)"""
    monkeypatch.setenv("MCG_TC_TR", "1")
    monkeypatch.setenv("MCG_TC_TR_R", str(r))
    for case in (("tr3d", 3, 3, 64, 128, (9, 12, 20), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
                 ("tr3d_wide", 3, 2, 128, 256, (7, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1))):
        run_conv_case(K, case, K.IMPL_TC, torch.bfloat16, TOL_BF16)


def test_conv_tc_padded_weight_rows(K):
    """MCG_W_ROWS: the generator's first layer (60 = dim_zc + dim_zm channels, net.py:44) on tcgen05 — dy is zero-padded
    to 64 channels, the weight / dw keep 60 rows; compared with the oracle convolution that has Cout = 60."""
    rng = np.random.default_rng(5)
    N, Cin, Cout, Cp = 70, 128, 60, 64
    x = bf16_round(rng.standard_normal((N, Cin, 4, 4)))
    W = bf16_round(rng.standard_normal((Cout, Cin, 4, 4)) * 0.1)
    gy = bf16_round(rng.standard_normal((N, Cout, 1, 1)))
    y_ref = ops.conv_nd_fwd(x, W, None, (1, 1), (0, 0))
    gx_ref, gW_ref, _ = ops.conv_nd_bwd(x, W, gy, (1, 1), (0, 0))
    g = K.make_geom(N, Cin, Cp, (1, 4, 4), (1, 4, 4), (1, 1, 1), (0, 0, 0))
    wr = K.w_rows(Cout)
    xd = to_cl(x, torch.bfloat16)
    gyd = torch.zeros((N, 1, 1, 1, Cp), dtype=torch.bfloat16, device="cuda")
    gyd[..., :Cout] = to_cl(gy, torch.bfloat16)
    wd = w_to_internal(W, torch.bfloat16)                       # 60 rows
    yd = torch.full((N, 1, 1, 1, Cp), 7.0, dtype=torch.bfloat16, device="cuda")
    K.conv_fprop(g, xd, wd, None, yd, K.IMPL_TC | wr)
    dxd = torch.empty_like(xd)
    K.conv_dgrad(g, gyd, wd, None, dxd, K.IMPL_TC | wr)
    guard = torch.full((Cout + 4,) + tuple(wd.shape[1:]), 3.0, dtype=torch.float32, device="cuda")
    guard[:Cout] = 0
    K.conv_wgrad(g, xd, gyd, guard, K.IMPL_TC | wr)             # rows 60..63 of the buffer must stay untouched
    torch.cuda.synchronize()
    assert K.tc_error_flag() == 0
    assert relerr(from_cl(yd[..., :Cout], 2), y_ref) < TOL_BF16 and float(yd[..., Cout:].abs().max()) == 0.0
    assert relerr(from_cl(dxd, 2), gx_ref) < TOL_BF16
    assert relerr(w_from_internal(guard[:Cout]), gW_ref) < TOL_BF16
    assert float((guard[Cout:] - 3.0).abs().max()) == 0.0


def test_conv_tc_rejects_unsupported(K):
    from mocogan_chainer_b200._lib import McgError
    g = K.make_geom(2, 24, 64, (1, 16, 16), (1, 4, 4), (1, 2, 2), (0, 1, 1))   # 24 channels: neither path takes it
    x = torch.zeros((2, 1, 16, 16, 24), dtype=torch.bfloat16, device="cuda")
    w = torch.zeros((64, 1, 4, 4, 24), dtype=torch.bfloat16, device="cuda")
    y = torch.zeros((2, 1, 8, 8, 64), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(McgError):
        K.conv_fprop(g, x, w, None, y, K.IMPL_TC)


# BASELINE config 2 (batch 35, n_filters 64): every convolution of the three networks at the size bench.py times, written
# as the conv geometry the kernels see (a generator deconvolution = the conv it is the data gradient of).
FULL_TC = [
    # name, N, Cin, Cout, in_sp (T,H,W), k, s, p
    ("Dv.dc1", 35, 3, 64, (16, 64, 64), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("Dv.dc2", 35, 64, 128, (13, 32, 32), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("Dv.dc3", 35, 128, 256, (10, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("Dv.dc4", 35, 256, 512, (7, 8, 8), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("Di.dc1", 35, 3, 64, (1, 64, 64), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("Di.dc2", 35, 64, 128, (1, 32, 32), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("Di.dc3", 35, 128, 256, (1, 16, 16), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("Di.dc4", 35, 256, 512, (1, 8, 8), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("G.dc1", 560, 512, 60, (1, 4, 4), (1, 4, 4), (1, 1, 1), (0, 0, 0)),      # 60 = dim_zc + dim_zm weight rows, MCG_W_ROWS
    ("G.dc2", 560, 256, 512, (1, 8, 8), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("G.dc3", 560, 128, 256, (1, 16, 16), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("G.dc4", 560, 64, 128, (1, 32, 32), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
    ("G.dc5", 560, 3, 64, (1, 64, 64), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
]
# measured on B200 (profiles/r02_parity_report.json): bf16 outputs sit within one bf16 rounding of the float64 result
# relative to the tensor's maximum, the fp32 weight gradient (atomics over up to 573,440 pixels) well inside 1e-3
FULL_TC_TOL = {"fprop": 6e-3, "dgrad": 6e-3, "wgrad": 1e-4}     # measured: <= 3.6e-3 / 3.6e-3 / 1.1e-5


@pytest.mark.parametrize("case", FULL_TC, ids=[c[0] for c in FULL_TC])
def test_conv_tc_full_size_vs_float64(K, case):
    """Every convolution of BASELINE config 2 at full size — the tile shapes pick_tile chooses there, incl. the im2col /
    merged-class paths of the 3-channel layers and the padded-row path of G.dc1 — against an independent float64 CPU
    convolution (torch conv2d / conv3d + autograd in float64 on the bf16-rounded operands)."""
    import torch.nn.functional as Fn
    name, N, Cin, Cout, in_sp, k, s, p = case
    Cp = (Cout + 63) // 64 * 64                      # G.dc1: activations zero-padded to 64 channels, weight keeps 60 rows
    wr = K.w_rows(Cout) if Cp != Cout else 0
    g = K.make_geom(N, Cin, Cp, in_sp, k, s, p)
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((N,) + in_sp + (Cin,), device="cuda", generator=gen).bfloat16()
    w = (torch.randn((Cout,) + k + (Cin,), device="cuda", generator=gen) * 0.05).bfloat16()
    gy = torch.zeros((N, g.To, g.Ho, g.Wo, Cp), device="cuda", dtype=torch.bfloat16)
    gy[..., :Cout] = torch.randn((N, g.To, g.Ho, g.Wo, Cout), device="cuda", generator=gen).bfloat16()
    y = torch.empty_like(gy)
    ws = K.conv_fprop(g, x, w, None, y, K.IMPL_TC | wr)
    dx = torch.empty_like(x)
    K.conv_dgrad(g, gy, w, None, dx, K.IMPL_TC | wr)
    dw = torch.zeros(w.shape, device="cuda")
    K.conv_wgrad(g, x, gy, dw, K.IMPL_TC | wr, ws=ws, cols_valid=ws is not None)
    torch.cuda.synchronize()
    assert K.tc_error_flag() == 0
    # float64 reference on the host: logical layouts (N,C,T,H,W) / (O,I,kT,kH,kW)
    x64 = x.cpu().double().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    w64 = w.cpu().double().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    gy64 = gy[..., :Cout].cpu().double().permute(0, 4, 1, 2, 3).contiguous()
    if in_sp[0] == 1 and k[0] == 1:
        y64 = Fn.conv2d(x64[:, :, 0], w64[:, :, 0], None, stride=s[1:], padding=p[1:]).unsqueeze(2)
    else:
        y64 = Fn.conv3d(x64, w64, None, stride=s, padding=p)
    gx64, gw64 = torch.autograd.grad(y64, (x64, w64), gy64)
    cl = lambda t: t.float().cpu().double().permute(0, 4, 1, 2, 3)
    errs = {"fprop": relerr(cl(y[..., :Cout]).numpy(), y64.detach().numpy()),
            "dgrad": relerr(cl(dx).numpy(), gx64.numpy()),
            "wgrad": relerr(cl(dw).numpy(), gw64.numpy())}
    if Cp != Cout:
        assert float(y[..., Cout:].abs().max()) == 0.0
    _report("conv_full_size", name, errs)
    for kind, e in errs.items():
        assert e < FULL_TC_TOL[kind], (name, kind, e)


def _report(section, name, errs):
    import json
    import os
    path = os.environ.get("MCG_PARITY_REPORT")
    if path:
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data.setdefault(section, {})[name] = errs
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)


# ---------------------------------------------------------------------------------------------- BN / activations
@pytest.mark.parametrize("dtype,tol", [(torch.float32, TOL_F32), (torch.bfloat16, TOL_BF16)])
@pytest.mark.parametrize("shape", [(4, 16, 1, 6, 6), (3, 64, 3, 8, 8),
                                   # BASELINE config 2, the sizes the one-wave occupancy-capped grids were tuned at:
                                   (560, 64, 1, 32, 32),     # G.bn4: M = 573,440 rows x 64 channels
                                   (560, 512, 1, 4, 4),      # G.bn1: M = 8,960 rows x 512 channels
                                   (35, 128, 10, 16, 16)],   # Dv.bn2: M = 89,600 rows x 128 channels
                         ids=["tiny16", "tiny64", "G.bn4", "G.bn1", "Dv.bn2"])
def test_bn_act_noise_forward_backward(K, dtype, tol, shape):
    rng = np.random.default_rng(5)
    N, Cc, T, H, W = shape
    y = rng.standard_normal(shape) * 1.7 + 0.3
    gamma, beta = 1 + 0.2 * rng.standard_normal(Cc), 0.1 * rng.standard_normal(Cc)
    noise = rng.standard_normal(shape)
    g_out = rng.standard_normal(shape)
    if dtype == torch.bfloat16:
        y, g_out = bf16_round(y), bf16_round(g_out)
    am, av = np.zeros(Cc), np.zeros(Cc)
    bn, stats = ops.batchnorm_fwd(y, gamma, beta, am, av)
    out_ref = ops.leaky_relu(bn, 0.2) + 0.2 * noise
    gpre = ops.leaky_relu_grad(bn, g_out, 0.2)
    gy_ref, gg_ref, gb_ref = ops.batchnorm_bwd(y, gamma + 0.05, stats, gpre)  # updated gamma, stale statistics

    M, P = N * T * H * W, T * H * W
    yd = to_cl(y, dtype)
    f = lambda a: torch.from_numpy(np.asarray(a)).float().cuda()
    gd, bd = f(gamma), f(beta)
    mean, invstd, scale, shift = (torch.empty(Cc, device="cuda") for _ in range(4))
    amd, avd = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    K.bn_stats(yd, M, Cc, gd, bd, ops.BN_EPS, ops.BN_DECAY, mean, invstd, scale, shift, amd, avd)
    nd_ = f(noise)  # reference layout (N,C,T,H,W): strides (C*P, P, 1)
    out = torch.empty_like(yd)
    K.affine_act_noise(yd, M, Cc, P, scale, shift, K.ACT_LRELU, 0.2, 0.2, nd_, (Cc * P, P, 1), None, 0, out)
    god = to_cl(g_out, dtype)
    dgam, dbet = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    acc_g, acc_b = torch.ones(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    K.act_bn_bwd_reduce(god, yd, M, Cc, mean, invstd, scale, shift, K.ACT_LRELU, 0.2, dgam, dbet, acc_g, acc_b)
    gyd = torch.empty_like(yd)
    K.act_bn_bwd_apply(god, yd, M, Cc, mean, invstd, f(gamma + 0.05), scale, shift, K.ACT_LRELU, 0.2, 0, dgam, dbet, gyd)
    torch.cuda.synchronize()
    stat_tol = 1e-5 if dtype == torch.float32 else 1e-5  # statistics accumulate in fp32/fp64 either way
    assert relerr(mean.cpu().numpy(), stats[0]) < stat_tol
    assert relerr(1 / invstd.cpu().numpy(), stats[1]) < stat_tol
    assert relerr(amd.cpu().numpy(), am) < 1e-5 and relerr(avd.cpu().numpy(), av) < 1e-5
    assert relerr(from_cl(out, 3), out_ref) < tol
    assert relerr(dgam.cpu().numpy(), gg_ref) < tol and relerr(dbet.cpu().numpy(), gb_ref) < tol
    assert relerr(acc_g.cpu().numpy(), gg_ref + 1) < tol and relerr(acc_b.cpu().numpy(), gb_ref + 1) < tol
    assert relerr(from_cl(gyd, 3), gy_ref) < tol


def test_colsum_and_small_channel_counts(K):
    rng = np.random.default_rng(0)
    for Cc in (1, 3, 7, 64):
        a = rng.standard_normal((5, Cc, 2, 4, 4))
        d = to_cl(a, torch.float32)
        out = torch.full((Cc,), 2.0, device="cuda")
        K.colsum(d, 5 * 2 * 16, Cc, out, True)
        torch.cuda.synchronize()
        assert relerr(out.cpu().numpy(), a.sum(axis=(0, 2, 3, 4)) + 2.0) < TOL_F32


def test_pack_video_frame_select_and_noise(K):
    rng = np.random.default_rng(3)
    N, Cc, T, H, W = 3, 3, 5, 6, 4
    x = rng.standard_normal((N, Cc, T, H, W)).astype(np.float32)
    noise = rng.standard_normal((N, Cc, H, W)).astype(np.float32)
    xd, nz = torch.from_numpy(x).cuda(), torch.from_numpy(noise).cuda()
    frame = torch.tensor([3], dtype=torch.int32, device="cuda")
    out = torch.empty((N, 1, H, W, Cc), device="cuda")
    K.pack_video(xd, N, Cc, T, H, W, (Cc * T * H * W, T * H * W, H * W, W, 1), frame, 0.2, nz, (Cc * H * W, H * W, 1), None, 0,
                 out)
    full = torch.empty((N, T, H, W, Cc), device="cuda", dtype=torch.bfloat16)
    K.pack_video(xd, N, Cc, T, H, W, (Cc * T * H * W, T * H * W, H * W, W, 1), None, 0.0, None, None, None, 0, full)
    torch.cuda.synchronize()
    assert relerr(from_cl(out, 3)[:, :, 0], x[:, :, 3] + 0.2 * noise) < TOL_F32
    assert relerr(from_cl(full, 3), x) < 1e-2


def test_philox_noise_moments_and_replay(K):
    st = K.step_state_new(1234, "cuda")
    a, b = torch.empty(1 << 20, device="cuda"), torch.empty(1 << 20, device="cuda")
    K.randn(a, 0.33, st, 1)
    K.randn(b, 0.33, st, 1)
    K.step_advance(st, 16)
    c = torch.empty(1 << 20, device="cuda")
    K.randn(c, 0.33, st, 1)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.mean().item()) < 2e-3 and abs(a.std().item() - 0.33) < 2e-3
    assert abs((a ** 4).mean().item() / a.var().item() ** 2 - 3.0) < 0.05  # Gaussian kurtosis
    assert 0 <= int(st[3].item()) < 16


# ---------------------------------------------------------------------------------------------- GRU, losses, Adam
@pytest.mark.parametrize("L", [0, 6])
def test_gru_forward_backward(K, L):
    rng = np.random.default_rng(9)
    T, N, H, Zc = 16, 35, 10, 50
    G = ref.ImageGenerator(Zc, H, L, 3, 4, T, rng=rng, dtype=np.float64)
    for k in list(G.params):
        if k.startswith("g0/") and k.endswith("/b"):
            G.params[k] += 0.1 * rng.standard_normal(H)
    lat = ref.ImageGenerator.draw_latents(rng, N, Zc, H, L, T, np.float64)
    g0 = {k[3:]: v for k, v in G.params.items() if k.startswith("g0/")}
    zl = np.eye(L)[lat["labels"]] if L else None
    h, caches, hs = lat["h0"], [], []
    for t in range(T):
        xt = np.concatenate((zl, lat["eps"][t]), 1) if L else lat["eps"][t]
        h, c = ops.gru_step_fwd(g0, h, xt)
        caches.append(c)
        hs.append(h)
    gz = rng.standard_normal((T, N, Zc + H))
    grads = {k: np.zeros_like(v) for k, v in g0.items()}
    gh = np.zeros((N, H))
    for t in range(T - 1, -1, -1):
        gh, _ = ops.gru_step_bwd(g0, caches[t], gh + gz[t, :, Zc:], grads)

    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
    order = ["W_r", "U_r", "W_z", "U_z", "W", "U"]
    p12 = [f(g0["%s/%s" % (n, wb)]) for n in order for wb in ("W", "b")]
    g12 = [torch.zeros_like(p) for p in p12]
    labels = torch.from_numpy(lat["labels"]).int().cuda() if L else None
    z = torch.empty((T * N, Zc + H), device="cuda")
    cache = torch.empty((T, N, 4, H), device="cuda")
    K.gru_forward(p12, labels, L, f(lat["h0"]), f(lat["eps"]), f(lat["zc"]), T, N, H, Zc, z, cache)
    K.gru_backward(p12, g12, labels, L, f(lat["eps"]), cache, f(gz.reshape(T * N, -1)), T, N, H, Zc)
    torch.cuda.synchronize()
    zz = z.cpu().numpy().reshape(T, N, Zc + H)
    assert relerr(zz[:, :, Zc:], np.stack(hs)) < TOL_F32
    assert relerr(zz[:, :, :Zc], np.tile(lat["zc"][None], (T, 1, 1))) < 1e-7
    i = 0
    for n in order:
        for wb in ("W", "b"):
            assert relerr(g12[i].cpu().numpy(), grads["%s/%s" % (n, wb)]) < 5e-5, (n, wb)
            i += 1


@pytest.mark.parametrize("Cc,use_ce", [(1, False), (7, True), (7, False)])
def test_losses(K, Cc, use_ce):
    rng = np.random.default_rng(2)
    N = 35
    yr, yf = rng.standard_normal((N, Cc, 1, 1, 1)) * 2, rng.standard_normal((N, Cc, 1, 1, 1)) * 2
    tr, tf = rng.integers(0, 6, N), rng.integers(0, 6, N)
    model = "infogan" if use_ce else "normal"
    l_ref, gr_ref, gf_ref = ref.loss_dis(model, "VideoDiscriminator", yr, yf, tr, tf)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()
    ti = lambda a: torch.from_numpy(a).int().cuda()
    loss, gr, gf = torch.empty(1, device="cuda"), torch.empty((N, Cc), device="cuda"), torch.empty((N, Cc), device="cuda")
    K.loss_dis(f(yr.reshape(N, Cc)), f(yf.reshape(N, Cc)), ti(tr), ti(tf), N, Cc, use_ce, loss, gr, gf)
    torch.cuda.synchronize()
    assert abs(loss.item() - l_ref) < 1e-5 * max(1, abs(l_ref))
    assert relerr(gr.cpu().numpy(), gr_ref.reshape(N, Cc)) < TOL_F32 and relerr(gf.cpu().numpy(), gf_ref.reshape(N, Cc)) < TOL_F32
    yi = rng.standard_normal((N, Cc, 1, 1)) * 2
    l_ref, gi_ref, gv_ref = ref.loss_gen(model, yi, yf, tf)
    gi, gv = torch.empty((N, Cc), device="cuda"), torch.empty((N, Cc), device="cuda")
    K.loss_gen(f(yi.reshape(N, Cc)), f(yf.reshape(N, Cc)), ti(tf), N, Cc, use_ce, loss, gi, gv)
    torch.cuda.synchronize()
    assert abs(loss.item() - l_ref) < 1e-5 * max(1, abs(l_ref))
    assert relerr(gi.cpu().numpy(), gi_ref.reshape(N, Cc)) < TOL_F32 and relerr(gv.cpu().numpy(), gv_ref.reshape(N, Cc)) < TOL_F32


def test_adam_weight_decay_three_steps(K):
    rng = np.random.default_rng(4)
    n = 10007
    p0 = rng.standard_normal(n).astype(np.float32)
    params = {"w": p0.astype(np.float64).copy()}
    st = ops.AdamState(params)
    pd = torch.from_numpy(p0.copy()).cuda()
    m, v = torch.zeros_like(pd), torch.zeros_like(pd)
    pb = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    t = torch.zeros(1, dtype=torch.int32, device="cuda")
    for step in range(3):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 1, n)).astype(np.float32)
        st.update(params, {"w": g.astype(np.float64)})
        K.int_add(t, 1)
        K.adam_step(pd, torch.from_numpy(g).cuda(), m, v, pb, 2e-4, 5e-5, 0.999, 1e-8, 1e-5, 1.0, t)
    torch.cuda.synchronize()
    assert np.abs(pd.cpu().numpy() - params["w"]).max() < 2e-6
    assert torch.equal(pb, pd.bfloat16())


def test_cast_round_trip_and_helpers(K):
    """mcg_cast_f32_to_bf16 (vector and scalar path) / mcg_cast_bf16_to_f32 (the bf16 gradient all-reduce's two passes),
    mcg_fill_zero, mcg_pad_channels — bit-exact against torch's conversions."""
    g = torch.Generator(device="cuda").manual_seed(3)
    for n in (8 * 1001, 1003):
        a = torch.randn(n, device="cuda", generator=g) * 3
        b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
        K.cast_bf16(a, b)
        assert torch.equal(b, a.bfloat16())
        if n % 8 == 0:
            c = torch.empty(n, device="cuda")
            K.cast_f32(b, c)
            assert torch.equal(c, b.float())
    z = torch.ones(4097, device="cuda")
    K.fill_zero(z)
    src = torch.randn((70, 1, 1, 1, 60), device="cuda", generator=g).bfloat16()
    dst = torch.full((70, 1, 1, 1, 64), 7.0, dtype=torch.bfloat16, device="cuda")
    K.pad_channels(src, dst)
    torch.cuda.synchronize()
    assert float(z.abs().max()) == 0.0
    assert torch.equal(dst[..., :60], src) and float(dst[..., 60:].abs().max()) == 0.0
