"""Typed wrappers over the C ABI (include/mcg.h): torch CUDA tensors in, kernels enqueued on torch's current stream.
torch is used for device memory and streams only — every operation here is a libmcg.so kernel."""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, BF16, F32, IMPL_SIMT, IMPL_TC, ConvGeom, check  # noqa: F401

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.uint8: _lib.U8}   # uint8 only as mcg_pack_video's source
_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16}


def lib():
    return _lib.load()


def dt_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("libmcg supports float32 / bfloat16 tensors, got %s" % t.dtype)


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.McgError("libmcg kernels need CUDA tensors (there is no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def conv_out(i, k, s, p):
    return (i + 2 * p - k) // s + 1


def make_geom(N, Cin, Cout, in_sp, ksize, stride, pad):
    """in_sp/ksize/stride/pad are (T,H,W) triples (use T=1,k=1,s=1,p=0 for 2-D)."""
    out_sp = tuple(conv_out(i, k, s, p) for i, k, s, p in zip(in_sp, ksize, stride, pad))
    return ConvGeom(N, Cin, Cout, *in_sp, *out_sp, *ksize, *stride, *pad)


_ws_cache = {}


def workspace(nbytes, device):
    """Per-(device, stream) scratch for the column reductions (allocated once; kernels never allocate).  ZERO-FILLED at
    allocation: the reductions keep their slot sums and ticket counter in it and hand it back zeroed (mcg.h)."""
    key = (device, torch.cuda.current_stream().cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def tc_ok(g):
    """Shapes the tcgen05 path takes: 64-multiple channels (implicit GEMM) or <= 16 input channels (im2col GEMM /
    GEMM + col2im dgrad for the 3-channel image layers)."""
    if os.environ.get("MCG_DISABLE_TC"):   # debugging aid: route every convolution through the CUDA-core kernel
        return False
    if max(g.sT, g.sH, g.sW) > 2 or g.kT * g.kH * g.kW > 64 or g.Cout % 64:
        return False
    return g.Cin % 64 == 0 or (g.Cin <= 16 and g.kW % 4 == 0)


def _conv_ws(g, impl, device):
    nb = lib().mcg_conv_workspace_bytes(C.byref(g), impl)
    if nb == 0:
        return None, 0
    buf = torch.empty(int(nb), dtype=torch.uint8, device=device)
    return buf, buf.numel()


COLS_VALID = 0x100


def w_rows(n):
    """MCG_W_ROWS(n): the weight tensor has only n rows, fewer than the (zero-padded) channel count (0 = no padding)."""
    return (int(n) & 0xffff) << 16


def conv_fprop(g, x, w, bias, y, impl, ws=None, cols_valid=False):
    """ws: optional caller-held workspace (kept alive to reuse the im2col of a small-Cin layer in wgrad)."""
    if ws is None:
        ws, nb = _conv_ws(g, impl, x.device)
    else:
        nb = ws.numel()
    check(lib().mcg_conv_fprop(C.byref(g), ptr(x), ptr(w), ptr(bias), ptr(y), dt_code(x), dt_code(y),
                               impl | (COLS_VALID if cols_valid else 0), ptr(ws), nb, stream()), "mcg_conv_fprop")
    return ws


def conv_dgrad(g, dy, w, bias, dx, impl, accumulate=False):
    ws, nb = _conv_ws(g, impl, dy.device)
    check(lib().mcg_conv_dgrad(C.byref(g), ptr(dy), ptr(w), ptr(bias), ptr(dx), dt_code(dy), dt_code(dx),
                               int(accumulate), impl, ptr(ws), nb, stream()), "mcg_conv_dgrad")


def conv_wgrad(g, x, dy, dw, impl, ws=None, cols_valid=False):
    assert dw.dtype == torch.float32
    if ws is None:
        ws, nb = _conv_ws(g, impl, x.device)
    else:
        nb = ws.numel()
    check(lib().mcg_conv_wgrad(C.byref(g), ptr(x), ptr(dy), ptr(dw), dt_code(x), impl | (COLS_VALID if cols_valid else 0),
                               ptr(ws), nb, stream()), "mcg_conv_wgrad")
    return ws


def bn_stats(y, M, Cc, gamma, beta, eps, decay, mean, invstd, scale, shift, avg_mean, avg_var):
    nb = lib().mcg_colreduce_workspace_bytes(M, Cc)
    ws = workspace(nb, y.device)
    check(lib().mcg_bn_stats(ptr(y), M, Cc, dt_code(y), ptr(gamma), ptr(beta), eps, decay, ptr(mean), ptr(invstd),
                             ptr(scale), ptr(shift), ptr(avg_mean), ptr(avg_var), ptr(ws), ws.numel(), stream()),
          "mcg_bn_stats")


def colsum(g, M, Cc, out, accumulate):
    nb = lib().mcg_colreduce_workspace_bytes(M, Cc)
    ws = workspace(nb, g.device)
    check(lib().mcg_colsum(ptr(g), M, Cc, dt_code(g), ptr(out), int(accumulate), ptr(ws), ws.numel(), stream()),
          "mcg_colsum")


def _noise_args(noise, strides):
    if noise is None:
        return None, 0, 0, 0
    assert noise.dtype == torch.float32
    return ptr(noise), strides[0], strides[1], strides[2]


def affine_act_noise(y, M, Cc, P, scale, shift, act, slope, sigma, noise, noise_strides, rng_state, call_id, out):
    np_, a, b, c = _noise_args(noise, noise_strides)
    check(lib().mcg_affine_act_noise(ptr(y), M, Cc, P, dt_code(y), ptr(scale), ptr(shift), act, slope, sigma, np_, a, b,
                                     c, ptr(rng_state), call_id, ptr(out), dt_code(out), stream()),
          "mcg_affine_act_noise")


def pack_video(src, N, Cc, T, H, W, strides, frame_ptr, sigma, noise, noise_strides, rng_state, call_id, out):
    np_, a, b, c = _noise_args(noise, noise_strides)
    check(lib().mcg_pack_video(ptr(src), dt_code(src), N, Cc, T, H, W, *[int(s) for s in strides], ptr(frame_ptr), sigma,
                               np_, a, b, c, ptr(rng_state), call_id, ptr(out), dt_code(out), stream()),
          "mcg_pack_video")


def act_bn_bwd_reduce(g, y, M, Cc, mean, invstd, scale, shift, act, slope, dgamma, dbeta, acc_dgamma, acc_dbeta):
    nb = lib().mcg_colreduce_workspace_bytes(M, Cc)
    ws = workspace(nb, g.device)
    check(lib().mcg_act_bn_bwd_reduce(ptr(g), ptr(y), M, Cc, dt_code(g), ptr(mean), ptr(invstd), ptr(scale), ptr(shift),
                                      act, slope, ptr(dgamma), ptr(dbeta), ptr(acc_dgamma), ptr(acc_dbeta), ptr(ws),
                                      ws.numel(), stream()), "mcg_act_bn_bwd_reduce")


def act_bn_bwd_apply(g, y, M, Cc, mean, invstd, gamma, scale, shift, act, slope, use_output, dgamma, dbeta, gy):
    assert g.dtype == y.dtype
    check(lib().mcg_act_bn_bwd_apply(ptr(g), ptr(y), M, Cc, dt_code(g), ptr(mean), ptr(invstd), ptr(gamma), ptr(scale),
                                     ptr(shift), act, slope, int(use_output), ptr(dgamma), ptr(dbeta), ptr(gy),
                                     dt_code(gy), stream()), "mcg_act_bn_bwd_apply")


def tanh_bwd_video(gv, gi, out_tn, N, T, HW, Cc, frame_ptr, g_tn):
    gref = gv if gv is not None else gi
    check(lib().mcg_tanh_bwd_video(ptr(gv), ptr(gi), dt_code(gref), ptr(out_tn), dt_code(out_tn), N, T, HW, Cc,
                                   ptr(frame_ptr), ptr(g_tn), dt_code(g_tn), stream()), "mcg_tanh_bwd_video")


def video_to_uint8(videos_phys, T, N, want_u8=True, grid_size=0):
    """videos_phys: the generator's output storage (T*N, 1, H, W, C).  Returns (u8 (T,N,C,H,W) | None, grid | None)."""
    B, _, H, W, Cc = videos_phys.shape
    assert B == T * N and videos_phys.is_contiguous()
    dev = videos_phys.device
    u8 = torch.empty((T, N, Cc, H, W), dtype=torch.uint8, device=dev) if want_u8 else None
    grid = torch.empty((T, Cc, grid_size * H, grid_size * W), dtype=torch.uint8, device=dev) if grid_size else None
    check(lib().mcg_video_to_uint8(ptr(videos_phys), dt_code(videos_phys), T, N, Cc, H, W, ptr(u8), ptr(grid), int(grid_size),
                                   stream()), "mcg_video_to_uint8")
    return u8, grid


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        assert t.dtype == torch.float32 and t.is_contiguous()
        arr[i] = t.data_ptr()
    return arr


def gru_forward(params12, labels, L, h0, eps, zc, T, N, H, Zc, z, cache):
    check(lib().mcg_gru_forward(_ptr_array(params12), ptr(labels), L, ptr(h0), ptr(eps), ptr(zc), T, N, H, Zc, ptr(z),
                                ptr(cache), stream()), "mcg_gru_forward")


def gru_backward(params12, grads12, labels, L, eps, cache, gz, T, N, H, Zc):
    check(lib().mcg_gru_backward(_ptr_array(params12), _ptr_array(grads12), ptr(labels), L, ptr(eps), ptr(cache), ptr(gz),
                                 T, N, H, Zc, stream()), "mcg_gru_backward")


def loss_dis(y_real, y_fake, t_real, t_fake, N, Cc, use_ce, loss, gy_real, gy_fake):
    check(lib().mcg_loss_dis(ptr(y_real), ptr(y_fake), ptr(t_real), ptr(t_fake), N, Cc, int(use_ce), ptr(loss),
                             ptr(gy_real), ptr(gy_fake), stream()), "mcg_loss_dis")


def loss_gen(y_i, y_v, t_fake, N, Cc, use_ce, loss, gy_i, gy_v):
    check(lib().mcg_loss_gen(ptr(y_i), ptr(y_v), ptr(t_fake), N, Cc, int(use_ce), ptr(loss), ptr(gy_i), ptr(gy_v),
                             stream()), "mcg_loss_gen")


def adam_step(p, g, m, v, p_bf16, alpha, beta1, beta2, eps, wd, grad_scale, t_ptr):
    check(lib().mcg_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), p.numel(), alpha, beta1, beta2, eps, wd,
                              grad_scale, ptr(t_ptr), stream()), "mcg_adam_step")


def fill_zero(t):
    """Zero-fills a contiguous tensor on the current stream (cudaMemsetAsync — no kernel)."""
    assert t.is_contiguous()
    check(lib().mcg_fill_zero(ptr(t), t.numel() * t.element_size(), stream()), "mcg_fill_zero")


def pad_channels(src, dst):
    """dst[..., :C] = src, dst[..., C:] = 0 for channels-last tensors of equal leading shape."""
    C_, Cp = src.shape[-1], dst.shape[-1]
    assert src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype and src.numel() // C_ == dst.numel() // Cp
    check(lib().mcg_pad_channels(ptr(src), ptr(dst), src.numel() // C_, C_, Cp, dt_code(src), stream()), "mcg_pad_channels")


def cast_bf16(src, dst):
    check(lib().mcg_cast_f32_to_bf16(ptr(src), ptr(dst), src.numel(), stream()), "mcg_cast_f32_to_bf16")


def cast_f32(src, dst):
    check(lib().mcg_cast_bf16_to_f32(ptr(src), ptr(dst), src.numel(), stream()), "mcg_cast_bf16_to_f32")


def step_state_new(seed, device):
    st = torch.zeros(8, dtype=torch.int32, device=device)
    check(lib().mcg_step_state_init(ptr(st), int(seed) & (2 ** 64 - 1), stream()), "mcg_step_state_init")
    return st


def step_advance(state, T):
    check(lib().mcg_step_advance(ptr(state), T, stream()), "mcg_step_advance")


def randn(out, sigma, state, call_id):
    check(lib().mcg_randn(ptr(out), out.numel(), sigma, ptr(state), call_id, stream()), "mcg_randn")


def randint(out, high, state, call_id):
    assert out.dtype == torch.int32
    check(lib().mcg_randint(ptr(out), out.numel(), high, ptr(state), call_id, stream()), "mcg_randint")


def int_add(t, delta):
    assert t.dtype == torch.int32
    check(lib().mcg_int_add(ptr(t), delta, stream()), "mcg_int_add")


def tc_error_flag(reset=True):
    return lib().mcg_tc_error_flag(int(reset))


def launch_count():
    return lib().mcg_launch_count()


def set_tc_sm_limit(sms):
    """SMs the persistent tcgen05 kernels may occupy (0 = all): the data-parallel layer leaves a few to NCCL's CTAs."""
    check(lib().mcg_set_tc_sm_limit(int(sms)), "mcg_set_tc_sm_limit")


def get_tc_sm_limit():
    return lib().mcg_get_tc_sm_limit()
