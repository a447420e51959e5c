#!/usr/bin/env python
"""Runs the HBM-bound BatchNorm / activation passes on the generator's widest layer (G.bn4: 573,440 rows x 64 channels,
73 MB per bf16 tensor) a few times — the short command `ncu --set full` is pointed at (profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocogan_chainer_b200 import kernels as K  # noqa: E402

M, C = 35 * 16 * 32 * 32, 64
y = torch.randn((M, C), device="cuda").bfloat16()
g = torch.randn((M, C), device="cuda").bfloat16()
out = torch.empty_like(y)
vec = lambda v: torch.full((C,), v, device="cuda")
mean, invstd, scale, shift, gamma = vec(0.1), vec(0.9), vec(0.9), vec(-0.09), vec(1.0)
dgam, dbet, am, av = (torch.zeros(C, device="cuda") for _ in range(4))
for _ in range(2):
    K.bn_stats(y, M, C, gamma, shift, 2e-5, 0.9, mean, invstd, scale, shift, am, av)
    K.affine_act_noise(y, M, C, 32 * 32, scale, shift, K.ACT_RELU, 0.2, 0.0, None, None, None, 0, out)
    K.act_bn_bwd_reduce(g, y, M, C, mean, invstd, scale, shift, K.ACT_RELU, 0.2, dgam, dbet, None, None)
    K.act_bn_bwd_apply(g, y, M, C, mean, invstd, gamma, scale, shift, K.ACT_RELU, 0.2, 0, dgam, dbet, out)
torch.cuda.synchronize()
print("ok")
