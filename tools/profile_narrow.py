#!/usr/bin/env python
"""Runs fprop / dgrad / wgrad of one 3-channel image layer of BASELINE config 2 (default Dv.dc1) a few times — the short
command the ncu launch-list pass is pointed at."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocogan_chainer_b200 import kernels as K  # noqa: E402

LAYERS = {"Dv.dc1": (35, 3, 64, (16, 64, 64), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
          "Di.dc1": (35, 3, 64, (1, 64, 64), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
          "G.dc5": (560, 3, 64, (1, 64, 64), (1, 4, 4), (1, 2, 2), (0, 1, 1))}
name = sys.argv[1] if len(sys.argv) > 1 else "Dv.dc1"
N, Cin, Cout, in_sp, k, s, p = LAYERS[name]
g = K.make_geom(N, Cin, Cout, in_sp, k, s, p)
x = torch.randn((N,) + in_sp + (Cin,), device="cuda").bfloat16()
w = (torch.randn((Cout,) + k + (Cin,), device="cuda") * 0.05).bfloat16()
gy = torch.randn((N, g.To, g.Ho, g.Wo, Cout), device="cuda").bfloat16()
y, dx, dw = torch.empty_like(gy), torch.empty_like(x), torch.zeros(w.shape, device="cuda")
for _ in range(3):
    ws = K.conv_fprop(g, x, w, None, y, K.IMPL_TC)
    K.conv_dgrad(g, gy, w, None, dx, K.IMPL_TC)
    K.conv_wgrad(g, x, gy, dw, K.IMPL_TC, ws=ws, cols_valid=True)
torch.cuda.synchronize()
assert K.tc_error_flag() == 0
print("ok", name)
