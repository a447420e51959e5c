#!/usr/bin/env python
"""Runs the three tcgen05 launches of one BASELINE config-2 layer (default Dv.dc2) a few times — the short command
`ncu --set full` is pointed at (profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocogan_chainer_b200 import kernels as K  # noqa: E402

LAYERS = {"Dv.dc2": (35, 64, 128, (13, 32, 32), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
          "Dv.dc3": (35, 128, 256, (10, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
          "Dv.dc4": (35, 256, 512, (7, 8, 8), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
          "Di.dc2": (35, 64, 128, (1, 32, 32), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
          "G.dc3": (560, 128, 256, (1, 16, 16), (1, 4, 4), (1, 2, 2), (0, 1, 1)),
          "G.dc4": (560, 64, 128, (1, 32, 32), (1, 4, 4), (1, 2, 2), (0, 1, 1))}
name = sys.argv[1] if len(sys.argv) > 1 else "Dv.dc2"
N, Cin, Cout, in_sp, k, s, p = LAYERS[name]
g = K.make_geom(N, Cin, Cout, in_sp, k, s, p)
x = torch.randn((N,) + in_sp + (Cin,), device="cuda").bfloat16()
w = (torch.randn((Cout,) + k + (Cin,), device="cuda") * 0.05).bfloat16()
gy = torch.randn((N, g.To, g.Ho, g.Wo, Cout), device="cuda").bfloat16()
y, dx, dw = torch.empty_like(gy), torch.empty_like(x), torch.zeros(w.shape, device="cuda")
for _ in range(3):
    K.conv_fprop(g, x, w, None, y, K.IMPL_TC)
    K.conv_dgrad(g, gy, w, None, dx, K.IMPL_TC)
    K.conv_wgrad(g, x, gy, dw, K.IMPL_TC)
torch.cuda.synchronize()
assert K.tc_error_flag() == 0
print("ok", name)
