#!/usr/bin/env python
"""wgrad-only timing for tile-shape experiments."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocogan_chainer_b200 import kernels as K
LAYERS = {"Dv.dc2": (35, 64, 128, (13, 32, 32), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
          "Dv.dc3": (35, 128, 256, (10, 16, 16), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
          "Dv.dc4": (35, 256, 512, (7, 8, 8), (4, 4, 4), (1, 2, 2), (0, 1, 1)),
          "G.dc3": (560, 128, 256, (1, 16, 16), (1, 4, 4), (1, 2, 2), (0, 1, 1))}
mode = sys.argv[1] if len(sys.argv) > 1 else "wgrad"
for name, (N, Cin, Cout, in_sp, k, s, p) in LAYERS.items():
    g = K.make_geom(N, Cin, Cout, in_sp, k, s, p)
    x = torch.randn((N,) + in_sp + (Cin,), device="cuda").bfloat16()
    w = (torch.randn((Cout,) + k + (Cin,), device="cuda") * 0.05).bfloat16()
    gy = torch.randn((N, g.To, g.Ho, g.Wo, Cout), device="cuda").bfloat16()
    y, dx, dw = torch.empty_like(gy), torch.empty_like(x), torch.zeros(w.shape, device="cuda")
    fn = {"wgrad": lambda: K.conv_wgrad(g, x, gy, dw, K.IMPL_TC), "fprop": lambda: K.conv_fprop(g, x, w, None, y, K.IMPL_TC),
          "dgrad": lambda: K.conv_dgrad(g, gy, w, None, dx, K.IMPL_TC)}[mode]
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * N * g.To * g.Ho * g.Wo * Cout * Cin * k[0] * k[1] * k[2]
    print("%s %s %.4f ms %.0f TF/s" % (name, mode, ms, fl / ms / 1e9), flush=True)
print("err", K.tc_error_flag())
