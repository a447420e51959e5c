import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocogan_chainer_b200 import kernels as K
N, Cin, Cout, in_sp, k, s, p = (35, 64, 128, (13, 32, 32), (4, 4, 4), (1, 2, 2), (0, 1, 1))
g = K.make_geom(N, Cin, Cout, in_sp, k, s, p)
w = (torch.randn((Cout,) + k + (Cin,), device="cuda") * 0.05).bfloat16()
gy = torch.randn((N, g.To, g.Ho, g.Wo, Cout), device="cuda").bfloat16()
dx = torch.empty((N,) + in_sp + (Cin,), device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    K.conv_dgrad(g, gy, w, None, dx, K.IMPL_TC)
torch.cuda.synchronize()
print("ok", K.tc_error_flag())
