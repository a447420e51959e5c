import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_step_gpu import build_pair, relerr
from oracle import mocogan_ref as ref
from mocogan_chainer_b200 import random as mrandom
model, (G, Di, Dv), (oG, oI, oV), up, oup = build_pair("mnist_normal", 16, "fp32")
N, C = 3, 1
x_real = np.random.default_rng(1234).uniform(-1, 1, size=(N, C, 16, 64, 64)).astype(np.float32)
r = ref.draw_step_randoms(np.random.default_rng(100), np.random.default_rng(200), oG, oI, oV, N, x_real.shape, t=7, dtype=np.float32)
trace = {}
ol = oup.update_core(x_real.astype(np.float64), None, r, trace=trace)
mrandom.set_source(mrandom.InjectedRandom(r))
up.step_on_device(torch.from_numpy(x_real).cuda(), None)
torch.cuda.synchronize()
for mine, key in ((Di, "grads_di"), (Dv, "grads_dv")):
    for path, p in mine.namedparams():
        k = path.lstrip("/")
        g = p.grad.float().cpu().numpy(); gr = trace[key][k]
        e = relerr(g, gr)
        if e > 1e-5: print("  %-20s %-12s %.3e  argmax %s" % (mine.name, k, e, np.unravel_index(np.abs(g-gr).argmax(), g.shape)))
print("done")
