import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_step_gpu import build_pair, relerr
from oracle import mocogan_ref as ref
from mocogan_chainer_b200 import random as mrandom
mode = sys.argv[1]; nf = int(sys.argv[2]); N = int(sys.argv[3])
model, (G, Di, Dv), (oG, oI, oV), up, oup = build_pair("mug_normal", nf, mode)
C = 3
x_real = np.random.default_rng(1234).uniform(-1, 1, size=(N, C, 16, 64, 64)).astype(np.float32)
t_real = np.random.default_rng(5).integers(0, 6, size=N)
r = ref.draw_step_randoms(np.random.default_rng(100), np.random.default_rng(200), oG, oI, oV, N, x_real.shape, t=7, dtype=np.float32)
trace = {}
ol = oup.update_core(x_real.astype(np.float64), t_real, r, trace=trace)
mrandom.set_source(mrandom.InjectedRandom(r))
up.step_on_device(torch.from_numpy(x_real).cuda(), torch.from_numpy(t_real).int().cuda())
torch.cuda.synchronize()
print(mode, nf, N, "TC off" if os.environ.get("MCG_DISABLE_TC") else "TC on")
print({k: float(v) for k, v in up.losses.items()}, ol)
for mine, key in ((Di, "grads_di"), (Dv, "grads_dv"), (G, "grads_g")):
    for path, p in mine.namedparams():
        k = path.lstrip("/")
        e = relerr(p.grad.float().cpu().numpy(), trace[key][k])
        print("  %-20s %-12s %.3e" % (mine.name, k, e))
