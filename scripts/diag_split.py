import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from conftest import pendulum_inputs
DEV = "cuda:0"
B, T = 1 << 18, 200
z0n, thn = pendulum_inputs(B)
z0, th = torch.from_numpy(z0n).to(DEV), torch.from_numpy(thn).to(DEV)
t = 0.05 * np.arange(T)
g = torch.Generator(device=DEV); g.manual_seed(7)
d = torch.randn(T, B, 2, device=DEV, generator=g)
def run(z0, th, d, **kw):
    z = z0.clone().requires_grad_(True); p = th.clone().requires_grad_(True)
    tr = ldeq.goku_solve(z, p, t, 0, ldeq.default_opts(**kw)); tr.backward(d)
    return tr.detach(), z.grad, p.grad
for sense in (1, 0):
    tr, gz, gp = run(z0, th, d, sensealg=sense)
    tr_b, gz_b, gp_b = run(z0, th, d, sensealg=sense)
    print("sense", sense, "repeat equal:", torch.equal(tr, tr_b), torch.equal(gz, gz_b), torch.equal(gp, gp_b))
    for lo, hi in ((0, 100_000), (100_000, 200_001), (128 * 1000, 128 * 1500), (32 * 4001, 32 * 6001), (200_001, B)):
        tr2, gz2, gp2 = run(z0[lo:hi], th[lo:hi], d[:, lo:hi].contiguous(), sensealg=sense)
        dz = (gz2 - gz[lo:hi]).abs(); dp = (gp2 - gp[lo:hi]).abs()
        print("  split", lo, hi, "tr eq", torch.equal(tr2, tr[:, lo:hi]), "gz eq", torch.equal(gz2, gz[lo:hi]), "max rel", float(dz.max() / gz.abs().max()),
              "n diff", int((dz.max(1).values > 0).sum()), "gp eq", torch.equal(gp2, gp[lo:hi]), float(dp.max() / gp.abs().max()), int((dp > 0).sum()))
