import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from conftest import pendulum_inputs
from oracle import goku as og
dev = torch.device("cuda:0")
def grads(rhs, z0, th, t, d, **kw):
    z = torch.from_numpy(z0).to(dev).requires_grad_(True); p = torch.from_numpy(th).to(dev).requires_grad_(True)
    st = []
    traj = ldeq.goku_solve(z, p, t, rhs, ldeq.default_opts(**kw), st)
    traj.backward(torch.from_numpy(d).to(dev)); torch.cuda.synchronize()
    return z.grad.cpu().numpy(), p.grad.cpu().numpy(), st[0].naccept.cpu().numpy(), traj.detach().cpu().numpy()
B, T = 512, 50
t = 0.05 * np.arange(T)
for dtype in ("float32", "float64"):
    z0, th = pendulum_inputs(B, dtype=dtype)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(dtype)
    for name, kw, okw in [("adaptive", {}, {}), ("fixed dt=0.07", dict(adaptive=False, dt=0.07), dict(adaptive=False, dt=0.07)),
                          ("fixed dt=0.05", dict(adaptive=False, dt=0.05), dict(adaptive=False, dt=0.05))]:
        gz, gp, na, tr = grads(0, z0, th, t, d, **kw)
        oz, op = og.grad(0, z0, th, t, d, og.Opts(**okw), norm_partials=False)
        otr, _, ona, _ = og.solve(0, z0, th, t, og.Opts(**okw))
        ez = np.abs(gz - oz).max(1) / np.abs(oz).max(); ep = np.abs(gp - op).max(1) / np.abs(op).max()
        same = na == ona
        print(f"{dtype} {name}: traj err {np.abs(tr-otr).max():.2e} | dz0 err q50 {np.quantile(ez,.5):.2e} q98 {np.quantile(ez,.98):.2e} max {ez.max():.2e}"
              f" | dth q50 {np.quantile(ep,.5):.2e} q98 {np.quantile(ep,.98):.2e} max {ep.max():.2e} | same-steps {same.mean():.3f}"
              f" | err on same-step trajs: dz0 {ez[same].max():.2e} dth {ep[same].max():.2e}")
