import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om
from oracle import goku as og
from test_mlp_gpu import _saturated_net
dev = "cuda:0"
rng = np.random.Generator(np.random.PCG64(21))
dims = [16, 200, 200, 16]
p = _saturated_net(rng, dims)
B = 128
z0 = (0.3 * rng.standard_normal((B, 16))).astype(np.float32)
for T, kw in ((3, dict(adaptive=False, dt=0.07)), (12, dict(adaptive=False, dt=0.07)), (12, dict(adaptive=False, dt=0.2)), (12, dict(norm_mode=0))):
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(4).standard_normal((T, B, 16)).astype(np.float32)
    o = ldeq.default_opts(mlp_math=1, **kw)
    res = {}
    for name, off in (("tc", False), ("exact", True)):
        if off: os.environ["LDEQ_MLP_TC_BWD_OFF"] = "1"
        else: os.environ.pop("LDEQ_MLP_TC_BWD_OFF", None)
        z = torch.from_numpy(z0).to(dev).requires_grad_(True); pp = torch.from_numpy(p).to(dev).requires_grad_(True)
        st = []
        tr = ldeq.mlp_solve(z, pp, dims, t, o, st); tr.backward(torch.from_numpy(d).to(dev)); torch.cuda.synchronize()
        res[name] = (z.grad.cpu().numpy(), pp.grad.cpu().numpy(), int(st[0].naccept.max()))
    gz, gp, na = res["tc"]; ez, ep, _ = res["exact"]
    rows = np.abs(gz - ez).max(1) / np.abs(ez).max()
    print(T, kw, "naccept", na, "rows>1e-5:", int((rows > 1e-5).sum()), "dz0 max", rows.max(), "median row", np.median(rows), "dparams", np.abs(gp - ep).max() / np.abs(ep).max(), flush=True)
