import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om
from oracle import goku as og
dev = "cuda:0"
rng = np.random.Generator(np.random.PCG64(1))
dims = [16, 200, 200, 16]
layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]), (0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32)) for i in range(3)]
p = om.pack_params(layers).astype(np.float32)
def run(B, T, dt, env=None):
    z0 = (0.5 * np.random.default_rng(3).standard_normal((B, 16))).astype(np.float32)
    t = dt * np.arange(T)
    d = np.random.default_rng(4).standard_normal((T, B, 16)).astype(np.float32)
    o = ldeq.default_opts(adaptive=False, dt=dt, mlp_math=1)
    out = {}
    for name, off in (("tc", False), ("exact", True)):
        if off: os.environ["LDEQ_MLP_TC_BWD_OFF"] = "1"
        else: os.environ.pop("LDEQ_MLP_TC_BWD_OFF", None)
        z = torch.from_numpy(z0).to(dev).requires_grad_(True); pp = torch.from_numpy(p).to(dev).requires_grad_(True)
        tr = ldeq.mlp_solve(z, pp, dims, t, o); tr.backward(torch.from_numpy(d).to(dev)); torch.cuda.synchronize()
        out[name] = (z.grad.cpu().numpy(), pp.grad.cpu().numpy())
    _, _, _, tape = om.solve(z0, p, dims, t, og.Opts(adaptive=False, dt=dt), record=True)
    oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d)
    out["oracle"] = (oz, op)
    return out
offs = np.cumsum([0, 200 * 16, 200, 200 * 200, 200, 16 * 200, 16])
names = ["W1", "b1", "W2", "b2", "W3", "b3"]
for B, T, dt in ((128, 2, 0.05), (128, 3, 0.05), (130, 20, 0.05)):
    r = run(B, T, dt)
    for a in ("tc", "exact"):
        gz, gp = r[a]; oz, op = r["oracle"]
        line = {"dz0": float(np.abs(gz - oz).max() / np.abs(oz).max())}
        for i, nme in enumerate(names):
            sl = slice(offs[i], offs[i + 1])
            line[nme] = float(np.abs(gp[sl] - op[sl]).max() / np.abs(op[sl]).max())
        print(B, T, a, json.dumps({k: round(v, 7) for k, v in line.items()}), flush=True)
