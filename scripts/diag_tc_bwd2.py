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
B, T, dt = 128, 3, 0.05
z0 = (0.5 * np.random.default_rng(3).standard_normal((B, 16))).astype(np.float32)
t = dt * np.arange(T)
d_full = np.random.default_rng(4).standard_normal((T, B, 16)).astype(np.float32)
for label, keep in (("only d[1]", [1]), ("only d[2]", [2]), ("all", [0, 1, 2])):
    d = np.zeros_like(d_full)
    for k in keep: d[k] = d_full[k]
    o = ldeq.default_opts(adaptive=False, dt=dt, mlp_math=1)
    os.environ.pop("LDEQ_MLP_TC_BWD_OFF", None)
    z = torch.from_numpy(z0).to(dev).requires_grad_(True); pp = torch.from_numpy(p).to(dev).requires_grad_(True)
    tr = ldeq.mlp_solve(z, pp, dims, t, o); tr.backward(torch.from_numpy(d).to(dev)); torch.cuda.synchronize()
    gz = z.grad.cpu().numpy()
    os.environ["LDEQ_MLP_TC_BWD_OFF"] = "1"
    z = torch.from_numpy(z0).to(dev).requires_grad_(True); pp = torch.from_numpy(p).to(dev).requires_grad_(True)
    tr = ldeq.mlp_solve(z, pp, dims, t, o); tr.backward(torch.from_numpy(d).to(dev)); torch.cuda.synchronize()
    ez = z.grad.cpu().numpy()
    err = np.abs(gz - ez).max(1) / np.abs(ez).max()
    print(label, "max", err.max(), "rows > 1e-5:", int((err > 1e-5).sum()), "worst rows", np.argsort(-err)[:6].tolist(), np.sort(err)[-6:][::-1].round(6).tolist(), flush=True)
    w = int(np.argmax(err))
    print("   worst row comps err", (np.abs(gz[w] - ez[w]) / np.abs(ez).max()).round(6).tolist())
