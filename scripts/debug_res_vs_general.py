import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om
dev = "cuda:0"
dims = [16, 200, 200, 16]
T = 50; t = 0.05 * np.arange(T)
def run(z, p, o):
    tr, st, _ = ldeq.mlp_solve_raw(z, p, dims, t, o)
    return tr.cpu().numpy(), st.naccept.cpu().numpy(), st.nreject.cpu().numpy()
for seed in range(6):
    rng = np.random.Generator(np.random.PCG64(seed))
    layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]), np.zeros(dims[i + 1], np.float32)) for i in range(3)]
    p = torch.from_numpy(om.pack_params(layers).astype(np.float32)).to(dev)
    z = torch.from_numpy((0.5 * rng.standard_normal((256, 16))).astype(np.float32)).to(dev)
    for name, kw in (("global", dict(norm_mode=0)), ("global rtol1e-6", dict(norm_mode=0, reltol=1e-6, abstol=1e-8)), ("per-traj", dict(norm_mode=1)), ("fixed", dict(adaptive=False, dt=0.05))):
        o = ldeq.default_opts(**kw)
        os.environ.pop("LDEQ_MLP_NO_RESIDENT", None)
        a = run(z, p, o)
        os.environ["LDEQ_MLP_NO_RESIDENT"] = "1"
        b = run(z, p, o)
        os.environ.pop("LDEQ_MLP_NO_RESIDENT", None)
        print(f"seed {seed} {name:16s} naccept res {a[1].mean():.2f} gen {b[1].mean():.2f} differ {int((a[1] != b[1]).sum())}  nreject {a[2].mean():.2f}/{b[2].mean():.2f}  max rel diff {np.abs(a[0]-b[0]).max()/np.abs(b[0]).max():.2e}")
