import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om, goku as og
rng = np.random.Generator(np.random.PCG64(1))
dims=[16,200,200,16]
layers=[(om.glorot_uniform(rng,dims[i+1],dims[i]), (0.1*rng.standard_normal(dims[i+1])).astype(np.float32)) for i in range(3)]
p=om.pack_params(layers).astype(np.float32)
B=int(sys.argv[1]) if len(sys.argv)>1 else 256
T=50
z0=(0.5*rng.standard_normal((B,16))).astype(np.float32); t=0.05*np.arange(T)
z=torch.from_numpy(z0).cuda(); pp=torch.from_numpy(p).cuda()
for name,kw in [("fixed dt=.05", dict(adaptive=False, dt=0.05)), ("adaptive per-traj", dict(norm_mode=ldeq.NORM_PER_TRAJ)), ("adaptive global", dict(norm_mode=ldeq.NORM_GLOBAL))]:
    ex,st_e,_=ldeq.mlp_solve_raw(z,pp,dims,t,ldeq.default_opts(**kw))
    tc,st_t,_=ldeq.mlp_solve_raw(z,pp,dims,t,ldeq.default_opts(mlp_math=ldeq.MLP_MATH_BF16X3, **kw))
    torch.cuda.synchronize()
    ex=ex.cpu().numpy(); tc=tc.cpu().numpy()
    print(name, "| tc vs exact max err", np.abs(tc-ex).max(), "rel", np.abs(tc-ex).max()/np.abs(ex).max(), "| nan", np.isnan(tc).sum(),
          "| naccept eq", (st_e.naccept==st_t.naccept).float().mean().item(), "ret", st_t.retcode.max().item())
    if B <= 256 and "fixed" in name:
        o,_,_,_=om.solve(z0,p,dims,t,og.Opts(adaptive=False,dt=0.05))
        print("   vs oracle: tc", np.abs(tc-o).max()/np.abs(o).max(), "exact", np.abs(ex-o).max()/np.abs(o).max())
