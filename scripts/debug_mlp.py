import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om, goku as og
rng = np.random.Generator(np.random.PCG64(1))
dims=[16,200,200,16]
layers=[(om.glorot_uniform(rng,dims[i+1],dims[i]).astype(np.float64), np.zeros(dims[i+1])) for i in range(3)]
p=om.pack_params(layers)
B,T=int(sys.argv[1]) if len(sys.argv)>1 else 8,50
z0=0.5*rng.standard_normal((B,16)); t=0.05*np.arange(T)
for cp in (1,0):
    opts=ldeq.default_opts(norm_mode=ldeq.NORM_PER_TRAJ, controller_pow=cp)
    z=torch.from_numpy(z0).cuda(); pp=torch.from_numpy(p).cuda()
    tr,st,tape=ldeq.mlp_solve_raw(z,pp,dims,t,opts,want_tape=True); torch.cuda.synchronize()
    na=st.naccept.cpu().numpy(); nr=st.nreject.cpu().numpy()
    otr,ona,onr,_=om.solve(z0[:8],p,dims,t,og.Opts(controller_pow=cp),norm_mode="per_traj"); na=na[:8]; nr=nr[:8]; tr=tr[:,:8]
    print("pow",cp,"kernel na",na,"nr",nr); print("      oracle na",ona,"nr",onr, "traj err", np.abs(tr.cpu().numpy()-otr).max())
    # step sizes of trajectory 0 from the kernel tape vs oracle
    import ctypes
    o2, n2, r2, tp = om.solve(z0[:1],p,dims,t,og.Opts(controller_pow=cp),record=True)
    print("      oracle dt[0]:", np.array(tp.dt)[:6])
