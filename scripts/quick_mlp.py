"""Scratch timing of the LatentODE kernels: exact CUDA-core path vs tcgen05 path (device-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om
rng = np.random.Generator(np.random.PCG64(1))
dims=[16,200,200,16]
layers=[(om.glorot_uniform(rng,dims[i+1],dims[i]), np.zeros(dims[i+1],np.float32)) for i in range(3)]
p=torch.from_numpy(om.pack_params(layers).astype(np.float32)).cuda()
T=50; t=0.05*np.arange(T)
only = sys.argv[1] if len(sys.argv) > 1 else None
for B in ([256, 2048, 18944, 65536] if not only else [int(only)]):
    z=(0.5*torch.randn(B,16,device="cuda"))
    for name,kw in [("exact per-traj", dict(norm_mode=1)), ("tc per-traj", dict(norm_mode=1, mlp_math=1)), ("exact global", dict(norm_mode=0)), ("tc global", dict(norm_mode=0, mlp_math=1)), ("tc fixed", dict(adaptive=False, dt=0.05, mlp_math=1))]:
        if "global" in name and B > 18944: continue
        if "exact global" in name and B > 4736: continue
        if name.startswith("exact") and B > 20000 and only is None: continue
        o=ldeq.default_opts(**kw)
        try:
            for _ in range(2): tr,st,_=ldeq.mlp_solve_raw(z,p,dims,t,o)
        except ldeq.LdeqError as e:
            print(B, name, "->", e); continue
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        n=5; e0.record()
        for _ in range(n): tr,st,_=ldeq.mlp_solve_raw(z,p,dims,t,o)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/n
        na=st.naccept.float().mean().item(); nr=st.nreject.float().mean().item()
        rhs = B*(6*(st.naccept+st.nreject).float().mean().item()+2)
        print(f"B={B:6d} {name:15s}: {ms:8.3f} ms | {B*(T-1)/ms/1e3:8.1f} M traj-steps/s | naccept {na:.1f} nreject {nr:.1f} | {rhs*92800/ms/1e9:7.2f} TFLOP/s (algorithmic MLP flops, 92 800 per RHS)")
