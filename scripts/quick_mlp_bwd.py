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
for B in [int(a) for a in sys.argv[1:]] or [256, 2048]:
    z=(0.5*torch.randn(B,16,device="cuda")); d=torch.randn(T,B,16,device="cuda")
    for name,kw in [("exact fwd", dict(norm_mode=1)), ("tc fwd", dict(norm_mode=1, mlp_math=1))]:
        o=ldeq.default_opts(**kw)
        def run():
            tr,st,tape=ldeq.mlp_solve_raw(z,p,dims,t,o,want_tape=True)
            g=ldeq.mlp_bwd_raw(tape,d); tape.free(); return st
        for _ in range(2): run()
        torch.cuda.synchronize()
        e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
        n=5; tf=tb=0
        for _ in range(n):
            e[0].record(); tr,st,tape=ldeq.mlp_solve_raw(z,p,dims,t,o,want_tape=True); e[1].record(); g=ldeq.mlp_bwd_raw(tape,d); e[2].record(); torch.cuda.synchronize()
            tf+=e[0].elapsed_time(e[1]); tb+=e[1].elapsed_time(e[2]); tape.free()
        print(f"B={B} {name}: fwd {tf/n:.3f} ms, bwd (exact adjoint kernel) {tb/n:.3f} ms, naccept {st.naccept.float().mean().item():.1f}")
