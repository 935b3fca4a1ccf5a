"""Where does the time of the tensor-core LatentODE path go at the C2 size (B = 256)?  Events around each call."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
import bench
dev = torch.device("cuda:0")
B, T, dims, p_np, z_np, d_np, t = bench._latentode_inputs(os.environ.get("WL", "c2"))
p, z, d = (torch.from_numpy(a).to(dev) for a in (p_np, z_np, d_np))
out = {}
for name, kw in (("tc_global", dict(norm_mode=0, mlp_math=1)), ("tc_per_traj", dict(norm_mode=1, mlp_math=1)), ("exact_global", dict(norm_mode=0))):
    o = ldeq.default_opts(**kw)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    f, b = [], []
    for it in range(8):
        ev[0].record()
        tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True)
        ev[1].record()
        g = ldeq.mlp_bwd_raw(tape, d)
        ev[2].record()
        torch.cuda.synchronize()
        tape.free()
        f.append(ev[0].elapsed_time(ev[1])); b.append(ev[1].elapsed_time(ev[2]))
    out[name] = {"fwd_tape_ms": f[3:], "bwd_ms": b[3:], "naccept": float(st.naccept.float().mean()), "nreject": float(st.nreject.float().mean())}
print(json.dumps(out))
