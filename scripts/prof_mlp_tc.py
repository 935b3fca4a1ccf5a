"""A few launches of the LatentODE tensor-core forward + reverse pass at B = 18 944 for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
import bench
dev = torch.device("cuda:0")
B, T, dims, p_np, z_np, d_np, t = bench._latentode_inputs("mlp")
p, z, d = (torch.from_numpy(a).to(dev) for a in (p_np, z_np, d_np))
o = ldeq.default_opts(norm_mode=int(os.environ.get("NORM", "0")), mlp_math=1)
for _ in range(3):
    tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True)
    g = ldeq.mlp_bwd_raw(tape, d)
    tape.free()
torch.cuda.synchronize()
