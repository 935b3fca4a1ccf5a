"""A handful of launches of the GOKU kernels at the C4 size for ncu (forward with tape, adjoint, forward-dual pullback)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import pendulum_inputs
dev = torch.device("cuda:0")
B, T = 1 << 20, 200
z0, th = pendulum_inputs(B)
z = torch.from_numpy(z0).to(dev); p = torch.from_numpy(th).to(dev)
t = 0.05 * np.arange(T)
d = torch.randn(T, B, 2, device=dev)
# launches: (fwd, bwd) x 3 warm, then fwd, bwd, [fwd], fwdsens_p, fwdsens_u
for sense in (0, 0, 0, 0, 1):
    o = ldeq.default_opts(sensealg=sense)
    traj, st, tape = ldeq.goku_solve_raw(z, p, t, 0, o, want_tape=True, want_stats=False)
    tape.p_dim = 1
    ldeq.goku_bwd_raw(tape, d)
    tape.free()
torch.cuda.synchronize()
