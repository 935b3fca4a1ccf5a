"""One exact forward + backward of the C2 LatentODE workload (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om
rng = np.random.Generator(np.random.PCG64(1))
dims = [16, 200, 200, 16]
layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]), np.zeros(dims[i + 1], np.float32)) for i in range(3)]
p = torch.from_numpy(om.pack_params(layers).astype(np.float32)).cuda()
T = 50; t = 0.05 * np.arange(T)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
z = 0.5 * torch.randn(B, 16, device="cuda"); d = torch.randn(T, B, 16, device="cuda")
o = ldeq.default_opts(norm_mode=1)
for _ in range(2):
    tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True)
    g = ldeq.mlp_bwd_raw(tape, d); tape.free()
torch.cuda.synchronize()
