"""One GOKU training step at one GPU's share of C5 on 8 GPUs (B = 8192), for the ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
import bench
dev = torch.device("cuda:0")
B, T = 8192, 50
z0n, thn = bench.pendulum_inputs(B, seed=1)
t = 0.05 * np.arange(T)
with torch.no_grad():
    ang, _, _ = ldeq.goku_solve_raw(torch.from_numpy(z0n).to(dev), torch.from_numpy(thn).to(dev), t, 0)
    x = bench.synthetic_frames(ang[..., 0], dev)
torch.manual_seed(333)
mt = ldeq.GOKU_basic()
enc, dec = ldeq.default_layers(mt, 784, ldeq.Pendulum(), device=dev)
model = ldeq.LatentDiffEqModel(mt, enc, dec)
flat = ldeq.FlatParams(model)
opt = ldeq.ADAMW(flat, 1e-3, (0.9, 0.999), 1e-3)
for _ in range(3):
    ldeq.train_step(model, flat, opt, x, t, beta=0.5, variational=True, global_batch=B)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("c5_step")
ldeq.train_step(model, flat, opt, x, t, beta=0.5, variational=True, global_batch=B)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
