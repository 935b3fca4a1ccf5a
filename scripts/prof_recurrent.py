"""One forward + reverse pass of the pattern-extractor kernels at one GPU's C5 share (B = 8192, T = 50), for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import recurrent as orr   # parameter initialisation only (test infrastructure; this is a profiling script)
from latentdiffeq_jl_b200.solve import _PatternExtractor
DEV = "cuda:0"
B, T, F = 8192, 50, 32
rng = np.random.default_rng(0)
x = torch.randn(T, B, F, device=DEV, requires_grad=True)
ps = [torch.from_numpy(orr.init_params(l, F, rng)).to(DEV).requires_grad_(True) for l in (False, True, True)]
w = torch.randn(B, 48, device=DEV)
for _ in range(2):
    z0, th = _PatternExtractor.apply(x, *ps)
    (torch.cat([z0, th], 1) * w).sum().backward()
torch.cuda.synchronize()
