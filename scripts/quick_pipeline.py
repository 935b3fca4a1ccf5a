"""Pipelined fwd(+tape)+bwd throughput and host-side cost per C-ABI call (scratch; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from conftest import pendulum_inputs
dev = torch.device("cuda:0")
B, T = 1 << 20, 200
z0, th = pendulum_inputs(B)
z = torch.from_numpy(z0).to(dev); p = torch.from_numpy(th).to(dev)
t = 0.05 * np.arange(T); d = torch.randn(T, B, 2, device=dev)
opts = ldeq.default_opts()
def step():
    t0 = time.perf_counter()
    traj, st, tape = ldeq.goku_solve_raw(z, p, t, 0, opts, want_tape=True, want_stats=False)
    tape.p_dim = 1
    t1 = time.perf_counter()
    g = ldeq.goku_bwd_raw(tape, d)
    t2 = time.perf_counter()
    tape.free()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
for _ in range(3): step()
torch.cuda.synchronize()
K = 20
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
hf = hb = hfree = 0.0
w0 = time.perf_counter(); e0.record()
for _ in range(K):
    a, b_, c = step(); hf += a; hb += b_; hfree += c
e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
ms = e0.elapsed_time(e1) / K
print(f"pipelined fwd+tape+bwd: {ms:.3f} ms/step device, wall {(w1-w0)/K*1e3:.3f} ms/step; host per call: fwd {hf/K*1e3:.3f} ms, bwd {hb/K*1e3:.3f} ms, free {hfree/K*1e3:.3f} ms")
print(f"  => {B*(T-1)/ms/1e6:.1f} G traj-steps/s fwd+adjoint, {B*(T-1)*16.1/ms/1e6:.0f} GB/s algorithmic")
