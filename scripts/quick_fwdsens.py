"""Forward-dual sensitivity mode (the reference's ForwardDiffSensitivity algorithm): parity numbers and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import goku as og
from conftest import pendulum_inputs
dev = "cuda:0"
for dtype in ("float64", "float32"):
    B, T = 4096, 50
    z0, th = pendulum_inputs(B, dtype=dtype); t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(dtype)
    rz, rp = og.grad(0, z0, th, t, d, norm_partials=True)
    for name, kw in (("forward_dual", dict(sensealg=1)), ("discrete_adjoint", dict())):
        z = torch.from_numpy(z0).to(dev).requires_grad_(True); p = torch.from_numpy(th).to(dev).requires_grad_(True)
        ldeq.goku_solve(z, p, t, 0, ldeq.default_opts(**kw)).backward(torch.from_numpy(d).to(dev))
        ez = np.abs(z.grad.cpu().numpy() - rz).max(1) / np.abs(rz).max(); ep = np.abs(p.grad.cpu().numpy() - rp).max(1) / np.abs(rp).max()
        print(f"{dtype} {name:17s} vs oracle ForwardDiff semantics (default tol): dz0 max {ez.max():.2e} q95 {np.quantile(ez,.95):.2e} | dtheta max {ep.max():.2e} q95 {np.quantile(ep,.95):.2e}")
B, T = 1 << 20, 200
z0, th = pendulum_inputs(B); t = 0.05 * np.arange(T)
z = torch.from_numpy(z0).to(dev); p = torch.from_numpy(th).to(dev); d = torch.randn(T, B, 2, device=dev)
for name, kw in (("forward_dual", dict(sensealg=1)), ("discrete_adjoint", dict())):
    o = ldeq.default_opts(**kw)
    def step():
        tr, st, tape = ldeq.goku_solve_raw(z, p, t, 0, o, want_tape=True)
        g = ldeq.goku_bwd_raw(tape, d); tape.free()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"2^20 x 200 forward + {name}: {ms:.3f} ms  ({B*(T-1)/ms/1e6:.1f} G traj-steps/s)")
