"""Scratch timing of the GOKU kernels (device-resident), used while iterating; bench.py is the contract."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import pendulum_inputs

dev = torch.device("cuda:0")
for B, T in [(1 << 20, 200), (1 << 16, 200), (1 << 20, 50)]:
    z0, th = pendulum_inputs(B)
    z = torch.from_numpy(z0).to(dev); p = torch.from_numpy(th).to(dev)
    t = 0.05 * np.arange(T)
    d = torch.randn(T, B, 2, device=dev)
    for adaptive in (True, False):
        opts = ldeq.default_opts(adaptive=adaptive, dt=0.0 if adaptive else 0.05)
        for it in range(3):
            traj, st, tape = ldeq.goku_solve_raw(z, p, t, 0, opts, want_tape=True)
            tape.p_dim = 1
            ldeq.goku_bwd_raw(tape, d); tape.free()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        n = 5
        e[0].record()
        for it in range(n):
            traj, st, _ = ldeq.goku_solve_raw(z, p, t, 0, opts, want_tape=False, want_stats=False)
        e[1].record()
        tapes = []
        tf = tb = 0.0
        for it in range(n):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
            a.record()
            traj, st, tape = ldeq.goku_solve_raw(z, p, t, 0, opts, want_tape=True)
            tape.p_dim = 1
            b.record()
            g = ldeq.goku_bwd_raw(tape, d)
            c.record()
            torch.cuda.synchronize()
            tf += a.elapsed_time(b); tb += b.elapsed_time(c)
            tape.free()
        torch.cuda.synchronize()
        f_ms = e[0].elapsed_time(e[1]) / n
        steps = B * (T - 1)
        na = st.naccept.float().mean().item()
        print(f"B={B} T={T} adaptive={adaptive}: fwd {f_ms:.3f} ms ({steps/f_ms/1e6:.1f} G traj-steps/s, {steps*8.06/f_ms/1e6:.0f} GB/s alg) | "
              f"fwd+tape {tf/n:.3f} ms, bwd {tb/n:.3f} ms, fwd+bwd {steps/((tf+tb)/n)/1e6:.1f} G ts/s | naccept mean {na:.1f} max {st.naccept.max().item()}")
