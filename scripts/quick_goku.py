"""Scratch timing of the GOKU kernels (device-resident), used while iterating; bench.py is the contract."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import pendulum_inputs

dev = torch.device("cuda:0")
cfgs = [(1 << 20, 200), (1 << 16, 200), (1 << 20, 50)]
if len(sys.argv) > 1:
    cfgs = cfgs[:int(sys.argv[1])]
modes = (True, False) if len(sys.argv) <= 2 else (True,)
for B, T in cfgs:
    z0, th = pendulum_inputs(B)
    z = torch.from_numpy(z0).to(dev); p = torch.from_numpy(th).to(dev)
    t = 0.05 * np.arange(T)
    d = torch.randn(T, B, 2, device=dev)
    for adaptive in modes:
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
        # compute-only bound: same kernel with the output stores disabled (traj_out = NULL)
        import ctypes as C
        h = ldeq.handle(0); na_t = torch.empty(B, dtype=torch.int32, device=dev)
        tg = np.ascontiguousarray(t)
        def nostore():
            h.check(h._lib.ldeq_solve_fwd(h.ptr, h.rhs_builtin(0), 0, C.c_void_p(z.data_ptr()), C.c_void_p(p.data_ptr()),
                    tg.ctypes.data_as(C.c_void_p), B, T, C.byref(opts), None, None, C.c_void_p(na_t.data_ptr()), None, None,
                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        nostore(); torch.cuda.synchronize(); e[2].record()
        for it in range(n): nostore()
        e[3].record(); torch.cuda.synchronize()
        print(f"   no-store fwd {e[2].elapsed_time(e[3]) / n:.3f} ms")
        steps = B * (T - 1)
        na = st.naccept.float().mean().item()
        print(f"B={B} T={T} adaptive={adaptive}: fwd {f_ms:.3f} ms ({steps/f_ms/1e6:.1f} G traj-steps/s, {steps*8.06/f_ms/1e6:.0f} GB/s alg) | "
              f"fwd+tape {tf/n:.3f} ms, bwd {tb/n:.3f} ms, fwd+bwd {steps/((tf+tb)/n)/1e6:.1f} G ts/s | naccept mean {na:.1f} max {st.naccept.max().item()}")
