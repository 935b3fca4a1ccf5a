"""Scratch timing of the GOKU kernels at the C4 size (2^20 x 200), used while iterating; bench.py is the contract.

    python scripts/quick_goku.py [B_log2=20] [T=200]
"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import pendulum_inputs

dev = torch.device("cuda:0")
B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
z0, th = pendulum_inputs(B)
z = torch.from_numpy(z0).to(dev); p = torch.from_numpy(th).to(dev)
t = 0.05 * np.arange(T)
d = torch.randn(T, B, 2, device=dev)
res = {"lib": os.environ.get("LDEQ_LIB", "default"), "B": B, "T": T}


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name, sense in (("adjoint", ldeq.SENSE_DISCRETE_ADJOINT), ("fwddual", ldeq.SENSE_FORWARD_DUAL)):
    o = ldeq.default_opts(sensealg=sense)
    tapes = []

    def fwd():
        traj, st, tape = ldeq.goku_solve_raw(z, p, t, 0, o, want_tape=True, want_stats=False)
        tape.p_dim = 1
        tapes.append(tape)

    def both():
        traj, st, tape = ldeq.goku_solve_raw(z, p, t, 0, o, want_tape=True, want_stats=False)
        tape.p_dim = 1
        ldeq.goku_bwd_raw(tape, d)
        tape.free()
    f = timed(lambda: (fwd(), tapes.pop().free()))
    fb = timed(both)
    res[name] = {"fwd_ms": f, "fwd_bwd_ms": fb, "bwd_ms": fb - f, "G_ts_per_s": B * (T - 1) / fb / 1e6}
traj, st, _ = ldeq.goku_solve_raw(z, p, t, 0, ldeq.default_opts())
res["naccept_mean"] = float(st.naccept.float().mean()); res["nreject_mean"] = float(st.nreject.float().mean())
res["naccept_max"] = int(st.naccept.max())

# host-buffer entry points, one caller thread
hz, hth = torch.from_numpy(z0).pin_memory(), torch.from_numpy(th).pin_memory()
hd = torch.empty(T, B, 2).pin_memory(); hd.copy_(d)
out = torch.empty(T, B, 2).pin_memory(); gz = torch.empty(B, 2).pin_memory(); gth = torch.empty(B, 1).pin_memory()
for name, sense in (("adjoint", ldeq.SENSE_DISCRETE_ADJOINT), ("fwddual", ldeq.SENSE_FORWARD_DUAL)):
    o = ldeq.default_opts(sensealg=sense)

    def sep():
        _, tape = ldeq.goku_solve_host(hz, hth, t, 0, o, want_tape=True, out=out)
        ldeq.goku_bwd_host(tape, hd, gz, gth)
        tape.free()

    def comb():
        ldeq.goku_fwd_bwd_host(hz, hth, t, hd, 0, o, out=out, dz0=gz, dtheta=gth)
    for nm, fn in (("separate", sep), ("combined", comb)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        w = time.perf_counter()
        n = 5
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - w) * 1e3 / n
        res[name]["e2e_" + nm + "_ms"] = ms
        res[name]["e2e_" + nm + "_G_ts_per_s"] = B * (T - 1) / ms / 1e6
print(json.dumps(res))
