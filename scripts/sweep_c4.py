"""C4 sweep (BASELINE.json configs[3]): N = 2^10 .. 2^20 trajectories x 200 save points, forward and forward + adjoint,
fp32 and fp64 state, device-resident, CUDA events, pipelined launches.  One JSON line per point."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from conftest import pendulum_inputs
dev = torch.device("cuda:0"); T = 200; t = 0.05 * np.arange(T)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for dtype in (torch.float32, torch.float64):
    es = 4 if dtype == torch.float32 else 8
    for lg in (10, 13, 16, 20):
        B = 1 << lg
        z0, th = pendulum_inputs(B)
        z = torch.from_numpy(z0).to(dev, dtype); p = torch.from_numpy(th).to(dev, dtype)
        d = torch.randn(T, B, 2, device=dev, dtype=dtype)
        opts = ldeq.default_opts()                                             # library default: the reference's dual-number pullback
        opts_da = ldeq.default_opts(sensealg=ldeq.SENSE_DISCRETE_ADJOINT)     # the explicit opt-in
        def fwd():
            return ldeq.goku_solve_raw(z, p, t, 0, opts, want_tape=False, want_stats=False)
        def fwdbwd(o):
            tr, st, tape = ldeq.goku_solve_raw(z, p, t, 0, o, want_tape=True, want_stats=False); tape.p_dim = 1
            g = ldeq.goku_bwd_raw(tape, d); tape.free(); return g
        res = {}
        for name, fn in (("forward", fwd), ("forward+adjoint", lambda: fwdbwd(opts_da)), ("forward+forward_dual", lambda: fwdbwd(opts))):
            n = 200 if lg <= 13 else 50 if lg <= 16 else 20
            for _ in range(5): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            alg = B * (3 * es + 2 * es * T) * (1 if name == "forward" else 2)
            res[name] = {"ms": round(ms, 4), "traj_steps_per_s": B * (T - 1) / (ms * 1e-3), "alg_GBps": alg / (ms * 1e-3) / 1e9,
                         "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak}
        print(json.dumps({"N": B, "T": T, "dtype": "f32" if es == 4 else "f64", **res}))
