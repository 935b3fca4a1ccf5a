"""Timing of the LatentODE tensor-core path: forward and forward + reverse pass at B = 18 944 (128 trajectories per SM) and C2."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
import bench
dev = torch.device("cuda:0")
res = {}
for wl in ("mlp", "c2"):
    B, T, dims, p_np, z_np, d_np, t = bench._latentode_inputs(wl)
    p, z, d = (torch.from_numpy(a).to(dev) for a in (p_np, z_np, d_np))
    for name, kw in (("tc_per_traj", dict(norm_mode=1, mlp_math=1)), ("tc_global", dict(norm_mode=0, mlp_math=1)), ("exact_per_traj", dict(norm_mode=1))):
        if wl == "mlp" and name == "exact_per_traj" and os.environ.get("SKIP_EXACT"):
            continue
        o = ldeq.default_opts(**kw)
        def fwd():
            tr, st, _ = ldeq.mlp_solve_raw(z, p, dims, t, o)
            return st
        def both():
            tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True)
            g = ldeq.mlp_bwd_raw(tape, d)
            tape.free()
            return g
        try:
            for _ in range(2):
                st = fwd(); both()
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            n = 5
            e[0].record()
            for _ in range(n): fwd()
            e[1].record()
            for _ in range(n): g = both()
            e[2].record()
            torch.cuda.synchronize()
            res[f"{wl}_{name}"] = {"B": B, "fwd_ms": e[0].elapsed_time(e[1]) / n, "fwd_bwd_ms": e[1].elapsed_time(e[2]) / n,
                                   "naccept": float(st.naccept.float().mean()), "finite": bool(torch.isfinite(g[0]).all() and torch.isfinite(g[1]).all())}
        except ldeq.LdeqError as ex:
            res[f"{wl}_{name}"] = {"error": str(ex)[:200]}
print(json.dumps(res))
