"""Host-side time of each call in an unsynchronised forward(tape) -> backward -> free loop at the C2 size."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
import bench
dev = torch.device("cuda:0")
B, T, dims, p_np, z_np, d_np, t = bench._latentode_inputs(os.environ.get("WL", "c2"))
p, z, d = (torch.from_numpy(a).to(dev) for a in (p_np, z_np, d_np))
out = {}
for name, kw in (("tc_global", dict(norm_mode=0, mlp_math=1)), ("tc_per_traj", dict(norm_mode=1, mlp_math=1)), ("exact_global", dict(norm_mode=0))):
    o = ldeq.default_opts(**kw)
    for it in range(3):
        tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True); g = ldeq.mlp_bwd_raw(tape, d); tape.free()
    torch.cuda.synchronize()
    rows = []
    t00 = time.perf_counter()
    for it in range(12):
        t0 = time.perf_counter()
        tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True)
        t1 = time.perf_counter()
        g = ldeq.mlp_bwd_raw(tape, d)
        t2 = time.perf_counter()
        tape.free()
        t3 = time.perf_counter()
        rows.append([round((t1 - t0) * 1e3, 3), round((t2 - t1) * 1e3, 3), round((t3 - t2) * 1e3, 3)])
    torch.cuda.synchronize()
    out[name] = {"total_ms_per_iter": (time.perf_counter() - t00) * 1e3 / 12, "fwd_bwd_free_host_ms": rows[-5:]}
print(json.dumps(out))
