"""Timing of the interpolating-adjoint backward solve at the C2 shape."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
import bench
dev = torch.device("cuda:0")
B, T, dims, p_np, z_np, d_np, t = bench._latentode_inputs(os.environ.get("WL", "c2"))
p, z, d = (torch.from_numpy(a).to(dev) for a in (p_np, z_np, d_np))
o = ldeq.default_opts(norm_mode=0, sensealg=ldeq.SENSE_INTERPOLATING_ADJOINT)
ms = []
for it in range(5):
    tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g = ldeq.mlp_bwd_raw(tape, d); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1)); stats = ldeq.mlp_bwd_stats(tape); tape.free()
print(json.dumps({"tb": os.environ.get("LDEQ_CADJ_TB", "auto"), "bwd_ms": ms[2:], "stats": stats, "us_per_attempt": 1e3 * ms[-1] / (stats[0] + stats[1])}))
