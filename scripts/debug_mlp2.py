import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import mlp as om, goku as og
from test_mlp_gpu import _net, _solve
dtype=sys.argv[1]
dims, p, rng = _net(dtype=dtype)
B, T = 256, 50
z0 = (0.5 * rng.standard_normal((B, 16))).astype(dtype)
t = 0.05 * np.arange(T)
tr, ret, na, nr = _solve(ldeq, z0, p, dims, t, norm_mode=ldeq.NORM_PER_TRAJ)
otr, ona, onr, _ = om.solve(z0[:16], p, dims, t, norm_mode="per_traj")
print("kernel na", na[:16], "nr", nr[:16]); print("oracle na", ona, "nr", onr)
print("traj err per traj", np.array2string(np.abs(tr[:, :16]-otr).max(axis=(0,2)), precision=1, max_line_width=200))
tr8, ret8, na8, nr8 = _solve(ldeq, z0[:16], p, dims, t, norm_mode=ldeq.NORM_PER_TRAJ)
print("kernel(B=16) na", na8, "nr", nr8)
if dtype == "float64":
    e = np.abs(tr[:, 8] - otr[:, 8]).max(axis=1)
    print("traj 8 err vs k:", np.array2string(e, precision=1, max_line_width=250))
    o1, n1, r1, tp = om.solve(z0[8:9], p, dims, t, record=True)
    print("oracle traj 8 t:", np.array2string(np.array(tp.t), precision=4, max_line_width=250))
    print("oracle traj 8 dt:", np.array2string(np.array(tp.dt), precision=4, max_line_width=250))
    for cp in (1,):
        tr2, ret2, na2, nr2 = _solve(ldeq, z0, p, dims, t, norm_mode=ldeq.NORM_PER_TRAJ, controller_pow=cp)
        o2, n2, r2, _ = om.solve(z0[8:9], p, dims, t, og.Opts(controller_pow=cp))
        print("exact pow: traj 8 err", np.abs(tr2[:, 8] - o2[:, 0]).max(), na2[8], n2)
