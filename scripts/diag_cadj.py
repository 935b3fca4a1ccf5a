"""Backward-solve step sequence of the interpolating adjoint: kernel vs oracle (fp64)."""
import os, sys, ctypes
os.environ["LDEQ_CADJ_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from oracle import goku as og, mlp as om
rng = np.random.Generator(np.random.PCG64(1))
dims = [16, 200, 200, 16]
layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]).astype(np.float64), (0.1 * rng.standard_normal(dims[i + 1]))) for i in range(3)]
p = om.pack_params(layers).astype(np.float64)
B, T = 24, 20
z0 = 0.5 * rng.standard_normal((B, 16)); t = 0.05 * np.arange(T); d = rng.standard_normal((T, B, 16))
opts = ldeq.default_opts(sensealg=ldeq.SENSE_INTERPOLATING_ADJOINT, controller_pow=1)
z = torch.from_numpy(z0).cuda(); pp = torch.from_numpy(p).cuda()
traj, st, tape = ldeq.mlp_solve_raw(z, pp, dims, t, opts, want_tape=True)
gz, gp = ldeq.mlp_bwd_raw(tape, torch.from_numpy(d).cuda())
na, nr, ret = ldeq.mlp_bwd_stats(tape)
n = na + nr
buf = (ctypes.c_double * (4 * n))()
tape.h.check(tape.h._lib.ldeq_debug_cadj_trace(tape.h.ptr, tape.ptr, buf, n))
tr = np.array(buf).reshape(n, 4)
otr, ona, _, otape = om.solve(z0, p, dims, t, og.Opts(controller_pow=1), record=True)
print("fwd naccept", int(st.naccept[0]), ona, "tape t diff", np.abs(np.array(otape.t) - tape_t).max() if False else "")
sto = {}
oz, op = om.interpolating_adjoint(z0, p, dims, t, d, og.Opts(controller_pow=1), tape=otape, stats=sto)
otrace = np.array(sto["trace"])
print("kernel", na, nr, "oracle", sto["naccept"], sto["nreject"])
for i in range(min(40, n, len(otrace))):
    print(i, "K %.9f %.6e %.6e %d" % tuple(tr[i]), " O %.9f %.6e %.6e %d" % tuple(otrace[i]))
