"""Small invocation of every CUDA entry point, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
from conftest import pendulum_inputs
dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "goku"):
    for dtype in ("float32", "float64"):
        for B, T in ((129, 23), (5, 2), (70, 200)):
            z0, th = pendulum_inputs(B, dtype=dtype)
            t = np.cumsum(np.r_[0.0, np.linspace(0.01, 0.09, T - 1)]) if B == 129 else 0.05 * np.arange(T)
            for kw in (dict(sensealg=0), dict(sensealg=0, adaptive=False, dt=0.07), dict(sensealg=0, tape_steps=2), dict()):
                z = torch.from_numpy(z0).to(dev).requires_grad_(True); p = torch.from_numpy(th).to(dev).requires_grad_(True)
                tr = ldeq.goku_solve(z, p, t, 1, ldeq.default_opts(**kw))
                tr.backward(torch.ones_like(tr))
    h = ldeq.handle(0)
    rhs = h.rhs_from_source("template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) { du[0] = u[1]; du[1] = -p[0]*sin(u[0]); du[2] = p[1]*u[0] - u[2]; }", 3, 2)
    z = torch.randn(37, 3, device=dev, requires_grad=True); p = (1 + torch.rand(37, 2, device=dev)).requires_grad_(True)
    tr = ldeq.goku_solve(z, p, 0.05 * np.arange(30), rhs, ldeq.default_opts(sensealg=0)); tr.backward(torch.ones_like(tr))
    # forward-dual pullback: built-in (both dtypes, friction) and the user function
    for dtype in ("float32", "float64"):
        z0, th = pendulum_inputs(131, dtype=dtype)
        for kind in (0, 1):
            z = torch.from_numpy(z0).to(dev).requires_grad_(True); p = torch.from_numpy(th).to(dev).requires_grad_(True)
            tr = ldeq.goku_solve(z, p, 0.05 * np.arange(23), kind, ldeq.default_opts(sensealg=1)); tr.backward(torch.ones_like(tr))
    z = torch.randn(37, 3, device=dev, requires_grad=True); p = (1 + torch.rand(37, 2, device=dev)).requires_grad_(True)
    tr = ldeq.goku_solve(z, p, 0.05 * np.arange(30), rhs, ldeq.default_opts(sensealg=1)); tr.backward(torch.ones_like(tr))
    # host-buffer entry points, several column slabs, both pullbacks, and the trig diagnostics
    z0, th = pendulum_inputs(70001)
    tt = 0.05 * np.arange(9)
    dd = torch.randn(9, 70001, 2)
    for sense in (0, 1):
        o = ldeq.default_opts(sensealg=sense)
        out, tape = ldeq.goku_solve_host(torch.from_numpy(z0), torch.from_numpy(th), tt, 0, o, want_tape=True)
        ldeq.goku_bwd_host(tape, dd)
        ldeq.goku_fwd_bwd_host(torch.from_numpy(z0), torch.from_numpy(th), tt, dd, 0, o)
    for w in (0, 1, 2):
        ldeq.debug_trig(torch.linspace(-20, 20, 1001, device=dev), w)
if which in ("all", "solvers"):
    # the diffeq struct's other solver values (csrc/ldeq_erk.cuh): forward, tape overflow + healing, both pullbacks, user RHS
    for dtype in ("float32", "float64"):
        z0, th = pendulum_inputs(131, dtype=dtype)
        for sv in (1, 2, 3):
            for kw in ((dict(adaptive=False, dt=0.07),) if sv == 3 else (dict(), dict(tape_steps=2), dict(adaptive=False, dt=0.07))):
                for sense in (0, 1):
                    z = torch.from_numpy(z0).to(dev).requires_grad_(True); p = torch.from_numpy(th).to(dev).requires_grad_(True)
                    tr = ldeq.goku_solve(z, p, 0.05 * np.arange(23), 1, ldeq.default_opts(solver=sv, sensealg=sense, **kw)); tr.backward(torch.ones_like(tr))
    h = ldeq.handle(0)
    rhs = h.rhs_from_source("template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) { du[0] = u[1]; du[1] = -p[0]*sin(u[0]); du[2] = p[1]*u[0] - u[2]; }", 3, 2)
    for sv in (1, 2):
        for sense in (0, 1):
            z = torch.randn(37, 3, device=dev, requires_grad=True); p = (1 + torch.rand(37, 2, device=dev)).requires_grad_(True)
            tr = ldeq.goku_solve(z, p, 0.05 * np.arange(30), rhs, ldeq.default_opts(solver=sv, sensealg=sense)); tr.backward(torch.ones_like(tr))
if which in ("all", "recurrent"):
    from oracle import recurrent as orr
    from latentdiffeq_jl_b200.solve import _PatternExtractor
    rng = np.random.default_rng(0)
    for F in (16, 32, 64):
        for B, T in ((37, 7), (1, 1), (70, 3)):
            x = torch.randn(T, B, F, device=dev, requires_grad=True)
            ps = [torch.from_numpy(orr.init_params(l, F, rng)).to(dev).requires_grad_(True) for l in (False, True, True)]
            z0, th = _PatternExtractor.apply(x, *ps); (z0.sum() + th.sum()).backward()
            z0, _ = _PatternExtractor.apply(x, ps[0], None, None); z0.sum().backward()
            if F >= 32:    # LatentODE's RNN stack with 32 hidden units
                q = torch.from_numpy(orr.init_params(False, F, rng, H=32)).to(dev).requires_grad_(True)
                z0, _ = _PatternExtractor.apply(x, q, None, None); z0.sum().backward()
if which in ("all", "mlp"):
    from oracle import mlp as om
    rng = np.random.Generator(np.random.PCG64(1)); dims = [16, 200, 200, 16]
    pp = torch.from_numpy(om.pack_params([(om.glorot_uniform(rng, dims[i + 1], dims[i]), np.zeros(dims[i + 1], np.float32)) for i in range(3)]).astype(np.float32)).to(dev)
    for dims2 in ([6, 50, 30, 6], [8, 64, 64, 32, 8], [10, 10]):
        q = (0.1 * torch.randn(om.n_params(dims2), device=dev)).requires_grad_(True)
        z = (0.5 * torch.randn(7, dims2[0], device=dev)).requires_grad_(True)
        tr = ldeq.mlp_solve(z, q, dims2, 0.05 * np.arange(9), ldeq.default_opts(norm_mode=1)); tr.backward(torch.ones_like(tr))
    for B in (3, 130):
        for kw in (dict(norm_mode=0), dict(norm_mode=1), dict(norm_mode=1, mlp_math=1), dict(norm_mode=0, mlp_math=1), dict(adaptive=False, dt=0.05, mlp_math=1)):
            z = (0.5 * torch.randn(B, 16, device=dev)).requires_grad_(True); q = pp.clone().requires_grad_(True)
            tr = ldeq.mlp_solve(z, q, dims, 0.05 * np.arange(12), ldeq.default_opts(**kw)); tr.backward(torch.ones_like(tr))
    # the reference's continuous adjoint (LDEQ_SENSE_INTERPOLATING_ADJOINT): resident-weights variant (Float32, TB = 2),
    # general variant with stage records (Float64) and without (LDEQ_CADJ_BATCH=0 is exercised by the A/B runs)
    for dt_, dims2 in ((torch.float32, [8, 24, 24, 8]), (torch.float64, [6, 20, 20, 6])):
        q = (0.3 * torch.randn(om.n_params(dims2), device=dev, dtype=dt_)).requires_grad_(True)
        z = (0.5 * torch.randn(5, dims2[0], device=dev, dtype=dt_)).requires_grad_(True)
        tr = ldeq.mlp_solve(z, q, dims2, 0.05 * np.arange(6), ldeq.default_opts(norm_mode=0, sensealg=ldeq.SENSE_INTERPOLATING_ADJOINT))
        tr.backward(torch.ones_like(tr))
    z = torch.randn(9, 5, device=dev, dtype=torch.float64, requires_grad=True)
    q = torch.randn(om.n_params([5, 33, 5]), device=dev, dtype=torch.float64).mul_(0.1).requires_grad_(True)
    tr = ldeq.mlp_solve(z, q, [5, 33, 5], 0.05 * np.arange(8)); tr.backward(torch.ones_like(tr))
if which in ("all", "loss"):
    x = torch.rand(7, 5, 101, device=dev); xh = torch.rand(7, 5, 101, device=dev)
    mu = [torch.randn(5, 16, device=dev), torch.randn(5, 16, device=dev)]; lv = [torch.randn(5, 16, device=dev) * .1 for _ in range(2)]
    ldeq.elbo_raw(x, xh, mu, lv, 0.5)
    n = 1003; n4 = 1004
    a = [torch.randn(n4, device=dev) for _ in range(4)]
    ldeq.adamw_step(a[0], a[1], a[2].abs(), a[3].abs(), 3)
    ldeq.sample_raw(mu[0], lv[0], 1, 0)
torch.cuda.synchronize()
print("sanitize run done:", which, "launches", ldeq.handle(0).launch_count())
