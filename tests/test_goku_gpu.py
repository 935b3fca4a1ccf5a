"""GPU parity tests of the GOKU hot path: CUDA (through the C ABI) vs the CPU oracle on the same
seeded inputs.  Tolerances are the north star's: trajectories rtol 1e-5 in fp64 / 1e-3 in fp32,
identical accepted-step counts in fixed-step mode, gradients 1e-4 relative."""
import numpy as np
import pytest
import torch

from conftest import pendulum_inputs
from oracle import goku as og

DEV = "cuda:0"

pytestmark = pytest.mark.gpu


def _run(ldeq, rhs, z0, th, t, **kw):
    dev = torch.device("cuda:0")
    opts = ldeq.default_opts(**kw)
    traj, st, _ = ldeq.goku_solve_raw(torch.from_numpy(z0).to(dev), torch.from_numpy(th).to(dev), t, rhs, opts)
    torch.cuda.synchronize()
    return traj.cpu().numpy(), st.retcode.cpu().numpy(), st.naccept.cpu().numpy(), st.nreject.cpu().numpy()


def _close(a, b, rtol, atol_scale=1.0):
    scale = np.abs(b).max()
    return np.abs(a - b).max() <= rtol * scale * atol_scale


@pytest.mark.parametrize("rhs", [0, 1])
@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-5), ("float64", 1e-11)])
def test_fixed_step_matches_oracle(ldeq, rhs, dtype, rtol):
    # C1 / C3 shapes, fixed step dt = 0.05: step counts must be identical, values tight
    B, T = 1024, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    tr, ret, na, nr = _run(ldeq, rhs, z0, th, t, adaptive=False, dt=0.05)
    otr, oret, ona, onr = og.solve(rhs, z0, th, t, og.Opts(adaptive=False, dt=0.05))
    assert (ret == 0).all() and (oret == 0).all()
    assert (na == ona).all() and (na == T - 1).all()
    assert np.abs(tr - otr).max() <= rtol * np.abs(otr).max()


@pytest.mark.parametrize("rhs", [0, 1])
@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-3), ("float64", 1e-5)])
def test_adaptive_matches_oracle(ldeq, rhs, dtype, rtol):
    # north star tolerances: fp64 rtol 1e-5, fp32 rtol 1e-3 (C3: B=1024, fp64 vs fp32)
    B, T = 1024, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    tr, ret, na, nr = _run(ldeq, rhs, z0, th, t)
    otr, oret, ona, onr = og.solve(rhs, z0, th, t)
    assert (ret == 0).all()
    assert np.abs(tr - otr).max() <= rtol * np.abs(otr).max()
    # same controller => the step sequences agree except for accept/reject flips at EEst ~ 1 (the oracle
    # rounds the fp32 stage arithmetic through Float64 as Julia does, the kernel stays in fp32)
    assert (na == ona).mean() > (0.9 if dtype == "float32" else 0.99)
    print("max abs diff", np.abs(tr - otr).max(), "naccept equal frac", (na == ona).mean())


def test_c4_shape_forward_vs_oracle(ldeq):
    # C4 sweep shape at the size the oracle still finishes in seconds: N = 2^16, T = 200
    B, T = 1 << 16, 200
    z0, th = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    tr, ret, na, nr = _run(ldeq, 0, z0, th, t)
    otr, oret, ona, onr = og.solve(0, z0, th, t)
    assert (ret == 0).all()
    assert np.abs(tr - otr).max() <= 1e-3 * np.abs(otr).max()
    assert (na == ona).mean() > 0.9


def _grads(ldeq, rhs, z0, th, t, d, **kw):
    dev = torch.device("cuda:0")
    opts = ldeq.default_opts(**kw)
    z = torch.from_numpy(z0).to(dev).requires_grad_(True)
    p = torch.from_numpy(th).to(dev).requires_grad_(True)
    traj = ldeq.goku_solve(z, p, t, rhs, opts)
    traj.backward(torch.from_numpy(d).to(dev))
    torch.cuda.synchronize()
    return z.grad.cpu().numpy(), p.grad.cpu().numpy()


@pytest.mark.parametrize("rhs", [0, 1])
@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-4), ("float64", 1e-10)])
def test_adjoint_fixed_step_matches_forward_sensitivity(ldeq, rhs, dtype, rtol):
    # fixed-step mode: the discrete adjoint IS the derivative ForwardDiffSensitivity computes
    B, T = 512, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(dtype)
    gz, gp = _grads(ldeq, rhs, z0, th, t, d, adaptive=False, dt=0.05)
    oz, op = og.grad(rhs, z0, th, t, d, og.Opts(adaptive=False, dt=0.05))
    assert np.abs(gz - oz).max() <= rtol * np.abs(oz).max()
    assert np.abs(gp - op).max() <= rtol * np.abs(op).max()


@pytest.mark.parametrize("dtype,rtol", [("float32", 2e-4), ("float64", 1e-9)])
def test_adjoint_adaptive_matches_frozen_step_sensitivity(ldeq, dtype, rtol):
    # adaptive mode: exact derivative of the primal discretisation (oracle: duals excluded from the norm)
    B, T = 512, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(dtype)
    gz, gp = _grads(ldeq, 0, z0, th, t, d)
    oz, op = og.grad(0, z0, th, t, d, norm_partials=False)
    # a trajectory whose accept/reject decision flipped has a different step sequence: compare the bulk
    ez = np.abs(gz - oz).max(1) / np.abs(oz).max()
    ep = np.abs(gp - op).max(1) / np.abs(op).max()
    assert np.quantile(ez, 0.98) <= rtol and np.quantile(ep, 0.98) <= rtol
    # and every trajectory agrees with the reference's ForwardDiff semantics at the solver tolerance
    rz, rp = og.grad(0, z0, th, t, d, norm_partials=True)
    assert np.abs(gz - rz).max() <= 2e-2 * np.abs(rz).max()
    assert np.abs(gp - rp).max() <= 2e-2 * np.abs(rp).max()


def test_adjoint_tight_tolerance_matches_reference_semantics(ldeq):
    # at abstol = reltol = 1e-10 (fp64) the reference's dual-number gradient and the discrete adjoint agree to 1e-4
    B, T = 256, 50
    z0, th = pendulum_inputs(B, dtype="float64")
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2))
    gz, gp = _grads(ldeq, 0, z0, th, t, d, abstol=1e-10, reltol=1e-10)
    rz, rp = og.grad(0, z0, th, t, d, og.Opts(abstol=1e-10, reltol=1e-10), norm_partials=True)
    assert np.abs(gz - rz).max() <= 1e-4 * np.abs(rz).max()
    assert np.abs(gp - rp).max() <= 1e-4 * np.abs(rp).max()


@pytest.mark.parametrize("rhs", [0, 1])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_forward_dual_mode_is_the_reference_gradient(ldeq, rhs, dtype):
    # sensealg = LDEQ_SENSE_FORWARD_DUAL: the reference's own algorithm (two dual-number re-solves per trajectory whose
    # error norm includes the partials) restated literally -- against the oracle's ForwardDiff-semantics gradient at the
    # DEFAULT tolerance (abstol 1e-6, reltol 1e-3), where the discrete adjoint only agrees to 2e-2.
    B, T = 512, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(dtype)
    gz, gp = _grads(ldeq, rhs, z0, th, t, d, sensealg=ldeq.SENSE_FORWARD_DUAL)
    rz, rp = og.grad(rhs, z0, th, t, d, norm_partials=True)
    ez = np.abs(gz - rz).max(1) / np.abs(rz).max()
    ep = np.abs(gp - rp).max(1) / np.abs(rp).max()
    if dtype == "float64":
        # north star: gradients within 1e-4 relative -- every trajectory, by a wide margin
        assert ez.max() <= 1e-9 and ep.max() <= 1e-9
    else:
        # Float32: sinf/cosf of CUDA and glibc differ in the last bit, which moves an accept/reject decision on a few
        # trajectories (a different, equally valid step sequence); the bulk agrees to rounding
        assert np.quantile(ez, 0.95) <= 2e-5 and np.quantile(ep, 0.95) <= 2e-5
        assert ez.max() <= 2e-2 and ep.max() <= 2e-2
    # the trajectories themselves are those of the default mode (the primal solve is the same kernel)
    a, _, _ = ldeq.goku_solve_raw(torch.from_numpy(z0).to(DEV), torch.from_numpy(th).to(DEV), t, rhs)
    b, _, _ = ldeq.goku_solve_raw(torch.from_numpy(z0).to(DEV), torch.from_numpy(th).to(DEV), t, rhs,
                                  ldeq.default_opts(sensealg=ldeq.SENSE_FORWARD_DUAL))
    assert torch.equal(a, b)


def test_forward_dual_mode_fixed_step_and_failures(ldeq):
    B, T = 100, 30
    z0, th = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(1).standard_normal((T, B, 2)).astype(np.float32)
    # fixed step: both sensitivity algorithms are the same derivative
    g1 = _grads(ldeq, 0, z0, th, t, d, adaptive=False, dt=0.05, sensealg=ldeq.SENSE_FORWARD_DUAL)
    g0 = _grads(ldeq, 0, z0, th, t, d, adaptive=False, dt=0.05)
    for a, b in zip(g1, g0):
        assert np.abs(a - b).max() <= 5e-6 * np.abs(b).max()
    # a trajectory that fails in the primal solve: zero gradient, the others untouched
    th2 = th.copy()
    th2[3, 0] = 1e-4
    gz, gp = _grads(ldeq, 0, z0, th2, t, d, maxiters=200, sensealg=ldeq.SENSE_FORWARD_DUAL)
    assert (gz[3] == 0).all() and (gp[3] == 0).all() and np.isfinite(gz).all() and np.abs(gz[4]).max() > 0


def test_forward_dual_mode_with_user_rhs(ldeq):
    # the same pendulum written as a user function (NVRTC): the dual-number pullback agrees with the built-in one;
    # a 3-dimensional system with 2 parameters in Float64: forward-dual == discrete adjoint in fixed-step mode
    h = ldeq.handle(0)
    r = h.rhs_from_source("template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) "
                          "{ du[0] = u[1]; du[1] = -S(10.0f) / p[0] * sin(u[0]); }", 2, 1)
    B, T = 256, 50
    z0, th = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(1).standard_normal((T, B, 2)).astype(np.float32)
    gu = _grads(ldeq, r, z0, th, t, d, sensealg=ldeq.SENSE_FORWARD_DUAL)
    gb = _grads(ldeq, 0, z0, th, t, d, sensealg=ldeq.SENSE_FORWARD_DUAL)
    for a, b in zip(gu, gb):
        e = np.abs(a - b).max(1) / np.abs(b).max()
        assert np.quantile(e, 0.95) <= 2e-5 and e.max() <= 2e-2   # fused vs unfused multiply-adds: last-bit differences
    r3 = h.rhs_from_source("template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) "
                           "{ du[0] = u[1]; du[1] = -p[0] * sin(u[0]) - S(0.1) * u[1]; du[2] = p[1] * u[0] - u[2] + exp(-t); }", 3, 2)
    rng = np.random.default_rng(7)
    B, T = 64, 20
    z3 = rng.uniform(-0.5, 0.5, (B, 3))
    p3 = rng.uniform(1.0, 2.0, (B, 2))
    t = 0.05 * np.arange(T)
    d3 = rng.standard_normal((T, B, 3))
    g1 = _grads(ldeq, r3, z3, p3, t, d3, adaptive=False, dt=0.025, sensealg=ldeq.SENSE_FORWARD_DUAL)
    g0 = _grads(ldeq, r3, z3, p3, t, d3, adaptive=False, dt=0.025)
    for a, b in zip(g1, g0):
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    # adaptive: the two algorithms agree within the solver tolerance
    g1 = _grads(ldeq, r3, z3, p3, t, d3, sensealg=ldeq.SENSE_FORWARD_DUAL)
    g0 = _grads(ldeq, r3, z3, p3, t, d3)
    for a, b in zip(g1, g0):
        assert np.abs(a - b).max() <= 2e-2 * np.abs(b).max()


def test_failed_trajectory_is_nan_block_with_zero_gradient(ldeq):
    # GOKU.jl:114: a non-Success retcode yields a NaN (z,T) block; here maxiters forces the failure
    B, T = 64, 50
    z0, th = pendulum_inputs(B)
    th[3, 0] = 1e-4  # G/L = 1e5: needs far more than 50 steps
    t = 0.05 * np.arange(T)
    tr, ret, na, nr = _run(ldeq, 0, z0, th, t, maxiters=50)
    otr, oret, _, _ = og.solve(0, z0, th, t, og.Opts(maxiters=50))
    assert ret[3] == 1 and oret[3] == 1
    assert np.isnan(tr[:, 3, :]).all() and np.isnan(otr[:, 3, :]).all()
    ok = np.arange(B) != 3
    assert np.isfinite(tr[:, ok, :]).all()
    d = np.ones((T, B, 2), dtype=np.float32)
    gz, gp = _grads(ldeq, 0, z0, th, t, d, maxiters=50)
    assert (gz[3] == 0).all() and (gp[3] == 0).all() and np.isfinite(gz).all()


def test_edge_shapes(ldeq):
    # B = 1, ragged B (not a multiple of the block), T = 1 and T = 2, non-uniform grid
    for B, T in [(1, 50), (129, 7), (5, 1), (5, 2)]:
        z0, th = pendulum_inputs(B)
        t = np.cumsum(np.r_[0.0, np.linspace(0.01, 0.2, T - 1)]) if T > 1 else np.zeros(1)
        tr, ret, na, nr = _run(ldeq, 0, z0, th, t)
        otr, oret, ona, _ = og.solve(0, z0, th, t)
        assert (ret == 0).all()
        assert np.abs(tr - otr).max() <= 1e-3 * max(np.abs(otr).max(), 1e-30)


def test_host_entry_points_match_device(ldeq):
    B, T = 4096, 50
    z0, th = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(5).standard_normal((T, B, 2)).astype(np.float32)
    out, tape = ldeq.goku_solve_host(torch.from_numpy(z0).pin_memory(), torch.from_numpy(th).pin_memory(), t, 0,
                                     want_tape=True)
    dz0, dth = ldeq.goku_bwd_host(tape, torch.from_numpy(d).pin_memory())
    tr, *_ = _run(ldeq, 0, z0, th, t)
    gz, gp = _grads(ldeq, 0, z0, th, t, d)
    assert np.array_equal(out.numpy(), tr)
    assert np.array_equal(dz0.numpy(), gz) and np.array_equal(dth.numpy(), gp)


def test_small_tape_heals_itself(ldeq):
    # a tape that is too small is replayed into a larger one before the backward pass: same gradients
    B, T = 300, 50
    z0, th = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(7).standard_normal((T, B, 2)).astype(np.float32)
    g_big = _grads(ldeq, 0, z0, th, t, d, tape_steps=512)
    g_small = _grads(ldeq, 0, z0, th, t, d, tape_steps=3)
    assert np.array_equal(g_big[0], g_small[0]) and np.array_equal(g_big[1], g_small[1])
    assert np.isfinite(g_small[0]).all()


@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-4), ("float64", 1e-10)])
def test_adjoint_fixed_step_off_grid(ldeq, dtype, rtol):
    # dt = 0.07 does not divide the 0.05 save grid: every save point goes through the dense interpolant,
    # some steps contain no save point at all, and the last step is truncated onto t_end
    B, T = 200, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(12).standard_normal((T, B, 2)).astype(dtype)
    for dt in (0.07, 0.013):
        gz, gp = _grads(ldeq, 1, z0, th, t, d, adaptive=False, dt=dt)
        oz, op = og.grad(1, z0, th, t, d, og.Opts(adaptive=False, dt=dt))
        assert np.abs(gz - oz).max() <= rtol * np.abs(oz).max()
        assert np.abs(gp - op).max() <= rtol * np.abs(op).max()
        tr, ret, na, nr = _run(ldeq, 1, z0, th, t, adaptive=False, dt=dt)
        otr, oret, ona, _ = og.solve(1, z0, th, t, og.Opts(adaptive=False, dt=dt))
        assert (na == ona).all()
        assert np.abs(tr - otr).max() <= (1e-5 if dtype == "float32" else 1e-11) * np.abs(otr).max()


def test_adjoint_every_trajectory_adaptive_fp64(ldeq):
    # regression: the last-finishing lanes of a warp must still sweep their early steps (which hold no
    # save point when the automatic first step is shorter than the grid spacing)
    B, T = 512, 50
    z0, th = pendulum_inputs(B, dtype="float64")
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2))
    gz, gp = _grads(ldeq, 0, z0, th, t, d, controller_pow=1)
    oz, op = og.grad(0, z0, th, t, d, og.Opts(controller_pow=1), norm_partials=False)
    assert np.abs(gz - oz).max() <= 1e-5 * np.abs(oz).max()
    assert np.abs(gp - op).max() <= 1e-5 * np.abs(op).max()
