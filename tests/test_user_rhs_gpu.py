"""User-defined right-hand sides (the reference's diffeq struct carries an arbitrary f!, pendulum.jl:19-26):
CUDA C source compiled at run time with NVRTC into the same integrator kernels."""
import numpy as np
import pytest
import torch
from scipy.integrate import solve_ivp

from conftest import pendulum_inputs
from oracle import goku as og

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

PENDULUM_SRC = r"""
template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) {
    const S G = S(10.0f);
    du[0] = u[1];
    du[1] = -G / p[0] * sin(u[0]) - S(0.7f) * u[1];      // Pendulum_friction, pendulum.jl:65-74
}
"""
LORENZ_SRC = r"""
template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) {
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = u[0] * (p[1] - u[2]) - u[1];
    du[2] = u[0] * u[1] - p[2] * u[2];
}
"""


def _solve(ldeq, rhs, z0, th, t, d=None, **kw):
    z = torch.from_numpy(z0).to(DEV).requires_grad_(d is not None)
    p = torch.from_numpy(th).to(DEV).requires_grad_(d is not None)
    st = []
    kw.setdefault("sensealg", ldeq.SENSE_DISCRETE_ADJOINT)   # these tests exercise the adjoint kernels of a user RHS
    tr = ldeq.goku_solve(z, p, t, rhs, ldeq.default_opts(**kw), st)
    if d is None:
        return tr.detach().cpu().numpy(), st[0]
    tr.backward(torch.from_numpy(d).to(DEV))
    return tr.detach().cpu().numpy(), z.grad.cpu().numpy(), p.grad.cpu().numpy()


@pytest.mark.parametrize("dtype,rtol", [("float64", 1e-10), ("float32", 1e-4)])
def test_user_pendulum_equals_builtin_and_oracle(ldeq, dtype, rtol):
    h = ldeq.handle(0)
    rhs = h.rhs_from_source(PENDULUM_SRC, 2, 1)
    B, T = 300, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(0).standard_normal((T, B, 2)).astype(dtype)
    for kw in (dict(adaptive=False, dt=0.07), dict()):
        a = _solve(ldeq, rhs, z0, th, t, d, **kw)
        b = _solve(ldeq, ldeq.RHS_PENDULUM_FRICTION, z0, th, t, d, **kw)
        for x, y in zip(a, b):
            assert np.abs(x - y).max() <= rtol * np.abs(y).max()
    # and against the oracle's forward sensitivities in fixed-step mode (dual-number VJP == hand-written VJP)
    tr, gz, gp = _solve(ldeq, rhs, z0, th, t, d, adaptive=False, dt=0.05)
    oz, op = og.grad(og.PENDULUM_FRICTION, z0, th, t, d, og.Opts(adaptive=False, dt=0.05))
    assert np.abs(gz - oz).max() <= rtol * np.abs(oz).max() and np.abs(gp - op).max() <= rtol * np.abs(op).max()


def test_user_lorenz_three_states_three_parameters(ldeq):
    h = ldeq.handle(0)
    rhs = h.rhs_from_source(LORENZ_SRC, 3, 3)
    rng = np.random.default_rng(1)
    B, T = 40, 30
    z0 = rng.uniform(-5, 5, (B, 3))
    th = np.stack([rng.uniform(9, 11, B), rng.uniform(20, 28, B), rng.uniform(2, 3, B)], 1)
    t = 0.02 * np.arange(T)
    tr, st = _solve(ldeq, rhs, z0, th, t, abstol=1e-10, reltol=1e-10)
    assert (st.retcode.cpu().numpy() == 0).all()
    for b in range(0, B, 7):
        s, r, be = th[b]
        ref = solve_ivp(lambda tt, u: [s * (u[1] - u[0]), u[0] * (r - u[2]) - u[1], u[0] * u[1] - be * u[2]], (0, t[-1]), z0[b],
                        method="DOP853", rtol=1e-12, atol=1e-12, t_eval=t).y.T
        assert np.abs(tr[:, b] - ref).max() <= 1e-6 * np.abs(ref).max()
    # gradient vs central finite differences of the same fixed-step CUDA solve (fp64)
    d = rng.standard_normal((T, B, 3))
    kw = dict(adaptive=False, dt=0.013)
    _, gz, gp = _solve(ldeq, rhs, z0, th, t, d, **kw)
    L = lambda z, p: (_solve(ldeq, rhs, z, p, t, **kw)[0] * d).sum(axis=(0, 2))   # noqa: E731
    eps = 1e-6
    for i in range(3):
        zp, zm, pp, pm = z0.copy(), z0.copy(), th.copy(), th.copy()
        zp[:, i] += eps; zm[:, i] -= eps; pp[:, i] += eps; pm[:, i] -= eps
        assert np.allclose((L(zp, th) - L(zm, th)) / (2 * eps), gz[:, i], rtol=2e-5, atol=1e-6)
        assert np.allclose((L(z0, pp) - L(z0, pm)) / (2 * eps), gp[:, i], rtol=2e-5, atol=1e-6)
    # fp32 instantiation of the same source
    tr32, _ = _solve(ldeq, rhs, z0.astype(np.float32), th.astype(np.float32), t)
    assert np.abs(tr32 - tr).max() <= 5e-3 * np.abs(tr).max()


def test_compile_error_is_reported(ldeq):
    h = ldeq.handle(0)
    with pytest.raises(ldeq.LdeqError) as e:
        h.rhs_from_source("template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) { du[0] = nope; }", 1, 1)
    assert e.value.code == -5 and "nope" in str(e.value)


def test_user_diffeq_struct_in_the_goku_model(ldeq):
    torch.manual_seed(0)
    mt = ldeq.GOKU_basic()
    diffeq = ldeq.UserDiffEq(LORENZ_SRC, u0=[1.0, 1.0, 1.0], p=[10.0, 28.0, 8.0 / 3.0], reltol=1e-4, abstol=1e-6)
    enc, dec = ldeq.default_layers(mt, 784, diffeq, device=DEV)
    model = ldeq.LatentDiffEqModel(mt, enc, dec)
    x = torch.rand(20, 6, 784, device=DEV)
    t = 0.02 * np.arange(20)
    (xh, zh, lh), mu, lv = model(x, t, True)
    assert zh.shape == (20, 6, 3) and lh[1].shape == (6, 3) and torch.isfinite(xh).all()
    ldeq.loss_batch(model, x, t, 1.0, True).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.decoder.latent_out.parameters())
