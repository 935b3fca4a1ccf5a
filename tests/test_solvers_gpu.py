"""GPU parity of the diffeq struct's other `solver` values: OrdinaryDiffEq's DP5 / BS3 / RK4 in the
GOKU integrator kernels (csrc/ldeq_erk.cuh) through the C ABI, against the oracle's restatement on the same seeded
inputs.  Same tolerances as the Tsit5 tests: trajectories rtol 1e-5 (fp64) / 1e-3 (fp32), identical accepted-step
counts in fixed-step mode, gradients 1e-4 relative."""
import numpy as np
import pytest
import torch

from conftest import pendulum_inputs
from oracle import goku as og

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SOLVERS = {"DP5": 1, "BS3": 2, "RK4": 3}


def _fwd(ldeq, rhs, z0, th, t, **kw):
    traj, st, _ = ldeq.goku_solve_raw(torch.from_numpy(z0).to(DEV), torch.from_numpy(th).to(DEV), t, rhs, ldeq.default_opts(**kw))
    torch.cuda.synchronize()
    return traj.cpu().numpy(), st.retcode.cpu().numpy(), st.naccept.cpu().numpy(), st.nreject.cpu().numpy()


def _grads(ldeq, rhs, z0, th, t, d, **kw):
    z = torch.from_numpy(z0).to(DEV).requires_grad_(True)
    p = torch.from_numpy(th).to(DEV).requires_grad_(True)
    ldeq.goku_solve(z, p, t, rhs, ldeq.default_opts(**kw)).backward(torch.from_numpy(d).to(DEV))
    torch.cuda.synchronize()
    return z.grad.cpu().numpy(), p.grad.cpu().numpy()


def test_enums_and_controller_defaults(ldeq):
    assert (ldeq.SOLVER_DP5, ldeq.SOLVER_BS3, ldeq.SOLVER_RK4) == (og.DP5, og.BS3, og.RK4)
    for name, sv in SOLVERS.items():
        o = ldeq.default_opts(solver=sv)
        assert (o.beta1, o.beta2) == pytest.approx(og.CONTROLLER_DEFAULTS[sv]) and o.solver == sv
    o = ldeq.default_opts()
    assert (o.beta1, o.beta2, o.solver) == (7 / 50, 2 / 25, ldeq.SOLVER_TSIT5)
    assert ldeq.DP5().code == 1 and ldeq.BS3().code == 2 and ldeq.RK4().code == 3


@pytest.mark.parametrize("name", list(SOLVERS))
@pytest.mark.parametrize("rhs", [0, 1])
@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-5), ("float64", 1e-11)])
def test_fixed_step_matches_oracle(ldeq, name, rhs, dtype, rtol):
    sv = SOLVERS[name]
    B, T = 1024, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    for dt in (0.05, 0.08):     # on the save grid, and off it (every save point through the dense output)
        tr, ret, na, nr = _fwd(ldeq, rhs, z0, th, t, solver=sv, adaptive=False, dt=dt)
        otr, oret, ona, onr = og.solve(rhs, z0, th, t, og.Opts.for_solver(sv, adaptive=False, dt=dt))
        assert (ret == 0).all() and (oret == 0).all()
        assert (na == ona).all()
        assert np.abs(tr - otr).max() <= rtol * np.abs(otr).max(), (dt, np.abs(tr - otr).max())


@pytest.mark.parametrize("name", ["DP5", "BS3"])
@pytest.mark.parametrize("rhs", [0, 1])
@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-3), ("float64", 1e-5)])
def test_adaptive_matches_oracle(ldeq, name, rhs, dtype, rtol):
    sv = SOLVERS[name]
    B, T = 1024, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    tr, ret, na, nr = _fwd(ldeq, rhs, z0, th, t, solver=sv)
    otr, oret, ona, onr = og.solve(rhs, z0, th, t, og.Opts.for_solver(sv))
    assert (ret == 0).all()
    assert np.abs(tr - otr).max() <= rtol * np.abs(otr).max()
    frac = (na == ona).mean()
    print(name, dtype, "max abs diff", np.abs(tr - otr).max(), "naccept equal frac", frac, "mean naccept", na.mean())
    assert frac > (0.85 if dtype == "float32" else 0.99)


def test_rk4_adaptive_and_latentode_other_solvers_are_refused(ldeq):
    z0, th = pendulum_inputs(8)
    t = 0.05 * np.arange(10)
    with pytest.raises(ldeq.LdeqError) as e:
        _fwd(ldeq, 0, z0, th, t, solver=ldeq.SOLVER_RK4)
    assert e.value.code == -3 and "fixed step" in str(e.value)
    with pytest.raises(ldeq.LdeqError) as e:
        _fwd(ldeq, 0, z0, th, t, solver=7)
    assert e.value.code in (-1, -3)
    dims = [4, 8, 4]
    n = sum(dims[i + 1] * dims[i] + dims[i + 1] for i in range(2))
    with pytest.raises(ldeq.LdeqError) as e:
        ldeq.mlp_solve_raw(torch.zeros(8, 4, device=DEV), torch.zeros(n, device=DEV), dims, t, ldeq.default_opts(solver=ldeq.SOLVER_DP5))
    assert e.value.code == -3


@pytest.mark.parametrize("name", list(SOLVERS))
@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-4), ("float64", 1e-10)])
def test_gradients_fixed_step_both_sensitivity_modes(ldeq, name, dtype, rtol):
    # fixed step (off the save grid: the dense output is differentiated too): the discrete adjoint and the dual solves
    # are the same derivative, the oracle's forward sensitivities
    sv = SOLVERS[name]
    B, T = 512, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(dtype)
    oz, op = og.grad(1, z0, th, t, d, og.Opts.for_solver(sv, adaptive=False, dt=0.08))
    for sense in (ldeq.SENSE_DISCRETE_ADJOINT, ldeq.SENSE_FORWARD_DUAL):
        gz, gp = _grads(ldeq, 1, z0, th, t, d, solver=sv, adaptive=False, dt=0.08, sensealg=sense)
        assert np.abs(gz - oz).max() <= rtol * np.abs(oz).max(), (sense, np.abs(gz - oz).max() / np.abs(oz).max())
        assert np.abs(gp - op).max() <= rtol * np.abs(op).max(), (sense, np.abs(gp - op).max() / np.abs(op).max())


@pytest.mark.parametrize("name", ["DP5", "BS3"])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_adaptive_gradients(ldeq, name, dtype):
    sv = SOLVERS[name]
    B, T = 512, 50
    z0, th = pendulum_inputs(B, dtype=dtype)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(dtype)
    o = og.Opts.for_solver(sv)
    # the library default: the reference's dual-number re-solves (partials in the error norm) vs the oracle's
    gz, gp = _grads(ldeq, 0, z0, th, t, d, solver=sv)
    rz, rp = og.grad(0, z0, th, t, d, o, norm_partials=True)
    ez = np.abs(gz - rz).max(1) / np.abs(rz).max()
    ep = np.abs(gp - rp).max(1) / np.abs(rp).max()
    print(name, dtype, "forward-dual q99", np.quantile(ez, 0.99), np.quantile(ep, 0.99), "max", ez.max(), ep.max())
    if dtype == "float64":
        assert ez.max() <= 1e-8 and ep.max() <= 1e-8
    else:
        assert np.quantile(ez, 0.98) <= 1e-4 and np.quantile(ep, 0.98) <= 1e-4
        assert ez.max() <= 5e-2 and ep.max() <= 5e-2
    # the opt-in: discrete adjoint of the taped steps vs the exact derivative of the primal discretisation
    gz, gp = _grads(ldeq, 0, z0, th, t, d, solver=sv, sensealg=ldeq.SENSE_DISCRETE_ADJOINT)
    fz, fp = og.grad(0, z0, th, t, d, o, norm_partials=False)
    ez = np.abs(gz - fz).max(1) / np.abs(fz).max()
    ep = np.abs(gp - fp).max(1) / np.abs(fp).max()
    tol = 1e-9 if dtype == "float64" else 2e-4
    assert np.quantile(ez, 0.97) <= tol and np.quantile(ep, 0.97) <= tol, (np.quantile(ez, 0.97), np.quantile(ep, 0.97))


USER_SRC = r"""
template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) {
    const S G = S(10.0f);
    du[0] = u[1];
    du[1] = -G / p[0] * sin(u[0]) - S(0.7f) * u[1];      // Pendulum_friction, pendulum.jl:65-74
}
"""


@pytest.mark.parametrize("name", ["DP5", "BS3", "RK4"])
def test_user_rhs_under_other_solvers(ldeq, name):
    # an NVRTC right-hand side compiles one module per solver on first use: equal to the built-in under the same solver
    sv = SOLVERS[name]
    rhs = ldeq.handle(0).rhs_from_source(USER_SRC, 2, 1)
    B, T = 300, 40
    z0, th = pendulum_inputs(B, dtype="float64")
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(0).standard_normal((T, B, 2))
    kws = [dict(adaptive=False, dt=0.07)] + ([] if name == "RK4" else [dict()])
    for kw in kws:
        for sense in (ldeq.SENSE_DISCRETE_ADJOINT, ldeq.SENSE_FORWARD_DUAL):
            a = _fwd(ldeq, rhs, z0, th, t, solver=sv, **kw)[0]
            b = _fwd(ldeq, ldeq.RHS_PENDULUM_FRICTION, z0, th, t, solver=sv, **kw)[0]
            assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()
            ga = _grads(ldeq, rhs, z0, th, t, d, solver=sv, sensealg=sense, **kw)
            gb = _grads(ldeq, ldeq.RHS_PENDULUM_FRICTION, z0, th, t, d, solver=sv, sensealg=sense, **kw)
            for x, y in zip(ga, gb):
                assert np.abs(x - y).max() <= 1e-8 * np.abs(y).max()


def test_solver_field_of_the_diffeq_struct_reaches_the_kernel(ldeq):
    # Pendulum(solver = DP5()) in the model's diffeq_layer: same trajectories as the raw call with solver = DP5
    torch.manual_seed(0)
    diffeq = ldeq.Pendulum(solver=ldeq.DP5())
    enc, dec = ldeq.default_layers(ldeq.GOKU(), 28 * 28, diffeq, device=DEV)
    model = ldeq.LatentDiffEqModel(ldeq.GOKU(), enc, dec)
    B, T = 16, 20
    z0, th = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    z = ldeq.diffeq_layer(model.decoder, (torch.from_numpy(z0).to(DEV), torch.from_numpy(th).to(DEV)), t)
    ref = _fwd(ldeq, 0, z0, th, t, solver=ldeq.SOLVER_DP5)[0]
    tsit = _fwd(ldeq, 0, z0, th, t)[0]
    assert np.array_equal(z.detach().cpu().numpy(), ref) and not np.array_equal(ref, tsit)
