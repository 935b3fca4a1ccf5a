"""Independent pins of the oracle's DP5 / BS3 / RK4 restatements (CPU tier).

Like tests/test_oracle_pins.py for Tsit5: nothing here depends on the restatement being right -- Butcher order conditions
of the oracle's own tableaus (read back through ``oracle_method_table``), order of the embedded pairs, order conditions of
the dense outputs (probed from the oracle's own interpolation routine), convergence against scipy DOP853 @1e-12, the
adaptive error against the requested tolerance, and finite differences for the ForwardDiff-style gradient.
"""
import numpy as np
import pytest
from scipy.integrate import solve_ivp

from conftest import pendulum_inputs
from oracle import goku as og

SOLVERS = {"DP5": og.DP5, "BS3": og.BS3, "RK4": og.RK4}


def _conds(b, a, c, order):
    e, ac = np.ones(len(c)), a @ c
    out = [(b @ e, 1.0)]
    if order >= 2:
        out += [(b @ c, 1 / 2)]
    if order >= 3:
        out += [(b @ c ** 2, 1 / 3), (b @ ac, 1 / 6)]
    if order >= 4:
        out += [(b @ c ** 3, 1 / 4), (b @ (c * ac), 1 / 8), (b @ (a @ c ** 2), 1 / 12), (b @ (a @ ac), 1 / 24)]
    if order >= 5:
        out += [(b @ c ** 4, 1 / 5), (b @ (c ** 2 * ac), 1 / 10), (b @ (c * (a @ c ** 2)), 1 / 15),
                (b @ (c * (a @ ac)), 1 / 30), (b @ (ac * ac), 1 / 20), (b @ (a @ c ** 3), 1 / 20),
                (b @ (a @ (c * ac)), 1 / 40), (b @ (a @ (a @ c ** 2)), 1 / 60), (b @ (a @ (a @ ac)), 1 / 120)]
    return out


@pytest.mark.parametrize("name", list(SOLVERS))
def test_tableau_order_conditions(name):
    ns, order, c, a, bt = og.method_table(SOLVERS[name])
    assert (ns, order) == {"DP5": (7, 5), "BS3": (4, 3), "RK4": (5, 4)}[name]
    assert np.allclose(a.sum(1), c, atol=1e-15)                       # c_i = sum_j a_ij
    b = a[ns - 1].copy()                                              # FSAL: the last row is b
    assert b[ns - 1] == 0.0
    for lhs, rhs in _conds(b, a, c, order):
        assert abs(lhs - rhs) < 1e-14
    if order < 5:                                                     # and not one order more (conditions listed up to 5)
        assert max(abs(l - r) for l, r in _conds(b, a, c, order + 1)) > 1e-4
    if name != "RK4":
        bhat = b - bt                                                 # btilde = b - bhat, embedded order - 1
        assert abs(bt.sum()) < 1e-15
        for lhs, rhs in _conds(bhat, a, c, order - 1):
            assert abs(lhs - rhs) < 1e-14
        assert max(abs(l - r) for l, r in _conds(bhat, a, c, order)) > 1e-4
    else:
        assert not bt.any()                                           # no embedded pair: fixed step only


@pytest.mark.parametrize("name,dense_order", [("DP5", 4), ("BS3", 3), ("RK4", 3)])
def test_dense_output_order_conditions(name, dense_order):
    sv = SOLVERS[name]
    ns, order, c, a, bt = og.method_table(sv)
    e, ac = np.ones(7), a @ c
    for th in (0.1, 0.35, 0.5, 0.77, 1.0):
        w = og.dense_weights(sv, th)
        want = [(w @ e, th), (w @ c, th ** 2 / 2), (w @ c ** 2, th ** 3 / 3), (w @ ac, th ** 3 / 6)]
        if dense_order >= 4:
            want += [(w @ c ** 3, th ** 4 / 4), (w @ (c * ac), th ** 4 / 8), (w @ (a @ c ** 2), th ** 4 / 12), (w @ (a @ ac), th ** 4 / 24)]
        for lhs, rhs in want:
            assert abs(lhs - rhs) < 1e-13
    assert np.allclose(og.dense_weights(sv, 1.0), a[ns - 1], atol=1e-14)      # continuous at the step end
    assert np.allclose(og.dense_weights(sv, 0.0), 0.0)
    h = 1e-6                                                                   # C1: slope k1 at the left end, k_last at the right
    assert np.allclose(og.dense_weights(sv, h) / h, np.eye(7)[0], atol=1e-5)
    assert np.allclose((og.dense_weights(sv, 1.0) - og.dense_weights(sv, 1.0 - h)) / h, np.eye(7)[ns - 1], atol=1e-5)


def _dop853(z0, L, t, friction=False):
    f = (lambda tt, u: [u[1], -10.0 / L * np.sin(u[0]) - (0.7 * u[1] if friction else 0.0)])
    return solve_ivp(f, (t[0], t[-1]), z0, method="DOP853", rtol=1e-12, atol=1e-12, t_eval=t).y.T


@pytest.mark.parametrize("name,order", [("DP5", 5), ("BS3", 3), ("RK4", 4)])
def test_fixed_step_convergence_order_against_dop853(name, order):
    z0, th = pendulum_inputs(8, dtype="float64")
    t = np.linspace(0.0, 2.4, 49)
    ref = np.stack([_dop853(z0[b], th[b, 0], t) for b in range(8)], 1)
    errs = []
    for dt in (0.1, 0.05, 0.025):          # every save point is a step end at 0.05 / 0.025, every other one is interpolated at 0.1
        tr, ret, na, _ = og.solve(og.PENDULUM, z0, th, t, og.Opts.for_solver(SOLVERS[name], adaptive=False, dt=dt))
        assert (ret == 0).all() and (na == round(2.4 / dt)).all()
        errs.append(np.abs(tr - ref).max())
    slope = np.log2(errs[1] / errs[2])
    assert order - 0.6 < slope < order + 1.5, (errs, slope)


@pytest.mark.parametrize("name", ["DP5", "BS3"])
@pytest.mark.parametrize("rhs", [og.PENDULUM, og.PENDULUM_FRICTION])
def test_adaptive_error_is_within_tolerance_of_dop853(name, rhs):
    z0, th = pendulum_inputs(16, dtype="float64")
    t = 0.05 * np.arange(50)
    ref = np.stack([_dop853(z0[b], th[b, 0], t, rhs == og.PENDULUM_FRICTION) for b in range(16)], 1)
    for tol, bound in ((1e-3, 5e-2), (1e-6, 2e-4), (1e-9, 1e-6)):
        tr, ret, na, nr = og.solve(rhs, z0, th, t, og.Opts.for_solver(SOLVERS[name], abstol=tol, reltol=tol))
        assert (ret == 0).all()
        assert np.abs(tr - ref).max() < bound, (tol, np.abs(tr - ref).max())


def test_rk4_matches_a_literal_python_rk4():
    # the classical method written out, Float64, on the save grid itself (no interpolation)
    z0, th = pendulum_inputs(4, dtype="float64")
    t = 0.05 * np.arange(20)
    tr, ret, na, _ = og.solve(og.PENDULUM, z0, th, t, og.Opts.for_solver(og.RK4, adaptive=False, dt=0.05))
    for b in range(4):
        f = lambda u: np.array([u[1], -10.0 / th[b, 0] * np.sin(u[0])])
        u = z0[b].copy()
        for k in range(1, 20):
            k1 = f(u); k2 = f(u + 0.025 * k1); k3 = f(u + 0.025 * k2); k4 = f(u + 0.05 * k3)
            u = u + 0.05 / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
            assert np.allclose(tr[k, b], u, rtol=0, atol=1e-13)


@pytest.mark.parametrize("name", ["DP5", "BS3", "RK4"])
def test_forward_sensitivity_gradient_matches_finite_differences(name):
    sv = SOLVERS[name]
    B, T = 6, 30
    z0, th = pendulum_inputs(B, dtype="float64")
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(0).standard_normal((T, B, 2))
    o = og.Opts.for_solver(sv, adaptive=False, dt=0.03)       # off-grid steps: the dense output is differentiated too
    dz0, dth = og.grad(og.PENDULUM, z0, th, t, d, o)
    loss = lambda z, p: float((og.solve(og.PENDULUM, z, p, t, o)[0] * d).sum())
    eps = 1e-6
    for b in range(B):
        for i in range(2):
            zp, zm = z0.copy(), z0.copy()
            zp[b, i] += eps; zm[b, i] -= eps
            assert abs((loss(zp, th) - loss(zm, th)) / (2 * eps) - dz0[b, i]) < 1e-6 * max(1.0, abs(dz0[b, i]))
        pp, pm = th.copy(), th.copy()
        pp[b, 0] += eps; pm[b, 0] -= eps
        assert abs((loss(z0, pp) - loss(z0, pm)) / (2 * eps) - dth[b, 0]) < 1e-6 * max(1.0, abs(dth[b, 0]))


def test_controller_defaults_per_algorithm():
    assert og.CONTROLLER_DEFAULTS[og.TSIT5] == (7 / 50, 2 / 25)
    assert og.CONTROLLER_DEFAULTS[og.DP5] == pytest.approx((0.17, 0.04))
    assert og.CONTROLLER_DEFAULTS[og.BS3] == pytest.approx((7 / 30, 2 / 15))
    o = og.Opts.for_solver(og.DP5, reltol=1e-5)
    assert (o.solver, o.beta2, o.reltol) == (og.DP5, 0.04, 1e-5)
