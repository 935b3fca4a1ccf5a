"""Independent pins of the CPU oracle (CPU tier).

The reference ships no tests, golden vectors or stored outputs (test/runtests.jl:4-6 is an empty test set) and
Julia is not available, so the oracle is PARITY-UNPINNED against the reference itself.  These tests pin it by
means that do not depend on the restatement being right: Runge-Kutta order conditions of the
tableau, convergence order against scipy DOP853 @1e-12, the pendulum energy invariant, finite differences.
"""
import itertools

import numpy as np
import pytest
from scipy.integrate import solve_ivp

from conftest import pendulum_inputs
from oracle import goku as og
from oracle import loss as ol
from oracle import mlp as om

A = om.A
C = np.array(om.CS)
B7 = np.array(A[6] + [0.0])          # b = a7 (FSAL), 7 stages
BT = np.array(om.BT)


def _amat():
    a = np.zeros((7, 7))
    for j in range(1, 7):
        a[j, :j] = A[j]
    return a


def _order_conditions(b, a, c, order):
    """Rooted-tree order conditions up to `order` (Butcher): list of (lhs, rhs)."""
    e = np.ones(7)
    ac = a @ c
    conds = [(b @ e, 1.0)]
    if order >= 2:
        conds += [(b @ c, 1 / 2)]
    if order >= 3:
        conds += [(b @ c ** 2, 1 / 3), (b @ ac, 1 / 6)]
    if order >= 4:
        conds += [(b @ c ** 3, 1 / 4), (b @ (c * ac), 1 / 8), (b @ (a @ c ** 2), 1 / 12), (b @ (a @ ac), 1 / 24)]
    if order >= 5:
        conds += [(b @ c ** 4, 1 / 5), (b @ (c ** 2 * ac), 1 / 10), (b @ (c * (a @ c ** 2)), 1 / 15),
                  (b @ (c * (a @ ac)), 1 / 30), (b @ (ac * ac), 1 / 20), (b @ (a @ c ** 3), 1 / 20),
                  (b @ (a @ (c * ac)), 1 / 40), (b @ (a @ (a @ c ** 2)), 1 / 60), (b @ (a @ (a @ ac)), 1 / 120)]
    return conds


def test_tableau_row_sums_and_order_5():
    a = _amat()
    assert np.allclose(a.sum(1), C, atol=1e-15)               # c_i = sum_j a_ij
    conds = _order_conditions(B7, a, C, 5)
    assert len(conds) == 17
    for lhs, rhs in conds:
        assert abs(lhs - rhs) < 1e-14


def test_embedded_method_is_order_4():
    a = _amat()
    bhat = B7 - BT                                            # btilde = b - bhat
    assert abs(BT.sum()) < 1e-15
    for lhs, rhs in _order_conditions(bhat, a, C, 4):
        assert abs(lhs - rhs) < 1e-13
    # and it is genuinely not order 5 (otherwise the error estimate would vanish)
    assert max(abs(l - r) for l, r in _order_conditions(bhat, a, C, 5)) > 1e-4


def test_dense_output_conditions():
    a = _amat()
    for th in (0.1, 0.35, 0.5, 0.77, 1.0):
        bw = np.array([float(x) for x in om.interp_weights(th)])
        # continuous extension of order 4: sum b_j(th) c_j^k a.. = th^{k+1}/..., for all trees up to order 4
        e, ac = np.ones(7), a @ C
        want = [(bw @ e, th), (bw @ C, th ** 2 / 2), (bw @ C ** 2, th ** 3 / 3), (bw @ ac, th ** 3 / 6),
                (bw @ C ** 3, th ** 4 / 4), (bw @ (C * ac), th ** 4 / 8), (bw @ (a @ C ** 2), th ** 4 / 12),
                (bw @ (a @ ac), th ** 4 / 24)]
        for lhs, rhs in want:
            assert abs(lhs - rhs) < 1e-13
    assert np.allclose([float(x) for x in om.interp_weights(1.0)], B7, atol=1e-14)   # b_j(1) = a7j
    assert np.allclose([float(x) for x in om.interp_weights(0.0)], 0.0)


def _dop853(z0, L, t, friction=False):
    f = (lambda tt, u: [u[1], -10.0 / L * np.sin(u[0]) - (0.7 * u[1] if friction else 0.0)])
    s = solve_ivp(f, (t[0], t[-1]), z0, method="DOP853", rtol=1e-12, atol=1e-12, t_eval=t)
    return s.y.T


def test_fixed_step_convergence_order_against_dop853():
    z0, th = pendulum_inputs(8, dtype="float64")
    t = np.linspace(0.0, 2.4, 49)
    ref = np.stack([_dop853(z0[b], th[b, 0], t) for b in range(8)], 1)
    errs = []
    for dt in (0.2, 0.1, 0.05):
        tr, ret, na, _ = og.solve(og.PENDULUM, z0, th, t, og.Opts(adaptive=False, dt=dt))
        assert (ret == 0).all() and (na == round(2.4 / dt)).all()
        errs.append(np.abs(tr - ref).max())
    slopes = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert (slopes > 4.5).all() and (slopes < 6.5).all(), (errs, slopes)   # global error ~ h^5


@pytest.mark.parametrize("rhs", [og.PENDULUM, og.PENDULUM_FRICTION])
def test_adaptive_error_is_within_tolerance_of_dop853(rhs):
    z0, th = pendulum_inputs(16, dtype="float64")
    t = 0.05 * np.arange(100)
    ref = np.stack([_dop853(z0[b], th[b, 0], t, friction=rhs == og.PENDULUM_FRICTION) for b in range(16)], 1)
    for tol, bound in ((1e-3, 5e-3), (1e-6, 2e-5), (1e-9, 2e-8)):
        tr, ret, na, nr = og.solve(rhs, z0, th, t, og.Opts(abstol=tol * 1e-3, reltol=tol))
        assert (ret == 0).all()
        assert np.abs(tr - ref).max() < bound, (tol, np.abs(tr - ref).max())


def test_energy_invariant_frictionless():
    z0, th = pendulum_inputs(32, dtype="float64")
    t = 0.05 * np.arange(200)
    tr, *_ = og.solve(og.PENDULUM, z0, th, t, og.Opts(abstol=1e-10, reltol=1e-10))
    E = 0.5 * tr[..., 1] ** 2 - 10.0 / th[:, 0] * np.cos(tr[..., 0])
    assert np.abs(E - E[0]).max() < 1e-7
    trf, *_ = og.solve(og.PENDULUM_FRICTION, z0, th, t, og.Opts(abstol=1e-10, reltol=1e-10))
    Ef = 0.5 * trf[..., 1] ** 2 - 10.0 / th[:, 0] * np.cos(trf[..., 0])
    assert (np.diff(Ef, axis=0) <= 1e-9).all()                 # friction only dissipates


def test_first_and_last_save_points_are_exact_and_failure_is_nan_block():
    z0, th = pendulum_inputs(4, dtype="float32")
    t = 0.05 * np.arange(50)
    tr, ret, na, nr = og.solve(og.PENDULUM, z0, th, t)
    assert np.array_equal(tr[0], z0)
    th2 = th.copy()
    th2[1, 0] = 1e-4                                           # G/L = 1e5: cannot finish in 50 iterations
    tr, ret, *_ = og.solve(og.PENDULUM, z0, th2, t, og.Opts(maxiters=50))
    assert ret[1] == og.RET_MAXITERS and np.isnan(tr[:, 1]).all() and np.isfinite(tr[:, [0, 2, 3]]).all()


def test_fastpow_is_the_rough_float32_approximation():
    xs = np.exp(np.random.default_rng(0).uniform(np.log(1e-6), np.log(10.0), 2000))
    for y in (0.14, 0.08):
        rel = np.array([og.fastpow(x, y) / x ** y - 1 for x in xs])
        assert np.abs(rel).max() < 2e-4 and np.abs(rel).max() > 1e-7   # approximate, not exact
    assert og.fastpow(0.0, 0.14) == 0.0


def test_fp32_and_fp64_states_agree():
    z0, th = pendulum_inputs(64, dtype="float64")
    t = 0.05 * np.arange(50)
    a, *_ = og.solve(og.PENDULUM_FRICTION, z0, th, t)
    b, *_ = og.solve(og.PENDULUM_FRICTION, z0.astype(np.float32), th.astype(np.float32), t)
    assert np.abs(a - b).max() < 1e-3 * np.abs(a).max()


def test_forward_sensitivity_gradient_matches_finite_differences():
    z0, th = pendulum_inputs(6, dtype="float64")
    t = 0.05 * np.arange(30)
    d = np.random.default_rng(1).standard_normal((30, 6, 2))
    for o in (og.Opts(adaptive=False, dt=0.05), og.Opts(adaptive=False, dt=0.07)):
        gz, gp = og.grad(og.PENDULUM_FRICTION, z0, th, t, d, o)
        gz2, gp2 = og.grad(og.PENDULUM_FRICTION, z0, th, t, d, o, norm_partials=False)
        assert np.allclose(gz, gz2, rtol=0, atol=1e-13) and np.allclose(gp, gp2, rtol=0, atol=1e-13)
        L = lambda z, p: (og.solve(og.PENDULUM_FRICTION, z, p, t, o)[0] * d).sum(axis=(0, 2))   # noqa: E731
        eps = 1e-6
        for i in range(2):
            zp, zm = z0.copy(), z0.copy()
            zp[:, i] += eps
            zm[:, i] -= eps
            assert np.allclose((L(zp, th) - L(zm, th)) / (2 * eps), gz[:, i], rtol=1e-6, atol=1e-8)
        assert np.allclose((L(z0, th + eps) - L(z0, th - eps)) / (2 * eps), gp[:, 0], rtol=1e-6, atol=1e-8)


def test_reference_gradient_semantics_converge_to_the_exact_derivative():
    # ForwardDiff duals enter the error norm (own step sequence); at tight tolerance both conventions agree
    z0, th = pendulum_inputs(8, dtype="float64")
    t = 0.05 * np.arange(50)
    d = np.random.default_rng(2).standard_normal((50, 8, 2))
    o = og.Opts(abstol=1e-11, reltol=1e-11)
    a = og.grad(og.PENDULUM, z0, th, t, d, o, norm_partials=True)
    b = og.grad(og.PENDULUM, z0, th, t, d, o, norm_partials=False)
    assert np.abs(a[0] - b[0]).max() < 1e-7 * np.abs(b[0]).max() and np.abs(a[1] - b[1]).max() < 1e-7 * np.abs(b[1]).max()
    o = og.Opts()
    a = og.grad(og.PENDULUM, z0, th, t, d, o, norm_partials=True)
    b = og.grad(og.PENDULUM, z0, th, t, d, o, norm_partials=False)
    assert np.abs(a[0] - b[0]).max() < 5e-2 * np.abs(b[0]).max()          # default tolerance: O(reltol) apart


# ---- LatentODE oracle -------------------------------------------------------------------------------------
def _net(seed=1, dims=(16, 200, 200, 16), bias=0.1):
    rng = np.random.Generator(np.random.PCG64(seed))
    layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]).astype(np.float64), bias * rng.standard_normal(dims[i + 1]))
              for i in range(len(dims) - 1)]
    return list(dims), layers, om.pack_params(layers), rng


def test_destructure_order_roundtrip():
    dims, layers, p, _ = _net(dims=(3, 5, 3))
    W0 = layers[0][0]
    assert p[0] == W0[0, 0] and p[1] == W0[1, 0] and p[5] == W0[0, 1]       # vec(W) column-major, W is (out, in)
    assert p.size == om.n_params(dims) == 3 * 5 + 5 + 5 * 3 + 3
    back = om.unpack_params(p, dims)
    for (W, b), (W2, b2) in zip(layers, back):
        assert np.array_equal(W, W2) and np.array_equal(b, b2)


def test_mlp_solve_against_dop853_and_global_vs_per_trajectory():
    dims, layers, p, rng = _net()
    z0 = 0.5 * rng.standard_normal((6, 16))
    t = 0.05 * np.arange(50)
    ref = np.stack([solve_ivp(lambda tt, u: om.mlp(layers, u[None])[0], (0, t[-1]), z0[b], method="DOP853", rtol=1e-12,
                              atol=1e-12, t_eval=t).y.T for b in range(6)], 1)
    trf, naf, _, _ = om.solve(z0, p, dims, t, og.Opts(adaptive=False, dt=0.025))
    assert naf == 98 and np.abs(trf - ref).max() < 1e-5   # relu kinks cap the observed order
    tra, na, nr, _ = om.solve(z0, p, dims, t)
    assert np.abs(tra - ref).max() < 5e-3
    trp, nap, _, _ = om.solve(z0, p, dims, t, norm_mode="per_traj")
    assert np.abs(trp - ref).max() < 5e-3 and len(set(nap.tolist())) > 1     # trajectories take different steps


def test_mlp_discrete_adjoint_matches_finite_differences():
    dims, layers, p, rng = _net(dims=(4, 9, 9, 4))
    z0 = 0.5 * rng.standard_normal((3, 4))
    t = 0.05 * np.arange(12)
    d = rng.standard_normal((12, 3, 4))
    o = og.Opts(adaptive=False, dt=0.07)
    _, _, _, tape = om.solve(z0, p, dims, t, o, record=True)
    gz, gp = om.discrete_adjoint(p, dims, t, tape, d)
    L = lambda z, pp: (om.solve(z, pp, dims, t, o)[0] * d).sum()   # noqa: E731
    eps = 1e-6
    for idx in itertools.product(range(3), range(4)):
        zp, zm = z0.copy(), z0.copy()
        zp[idx] += eps
        zm[idx] -= eps
        assert abs((L(zp, p) - L(zm, p)) / (2 * eps) - gz[idx]) < 1e-6 * max(1, abs(gz[idx]))
    for i in range(0, p.size, 7):
        pp, pm = p.copy(), p.copy()
        pp[i] += eps
        pm[i] -= eps
        assert abs((L(z0, pp) - L(z0, pm)) / (2 * eps) - gp[i]) < 1e-6 * max(1, abs(gp[i]))


# ---- loss / optimiser oracle ----------------------------------------------------------------------------------
def test_loss_oracle_against_direct_formulas():
    rng = np.random.default_rng(0)
    x, xh = rng.random((5, 3, 7), dtype=np.float32), rng.random((5, 3, 7), dtype=np.float32)
    mu = (rng.standard_normal((3, 4)).astype(np.float32), rng.standard_normal((3, 2)).astype(np.float32))
    lv = (rng.standard_normal((3, 4)).astype(np.float32), rng.standard_normal((3, 2)).astype(np.float32))
    tot, rec, k = ol.loss_batch(x, xh, mu, lv, 0.25)
    rec_direct = sum(((x[:, :, pix] - xh[:, :, pix]) ** 2).mean() for pix in range(7))
    k_direct = sum(((np.exp(l) + m ** 2 - l - 1) / 2).sum() / 3 for m, l in zip(mu, lv))
    assert abs(rec - rec_direct) < 1e-6 and abs(k - k_direct) < 1e-5 and abs(tot - (rec_direct + 0.25 * k_direct)) < 1e-5
    dx, dmu, dlv = ol.loss_batch_grads(x, xh, mu, lv, 0.25)
    eps = 1e-3
    xp = xh.copy()
    xp[1, 2, 3] += eps
    assert abs((ol.loss_batch(x, xp, mu, lv, 0.25)[0] - tot) / eps - dx[1, 2, 3]) < 2e-3


def test_adamw_first_step_and_decay_not_scaled_by_lr():
    opt = ol.ADAMW(1e-3, (0.9, 0.999), np.float32(0.001))
    x = np.array([1.0, -2.0, 0.5], dtype=np.float32)
    g = np.array([0.3, -0.1, 0.0], dtype=np.float32)
    x1 = opt.update("w", x.copy(), g)
    # first step: m_hat = g, v_hat = g^2 -> step = lr * g / (|g| + eps); plus decay * x (NOT lr * decay * x)
    want = x - (1e-3 * g / (np.abs(g) + 1e-8) + np.float32(0.001) * x)
    assert np.allclose(x1, want, rtol=1e-6, atol=1e-9)


def test_reference_continuous_adjoint_converges_to_the_discrete_adjoint():
    # Row a13: the reference differentiates LatentODE with the continuous InterpolatingAdjoint; the product computes the
    # discrete adjoint of the accepted steps.  They are two discretisations of the same derivative: the gap closes as the
    # tolerance tightens (slowly, the relu right-hand side is only piecewise smooth).
    dims, layers, p, rng = _net(seed=2, dims=(6, 24, 24, 6))
    z0 = 0.5 * rng.standard_normal((5, 6))
    t = 0.05 * np.arange(20)
    d = rng.standard_normal((20, 5, 6))
    gaps = []
    for tol in (1e-3, 1e-6, 1e-9):
        o = og.Opts(abstol=tol * 1e-3, reltol=tol, controller_pow=1)
        _, _, _, tape = om.solve(z0, p, dims, t, o, record=True)
        gz, gp = om.discrete_adjoint(p, dims, t, tape, d)
        cz, cp = om.interpolating_adjoint(z0, p, dims, t, d, o)
        gaps.append((np.abs(gz - cz).max() / np.abs(gz).max(), np.abs(gp - cp).max() / np.abs(gp).max()))
    assert gaps[0][0] > gaps[1][0] > gaps[2][0] and gaps[0][1] > gaps[1][1] > gaps[2][1]
    assert gaps[0][0] < 5e-2 and gaps[0][1] < 2e-1          # default tolerance: O(reltol^(1/3)) apart
    assert gaps[2][0] < 1e-3 and gaps[2][1] < 5e-3


def test_julia_float32_trig_restatement():
    """Base.sin / Base.cos(::Float32) as restated in the oracle (msun kernels in Float64, rounded once): the published
    bounds of the kernels hold (|sin - s| < 2^-37.5, |cos - c| < 2^-34.1 on [-pi/4, pi/4] before the final rounding),
    the Float32 results are within 0.51 ulp of the truth on the pendulum's range and across the quadrant boundaries of
    rem_pio2_kernel, and they differ from glibc's sinf in about 1 % of the arguments -- which is why the oracle does
    not simply call sinf."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-np.pi / 4, np.pi / 4, 400000), rng.uniform(-10, 10, 400000), rng.uniform(-1e3, 1e3, 100000),
                        rng.uniform(-3e-4, 3e-4, 1000)]).astype(np.float32)
    s, c = og.jl_sincosf(x)
    xd = x.astype(np.float64)
    for got, ref in ((s, np.sin(xd)), (c, np.cos(xd))):
        ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
        e = np.abs(got.astype(np.float64) - ref) / ulp
        assert e.max() <= 0.51, e.max()
    # kernel bounds, evaluated in Float64 exactly as the oracle does
    S = [float.fromhex(h) for h in ("-0x15555554cbac77.0p-55", "0x111110896efbb2.0p-59", "-0x1a00f9e2cae774.0p-65", "0x16cd878c3b46a7.0p-71")]
    C = [float.fromhex(h) for h in ("-0x1ffffffd0c5e81.0p-54", "0x155553e1053a42.0p-57", "-0x16c087e80f1e27.0p-62", "0x199342e0ee5069.0p-68")]
    y = np.linspace(-np.pi / 4, np.pi / 4, 200001)
    z = y * y
    w = z * z
    sk = (y + (z * y) * (S[0] + z * S[1])) + (z * y) * w * (S[2] + z * S[3])
    ck = ((1 + z * C[0]) + w * C[1]) + (w * z) * (C[2] + z * C[3])
    assert np.abs(sk - np.sin(y)).max() < 2.0 ** -37.5 and np.abs(ck - np.cos(y)).max() < 2.0 ** -34.0
    # the first 50000 arguments lie in [-pi/4, pi/4]: the Float32 result is the rounded kernel value
    assert np.array_equal(s[:50000], sk_f32(x[:50000], S))
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.sinf.restype = ctypes.c_float
    libm.sinf.argtypes = [ctypes.c_float]
    sub = x[400000:420000]
    g = np.array([libm.sinf(float(v)) for v in sub], np.float32)
    frac = float((g != s[400000:420000]).mean())
    assert 0.001 < frac < 0.05, frac


def sk_f32(x, S):
    xd = x.astype(np.float64)
    z = xd * xd
    w = z * z
    r = S[2] + z * S[3]
    s = z * xd
    out = ((xd + s * (S[0] + z * S[1])) + s * w * r).astype(np.float32)
    tiny = np.abs(x) < np.float32(0.00034526698)
    out[tiny] = x[tiny]
    return out
