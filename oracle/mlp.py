"""numpy restatement of the LatentODE hot path -- TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows reference ``diffeq_layer(::Decoder{LatentODE}, z0, t)`` (``src/models/LatentODE.jl:61-78``):
``NeuralODE(dudt, (t[1], t[end]), Tsit5(); saveat = t)`` applied to the whole ``(D,B)`` matrix --
ONE integrator, ONE step size for the batch, RMS error norm over all ``D*B`` entries
(DiffEqFlux 1.52 / OrdinaryDiffEq 6.27).  ``dudt`` is the ``NODE`` struct's
``Chain(Dense(D,H,relu), Dense(H,H,relu), Dense(H,D))`` (``examples/pendulum_friction-less/nODE.jl:14-16``)
with parameters in ``Flux.destructure`` order.

Mixed precision as Julia's promotion rules give it for a Float32 state and a Float64 time span
(the out-of-place Tsit5 step): the stage arithmetic is promoted to Float64 by ``dt``; ``u`` and the
``k_j`` are rounded to Float32 when they are stored back into the integrator.

Arrays: the matrix state ``(D,B)`` column-major is a numpy ``[B, D]`` array; trajectories are
``[T, B, D]``.  ``norm_mode='per_traj'`` is the documented deviation (per-trajectory error control)
that the product offers as a performance mode.

PARITY UNPINNED (no reference tests / golden vectors; Julia unavailable): pinned by finite
differences and a scipy DOP853 cross-check in ``tests/``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .goku import Opts, fastpow

C2, C3, C4, C5 = 0.161, 0.327, 0.9, 0.9800255409045097
A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
]
CS = [0.0, C2, C3, C4, C5, 1.0, 1.0]
BT = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
      0.5823571654525552, -0.45808210592918697, 0.015151515151515152]
R = [
    [1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216],
    [0.13169999999999998, -0.2234, 0.1017],
    [3.9302962368947516, -5.941033872131505, 2.490627285651253],
    [-12.411077166933676, 30.33818863028232, -16.548102889244902],
    [37.50931341651104, -88.1789048947664, 47.37952196281928],
    [-27.896526289197286, 65.09189467479366, -34.87065786149661],
    [1.5, -4.0, 2.5],
]


def interp_weights(th):
    """b_1..b_7 of the Tsit5 dense output at Theta.  ``th`` scalar or array."""
    th = np.asarray(th, dtype=np.float64)
    b = [th * (R[0][0] + th * (R[0][1] + th * (R[0][2] + th * R[0][3])))]
    for j in range(1, 7):
        b.append(th * th * (R[j][0] + th * (R[j][1] + th * R[j][2])))
    return b


def glorot_uniform(rng, out_dim, in_dim):
    """Flux.glorot_uniform: U(+-sqrt(6/(in+out))) -- the default ``Dense`` init of nODE.jl:14-16."""
    lim = np.sqrt(6.0 / (in_dim + out_dim))
    return rng.uniform(-lim, lim, (out_dim, in_dim)).astype(np.float32)


def pack_params(layers):
    """[(W(out,in), b(out))...] -> flat vector in Flux.destructure order (vec(W) column-major, then b)."""
    return np.concatenate([np.concatenate([W.T.reshape(-1), b.reshape(-1)]) for W, b in layers])


def unpack_params(p, dims):
    out, off = [], 0
    for i in range(len(dims) - 1):
        fi, fo = dims[i], dims[i + 1]
        W = p[off:off + fi * fo].reshape(fi, fo).T
        off += fi * fo
        b = p[off:off + fo]
        off += fo
        out.append((W, b))
    assert off == p.size
    return out


def n_params(dims):
    return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))


def mlp(layers, U):
    """dudt(U): relu on every layer but the last.  U is [B, D]."""
    h = U
    for i, (W, b) in enumerate(layers):
        h = h @ W.T + b
        if i + 1 < len(layers):
            h = np.maximum(h, 0)
    return h


def mlp_vjp(layers, U, kbar):
    """(d mlp/dU)^T kbar, and [(dW, db)...] = (d mlp/dparams)^T kbar."""
    acts, pre = [U], []
    h = U
    for i, (W, b) in enumerate(layers):
        z = h @ W.T + b
        pre.append(z)
        h = np.maximum(z, 0) if i + 1 < len(layers) else z
        acts.append(h)
    g = kbar
    grads = [None] * len(layers)
    for i in reversed(range(len(layers))):
        if i + 1 < len(layers):
            g = g * (pre[i] > 0)
        W, _ = layers[i]
        grads[i] = (g.T @ acts[i], g.sum(0))
        g = g @ W
    return g, grads


@dataclass
class Tape:
    t: list
    dt: list
    u: list


def _norm(a, mode):
    if mode == "global":
        return np.sqrt(np.mean(a * a))
    return np.sqrt(np.mean(a * a, axis=1))  # per trajectory [B]


def solve(z0, p_flat, dims, t, opts: Opts | None = None, norm_mode="global", record=False):
    """Tsit5 solve of U' = mlp(U) on the [B, D] matrix state, saved on grid ``t``.

    Returns ``(traj[T,B,D], naccept, nreject, tape_or_None)``.  In 'per_traj' mode every
    trajectory carries its own dt (implemented by solving the rows independently).
    """
    opts = opts or Opts()
    S = z0.dtype
    t = np.asarray(t, dtype=np.float64)
    if norm_mode == "per_traj":
        outs, nas, nrs = [], [], []
        for b in range(z0.shape[0]):
            o, na, nr, _ = solve(z0[b:b + 1], p_flat, dims, t, opts, "global")
            outs.append(o)
            nas.append(na)
            nrs.append(nr)
        return np.concatenate(outs, 1), np.array(nas), np.array(nrs), None
    layers64 = [(W.astype(np.float64), b.astype(np.float64)) for W, b in unpack_params(p_flat, dims)]
    f = lambda U: mlp(layers64, U)  # noqa: E731  Float32 weights promoted exactly
    T = t.shape[0]
    t0, tend = t[0], t[-1]
    dtmax = opts.dtmax if opts.dtmax > 0 else tend - t0
    dtmin = opts.dtmin if opts.dtmin > 0 else max(np.finfo(np.float64).eps, np.spacing(abs(t0)))
    abstol, reltol = S.type(opts.abstol), S.type(opts.reltol)
    u = z0.copy()
    k = [None] * 7
    k[0] = f(u.astype(np.float64)).astype(S)
    if opts.adaptive and not opts.dt > 0:
        sk = abstol + np.abs(u) * reltol
        d0 = float(_norm(u / sk, "global"))
        d1 = float(_norm(k[0] / sk, "global"))
        dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        dt0 = min(dt0, dtmax)
        u1 = (u.astype(np.float64) + dt0 * k[0].astype(np.float64))
        f1 = f(u1).astype(S)
        d2 = float(_norm((f1 - k[0]) / sk, "global")) / dt0
        m = max(d1, d2)
        dt1 = max(1e-6, dt0 * 1e-3) if m <= 1e-15 else 10.0 ** (-(2.0 + np.log10(m)) / 5.0)
        dt = max(dtmin, min(100 * dt0, dt1, dtmax))
    else:
        dt = opts.dt
    qold = opts.qoldinit
    traj = np.empty((T,) + z0.shape, dtype=S)
    traj[0] = u
    ks, na, nr, iters = 1, 0, 0, 0
    tcur = t0
    tape = Tape([], [], []) if record else None
    pw = fastpow if opts.controller_pow == 0 else (lambda x, y: x ** y)
    while ks < T:
        if iters >= opts.maxiters:
            traj[:] = np.nan
            break
        iters += 1
        dts = min(dt, tend - tcur)
        tnew = tcur + dts
        if abs(tnew - tend) < 100 * np.spacing(max(abs(tcur), abs(tend))):
            tnew = tend
        u64 = u.astype(np.float64)
        kk = [k[0].astype(np.float64)]
        for j in range(1, 7):
            acc = sum(A[j][i] * kk[i] for i in range(j))
            g = u64 + dts * acc
            if j == 6:
                unew64 = g
            kk.append(f(g))
        unew = unew64.astype(S)
        kS = [x.astype(S) for x in kk]
        accept, q, q11 = True, 1.0, 1.0
        if opts.adaptive:
            utilde = (dts * sum(BT[i] * kk[i] for i in range(7))).astype(S)
            atmp = utilde / (abstol + np.maximum(np.abs(u), np.abs(unew)) * reltol)
            EEst = float(_norm(atmp, "global"))
            if not np.isfinite(EEst) or not np.isfinite(unew).all():
                traj[:] = np.nan
                break
            if EEst == 0:
                q = 1 / opts.qmax
            else:
                q11 = pw(EEst, opts.beta1)
                q = q11 / pw(qold, opts.beta2)
                q = max(1 / opts.qmax, min(1 / opts.qmin, q / opts.gamma))
            accept = EEst <= 1
            if accept:
                if opts.qsteady_min <= q <= opts.qsteady_max:
                    q = 1.0
                qold = max(EEst, opts.qoldinit)
        if accept:
            if record:
                tape.t.append(tcur)
                tape.dt.append(dts)
                tape.u.append(u.copy())
            na += 1
            while ks < T and t[ks] <= tnew:
                if t[ks] == tnew:
                    traj[ks] = unew
                else:
                    bw = interp_weights((t[ks] - tcur) / dts)
                    acc = sum(bw[i] * kS[i].astype(np.float64) for i in range(7))
                    traj[ks] = (u64 + dts * acc).astype(S)
                ks += 1
            u, tcur = unew, tnew
            k[0] = kS[6]
            if opts.adaptive:
                dt = min(dtmax, dts / q)
        else:
            nr += 1
            dt = dts / min(1 / opts.qmin, q11 / opts.gamma)
    return traj, na, nr, tape


def discrete_adjoint(p_flat, dims, t, tape: Tape, dtraj):
    """Exact reverse-mode derivative of the taped steps (float64): returns ``(dz0[B,D], dp_flat)``.

    This is what the product's backward kernel computes.  The reference itself uses the continuous
    ``InterpolatingAdjoint``, which agrees with this only to the solver tolerance.
    """
    layers = [(W.astype(np.float64), b.astype(np.float64)) for W, b in unpack_params(p_flat, dims)]
    t = np.asarray(t, dtype=np.float64)
    T = t.shape[0]
    dtraj = dtraj.astype(np.float64)
    gW = [(np.zeros_like(W), np.zeros_like(b)) for W, b in layers]
    ubn = np.zeros_like(dtraj[0])
    ks = T - 1
    tnext = t[-1]
    for n in reversed(range(len(tape.t))):
        tn, dtn, u = tape.t[n], tape.dt[n], tape.u[n].astype(np.float64)
        g, k = [u], [mlp(layers, u)]
        for j in range(1, 7):
            gj = u + dtn * sum(A[j][i] * k[i] for i in range(j))
            g.append(gj)
            k.append(mlp(layers, gj))
        kbar = [np.zeros_like(u) for _ in range(7)]
        ub = np.zeros_like(u)
        while ks >= 1 and t[ks] > tn:
            if t[ks] == tnext:
                ubn = ubn + dtraj[ks]
            else:
                bw = interp_weights((t[ks] - tn) / dtn)
                for j in range(7):
                    kbar[j] += dtn * bw[j] * dtraj[ks]
                ub += dtraj[ks]
            ks -= 1

        def vjp(j, kb):
            gu, gp = mlp_vjp(layers, g[j], kb)
            for i, (dW, db) in enumerate(gp):
                gW[i][0][...] += dW
                gW[i][1][...] += db
            return gu

        ubn = ubn + vjp(6, kbar[6])
        ub += ubn
        for i in range(6):
            kbar[i] += dtn * A[6][i] * ubn
        for j in (5, 4, 3, 2, 1):
            gb = vjp(j, kbar[j])
            ub += gb
            for i in range(j):
                kbar[i] += dtn * A[j][i] * gb
        ub += vjp(0, kbar[0])
        ubn, tnext = ub, tn
    ubn = ubn + dtraj[0]
    return ubn, pack_params(gW)


# ---- the reference's own sensitivity algorithm for the LatentODE path ------------------------------------------
def _dense_eval(tape_full, tq):
    """u(tq) from the forward solve's dense output (Tsit5 interpolant of the step that contains tq)."""
    ts = tape_full["t"]
    n = min(max(np.searchsorted(ts, tq, side="right") - 1, 0), len(ts) - 2)
    t0, dt = ts[n], tape_full["dt"][n]
    th = (tq - t0) / dt
    bw = interp_weights(th)
    acc = sum(bw[j] * tape_full["k"][n][j] for j in range(7))
    return tape_full["u"][n] + dt * acc


def _forward_dense(z0, p_flat, dims, t, opts):
    """Float64 forward solve (global norm) that keeps every accepted step's (t, dt, u, k1..k7): the `dense = true`
    solution SciMLSensitivity's InterpolatingAdjoint interpolates during the backward sweep."""
    layers = [(W.astype(np.float64), b.astype(np.float64)) for W, b in unpack_params(p_flat, dims)]
    f = lambda U: mlp(layers, U)  # noqa: E731
    t = np.asarray(t, dtype=np.float64)
    t0, tend = t[0], t[-1]
    u = z0.astype(np.float64).copy()
    k1 = f(u)
    sk = opts.abstol + np.abs(u) * opts.reltol
    d0, d1 = _norm(u / sk, "global"), _norm(k1 / sk, "global")
    dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = f(u + dt0 * k1)
    d2 = _norm((f1 - k1) / sk, "global") / dt0
    dt = min(100 * dt0, 10.0 ** (-(2.0 + np.log10(max(d1, d2))) / 5.0), tend - t0) if opts.adaptive and not opts.dt > 0 else opts.dt
    rec = {"t": [], "dt": [], "u": [], "k": []}
    tc, qold = t0, opts.qoldinit
    while tc < tend:
        dts = min(dt, tend - tc)
        kk = [k1]
        for j in range(1, 7):
            g = u + dts * sum(A[j][i] * kk[i] for i in range(j))
            kk.append(f(g))
        unew = g
        accept, q, q11 = True, 1.0, 1.0
        if opts.adaptive:
            ut = dts * sum(BT[i] * kk[i] for i in range(7))
            EEst = float(_norm(ut / (opts.abstol + np.maximum(np.abs(u), np.abs(unew)) * opts.reltol), "global"))
            q11 = EEst ** opts.beta1 if EEst > 0 else 1.0
            q = max(1 / opts.qmax, min(1 / opts.qmin, (q11 / qold ** opts.beta2) / opts.gamma)) if EEst > 0 else 1 / opts.qmax
            accept = EEst <= 1
        if accept:
            rec["t"].append(tc); rec["dt"].append(dts); rec["u"].append(u); rec["k"].append(kk)
            if opts.adaptive:
                qold = max(EEst, opts.qoldinit)
                dt = min(tend - t0, dts / q)
            tc = tend if abs(tc + dts - tend) < 1e-13 else tc + dts
            u, k1 = unew, kk[6]
        else:
            dt = dts / min(1 / opts.qmin, q11 / opts.gamma)
    rec["t"].append(tend)
    rec["t"] = np.array(rec["t"])
    return rec, layers


def _dense_from_tape(tape: Tape, layers):
    """The dense forward solution rebuilt from a step tape (t_n, dt_n, u_n): the stage derivatives k_1..k_7 of every
    accepted step are recomputed from u_n (what the CUDA kernel does: the tape holds no stages)."""
    f = lambda U: mlp(layers, U)  # noqa: E731
    rec = {"t": [], "dt": [], "u": [], "k": []}
    for tn, dtn, un in zip(tape.t, tape.dt, tape.u):
        u = un.astype(np.float64)
        kk = [f(u)]
        for j in range(1, 7):
            kk.append(f(u + dtn * sum(A[j][i] * kk[i] for i in range(j))))
        rec["t"].append(tn); rec["dt"].append(dtn); rec["u"].append(u); rec["k"].append(kk)
    rec["t"].append(tape.t[-1] + tape.dt[-1])
    rec["t"] = np.array(rec["t"])
    return rec


def interpolating_adjoint(z0, p_flat, dims, t, dtraj, opts: Opts | None = None, tape: Tape | None = None, stats=None):
    """Gradients the way the REFERENCE computes them for LatentODE: DiffEqFlux's NeuralODE default
    ``InterpolatingAdjoint(autojacvec = ZygoteVJP())`` [3P SciMLSensitivity 7.10, restated from the published algorithm]
    -- the continuous adjoint
        lambda' = -(df/du)^T lambda,   mu' = -(df/dp)^T lambda
    solved as ONE ODE on the augmented state [lambda (D*B); mu (n_params)] from t_end back to t_0 with adaptive Tsit5 at
    the solve's abstol / reltol: RMS error norm over the whole augmented vector, Hairer initial step, PI controller (the
    same OrdinaryDiffEq machinery as the forward solve, ``solve`` above), u(t) from the forward solution's dense output,
    and a ``PresetTimeCallback`` at every save time: the step is clipped to the next save time (tstop), there
    ``lambda += dtraj[k]`` and the FSAL derivative is re-evaluated.  Float64 throughout.

    ``tape``: accepted steps of the forward solve to interpolate (``solve(..., record=True)``); without it a Float64
    forward solve is made here.  Returns ``(dz0[B,D], dparams_flat)``; ``stats`` (a dict) receives naccept / nreject.
    The product offers this as LDEQ_SENSE_INTERPOLATING_ADJOINT next to the discrete adjoint of the accepted steps; the
    two agree to the solver tolerance, and tests quantify by how much."""
    opts = opts or Opts()
    t = np.asarray(t, dtype=np.float64)
    if tape is None:
        rec, layers = _forward_dense(z0, p_flat, dims, t, opts)
    else:
        layers = [(W.astype(np.float64), b.astype(np.float64)) for W, b in unpack_params(p_flat, dims)]
        rec = _dense_from_tape(tape, layers)
    T = len(t)
    npar = n_params(dims)
    lam = np.asarray(dtraj[T - 1], dtype=np.float64).copy()
    mu = np.zeros(npar)
    nall = lam.size + npar

    def rhs(tq, lam_):
        gu, gp = mlp_vjp(layers, _dense_eval(rec, tq), lam_)
        return -gu, -pack_params(gp).astype(np.float64)

    def rms(a_l, a_m):
        return float(np.sqrt((np.sum(a_l * a_l) + np.sum(a_m * a_m)) / nall))

    t0, tend = t[0], t[-1]
    dtmax = opts.dtmax if opts.dtmax > 0 else tend - t0
    dtmin = opts.dtmin if opts.dtmin > 0 else max(np.finfo(np.float64).eps, np.spacing(abs(tend)))
    pw = fastpow if opts.controller_pow == 0 else (lambda x, y: x ** y)
    tc, ks = tend, T - 2
    kl, km = rhs(tc, lam)
    if opts.adaptive and not opts.dt > 0:
        # ode_determine_initdt on the augmented state, time running backwards (tdir = -1)
        sl, sm = opts.abstol + np.abs(lam) * opts.reltol, opts.abstol + np.abs(mu) * opts.reltol
        d0, d1 = rms(lam / sl, mu / sm), rms(kl / sl, km / sm)
        dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        dt0 = min(dt0, dtmax)
        fl, fm = rhs(tc - dt0, lam - dt0 * kl)
        d2 = rms((fl - kl) / sl, (fm - km) / sm) / dt0
        m = max(d1, d2)
        dt1 = max(1e-6, dt0 * 1e-3) if m <= 1e-15 else 10.0 ** (-(2.0 + np.log10(m)) / 5.0)
        dt = max(dtmin, min(100 * dt0, dt1, dtmax))
    else:
        dt = opts.dt
    qold = opts.qoldinit
    na = nr = iters = 0
    while ks >= 0:
        if iters >= opts.maxiters:
            lam[:], mu[:] = np.nan, np.nan
            break
        iters += 1
        tstop = t[ks]
        dts = min(dt, tc - tstop)
        tnew = tc - dts
        if abs(tnew - tstop) < 100 * np.spacing(max(abs(tc), abs(tstop))):
            tnew = tstop
        kls, kms = [kl], [km]
        for j in range(1, 7):
            gl = lam - dts * sum(A[j][i] * kls[i] for i in range(j))
            a, b_ = rhs(tc - CS[j] * dts, gl)
            kls.append(a); kms.append(b_)
        new_l = gl
        new_m = mu - dts * sum(A[6][i] * kms[i] for i in range(6))
        accept, q, q11 = True, 1.0, 1.0
        if opts.adaptive:
            el = dts * sum(BT[i] * kls[i] for i in range(7))
            em = dts * sum(BT[i] * kms[i] for i in range(7))
            EEst = rms(el / (opts.abstol + np.maximum(np.abs(lam), np.abs(new_l)) * opts.reltol),
                       em / (opts.abstol + np.maximum(np.abs(mu), np.abs(new_m)) * opts.reltol))
            if not np.isfinite(EEst):
                lam[:], mu[:] = np.nan, np.nan
                break
            if EEst == 0:
                q = 1 / opts.qmax
            else:
                q11 = pw(EEst, opts.beta1)
                q = max(1 / opts.qmax, min(1 / opts.qmin, (q11 / pw(qold, opts.beta2)) / opts.gamma))
            accept = EEst <= 1
            if stats is not None:
                stats.setdefault("trace", []).append((tc, dts, EEst, float(accept)))
            if accept:
                if opts.qsteady_min <= q <= opts.qsteady_max:
                    q = 1.0
                qold = max(EEst, opts.qoldinit)
        if accept:
            na += 1
            tc, lam, mu = tnew, new_l, new_m
            kl, km = kls[6], kms[6]
            if tc == tstop:
                lam = lam + dtraj[ks]      # the callback: cotangent of save point ks, then f is re-evaluated
                ks -= 1
                if ks >= 0:
                    kl, km = rhs(tc, lam)
            if opts.adaptive:
                dt = min(dtmax, dts / q)
        else:
            nr += 1
            dt = dts / min(1 / opts.qmin, q11 / opts.gamma)
        if ks >= 0 and opts.adaptive and not (abs(dt) > dtmin and np.isfinite(dt)):
            lam[:], mu[:] = np.nan, np.nan
            break
    if stats is not None:
        stats["naccept"], stats["nreject"] = na, nr
    return lam, mu
