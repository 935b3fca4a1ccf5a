"""Freezes oracle outputs for the seeded inputs of SURVEY.md 8(d) into small .npz fixtures, so that later
refactors of the oracle cannot silently move the target.  The reference itself has no golden vectors (and
cannot be run here: no Julia), so these pin the ORACLE, not the reference.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import pendulum_inputs  # noqa: E402
from oracle import goku as og  # noqa: E402
from oracle import mlp as om  # noqa: E402
from oracle import recurrent as orr  # noqa: E402


def main():
    # C1: GOKU friction-less pendulum, B = 64, T = 50, fp32 state / fp64 time, seed 333; adaptive and fixed step
    z0, th = pendulum_inputs(64, seed=333, dtype="float32")
    t = 0.05 * np.arange(50)
    d = np.random.default_rng(334).standard_normal((50, 64, 2)).astype(np.float32)
    tr_a, ret_a, na_a, nr_a = og.solve(og.PENDULUM, z0, th, t)
    tr_f, ret_f, na_f, nr_f = og.solve(og.PENDULUM, z0, th, t, og.Opts(adaptive=False, dt=0.05))
    gz_f, gp_f = og.grad(og.PENDULUM, z0, th, t, d, og.Opts(adaptive=False, dt=0.05))
    gz_a, gp_a = og.grad(og.PENDULUM, z0, th, t, d, norm_partials=True)
    np.savez_compressed(os.path.join(HERE, "c1_goku_pendulum_f32.npz"), z0=z0, theta=th, t=t, dtraj=d, traj_adaptive=tr_a,
                        naccept_adaptive=na_a, nreject_adaptive=nr_a, traj_fixed=tr_f, naccept_fixed=na_f,
                        dz0_fixed=gz_f, dtheta_fixed=gp_f, dz0_adaptive_fwddiff=gz_a, dtheta_adaptive_fwddiff=gp_a)
    # C3: pendulum with friction, fp64 and fp32 on identical inputs (first 64 of the B = 1024 batch)
    z0, th = pendulum_inputs(1024, seed=333, dtype="float64")
    z0, th = z0[:64], th[:64]
    d = np.random.default_rng(335).standard_normal((50, 64, 2))
    tr64, _, na64, _ = og.solve(og.PENDULUM_FRICTION, z0, th, t)
    tr32, _, na32, _ = og.solve(og.PENDULUM_FRICTION, z0.astype(np.float32), th.astype(np.float32), t)
    gz, gp = og.grad(og.PENDULUM_FRICTION, z0, th, t, d, norm_partials=False)
    # the reference's own semantics (ForwardDiffSensitivity: partials in the error norm), Float64, default tolerance
    rz, rp = og.grad(og.PENDULUM_FRICTION, z0, th, t, d, norm_partials=True)
    np.savez_compressed(os.path.join(HERE, "c3_goku_friction.npz"), z0=z0, theta=th, t=t, dtraj=d, traj_f64=tr64,
                        naccept_f64=na64, traj_f32=tr32, naccept_f32=na32, dz0_f64_frozen=gz, dtheta_f64_frozen=gp,
                        dz0_f64_fwddiff=rz, dtheta_f64_fwddiff=rp)
    # C2: LatentODE, D = 16, H = 200, first 8 of the B = 256 batch solved as a batch of 8 (global norm), seed 1
    rng = np.random.Generator(np.random.PCG64(1))
    dims = [16, 200, 200, 16]
    layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]), np.zeros(dims[i + 1], np.float32)) for i in range(3)]
    p = om.pack_params(layers)
    z0 = (0.5 * rng.standard_normal((256, 16))).astype(np.float32)[:8]
    d = rng.standard_normal((50, 8, 16))
    tr32, na32, nr32, _ = om.solve(z0, p, dims, t)
    tr64, na64, nr64, tape = om.solve(z0.astype(np.float64), p.astype(np.float64), dims, t, og.Opts(adaptive=False, dt=0.05),
                                      record=True)
    gz, gp = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d)
    np.savez_compressed(os.path.join(HERE, "c2_latentode_mlp.npz"), z0=z0, params=p, dims=np.array(dims), t=t, dtraj=d,
                        traj_f32_adaptive_global=tr32, naccept_f32=np.array(na32), traj_f64_fixed=tr64,
                        dz0_f64_fixed=gz, dparams_f64_fixed=gp.astype(np.float32))
    # Other solvers of the diffeq struct: C3's right-hand side, first 32 trajectories, Float64, per solver
    # adaptive (DP5, BS3) and fixed step off the save grid (all three), with the ForwardDiff-semantics gradient
    z0, th = pendulum_inputs(1024, seed=333, dtype="float64")
    z0, th = z0[:32], th[:32]
    d = np.random.default_rng(336).standard_normal((50, 32, 2))
    out = dict(z0=z0, theta=th, t=t, dtraj=d)
    for name, sv in (("dp5", og.DP5), ("bs3", og.BS3), ("rk4", og.RK4)):
        of = og.Opts.for_solver(sv, adaptive=False, dt=0.08)
        out[f"traj_{name}_fixed"] = og.solve(og.PENDULUM_FRICTION, z0, th, t, of)[0]
        out[f"dz0_{name}_fixed"], out[f"dtheta_{name}_fixed"] = og.grad(og.PENDULUM_FRICTION, z0, th, t, d, of)
        if sv != og.RK4:
            oa = og.Opts.for_solver(sv)
            tr, _, na, nr = og.solve(og.PENDULUM_FRICTION, z0, th, t, oa)
            out[f"traj_{name}_adaptive"], out[f"naccept_{name}_adaptive"], out[f"nreject_{name}_adaptive"] = tr, na, nr
            out[f"dz0_{name}_adaptive_fwddiff"], out[f"dtheta_{name}_adaptive_fwddiff"] = og.grad(og.PENDULUM_FRICTION, z0, th, t, d, oa)
    np.savez_compressed(os.path.join(HERE, "solvers_goku_friction_f64.npz"), **out)
    # Recurrent pattern extractor: GOKU's three stacks (F = 32, H = 16) and LatentODE's RNN stack (H = 32),
    # 12 sequences of 20 frames, seed 1; final states and all gradients (float64 restatement of the Flux cells)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((20, 12, 32)).astype(np.float32)
    rnn, lf, lb = orr.init_params(False, 32, rng), orr.init_params(True, 32, rng), orr.init_params(True, 32, rng)
    rnn32 = orr.init_params(False, 32, rng, H=32)
    dz0, dth, dz32 = (rng.standard_normal((12, 16)).astype(np.float32), rng.standard_normal((12, 32)).astype(np.float32),
                      rng.standard_normal((12, 32)).astype(np.float32))
    zo, tho, g = orr.pattern_extractor(x, rnn, lf, lb, dz0, dth)
    zo32, _, g32 = orr.pattern_extractor(x, rnn32, None, None, dz32, H=32)
    np.savez_compressed(os.path.join(HERE, "pattern_extractor.npz"), x=x, rnn=rnn, lstm_f=lf, lstm_b=lb, rnn32=rnn32, dz0=dz0, dtheta=dth,
                        dz0_32=dz32, z0_out=zo, theta_out=tho, dx=g[0], d_rnn=g[1], d_lstm_f=g[2], d_lstm_b=g[3], z0_out_32=zo32,
                        dx_32=g32[0], d_rnn32=g32[1])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
