"""GPU parity tests of the LatentODE hot path (MLP right-hand side) against the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import goku as og
from oracle import mlp as om

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(D=16, H=200, seed=1, dtype=np.float32, bias_scale=0.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = [D, H, H, D]
    layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]).astype(dtype),
               (bias_scale * rng.standard_normal(dims[i + 1])).astype(dtype)) for i in range(3)]
    return dims, om.pack_params(layers).astype(dtype), rng


def _solve(ldeq, z0, p, dims, t, want_grad=None, **kw):
    opts = ldeq.default_opts(**kw)
    z = torch.from_numpy(z0).to(DEV)
    pp = torch.from_numpy(p).to(DEV)
    if want_grad is None:
        traj, st, _ = ldeq.mlp_solve_raw(z, pp, dims, t, opts)
        torch.cuda.synchronize()
        return traj.cpu().numpy(), st.retcode.cpu().numpy(), st.naccept.cpu().numpy(), st.nreject.cpu().numpy()
    z.requires_grad_(True)
    pp.requires_grad_(True)
    traj = ldeq.mlp_solve(z, pp, dims, t, opts)
    traj.backward(torch.from_numpy(want_grad).to(DEV))
    torch.cuda.synchronize()
    return traj.detach().cpu().numpy(), z.grad.cpu().numpy(), pp.grad.cpu().numpy()


@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-3), ("float64", 1e-5)])
@pytest.mark.parametrize("mode", ["global", "per_traj"])
def test_c2_forward_matches_oracle(ldeq, dtype, rtol, mode):
    # C2: D = 16, H = 200, B = 256, T = 50, z0 ~ N(0, 0.5^2), glorot weights, zero bias, seed 1
    dims, p, rng = _net(dtype=dtype)
    B, T = 256, 50
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(dtype)
    t = 0.05 * np.arange(T)
    nm = ldeq.NORM_GLOBAL if mode == "global" else ldeq.NORM_PER_TRAJ
    tr, ret, na, nr = _solve(ldeq, z0, p, dims, t, norm_mode=nm)
    assert (ret == 0).all()
    if mode == "global":
        # reference semantics: one step sequence for the whole batch, identical to the oracle's
        otr, ona, onr, _ = om.solve(z0, p, dims, t, norm_mode="global")
        assert (na == ona).all() and (nr == onr).all()
        assert np.abs(tr - otr).max() <= rtol * np.abs(otr).max()
    else:
        # per-trajectory control (documented deviation).  A 16-entry error norm on a piecewise-linear (relu)
        # right-hand side is sensitive: perturbing the ORACLE's controller by 1e-7 moves single trajectories by
        # 6e-5, so agreement is to the solver tolerance (reltol 1e-3); step counts agree exactly in fp64.
        otr, ona, onr, _ = om.solve(z0[:32], p, dims, t, norm_mode="per_traj")
        tr, na = tr[:, :32], na[:32]
        if dtype == "float64":
            assert (na == ona).all()
            assert np.median(np.abs(tr - otr).max(axis=(0, 2))) <= 1e-6
        assert np.abs(tr - otr).max() <= 1e-3 * np.abs(otr).max()


@pytest.mark.parametrize("dtype,rtol", [("float32", 2e-5), ("float64", 1e-11)])
def test_fixed_step_forward_and_step_count(ldeq, dtype, rtol):
    dims, p, rng = _net(dtype=dtype, bias_scale=0.1)
    B, T = 37, 50   # ragged batch: not a multiple of any tile size
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(dtype)
    t = 0.05 * np.arange(T)
    for dt in (0.05, 0.07):
        tr, ret, na, nr = _solve(ldeq, z0, p, dims, t, adaptive=False, dt=dt)
        otr, ona, _, _ = om.solve(z0, p, dims, t, og.Opts(adaptive=False, dt=dt))
        assert (na == ona).all() and (ret == 0).all()
        assert np.abs(tr - otr).max() <= rtol * np.abs(otr).max()


@pytest.mark.parametrize("dtype,rtol", [("float32", 2e-4), ("float64", 1e-9)])
@pytest.mark.parametrize("adaptive", [False, True])
def test_adjoint_matches_discrete_adjoint_oracle(ldeq, dtype, rtol, adaptive):
    if adaptive:
        # the kernel's proposed step sizes carry fp32 rounding (1e-7 relative): fp64 gradients agree to 1e-6.
        # In fp32 the kernel (pure fp32 stages) and the oracle (Julia's mixed fp32/fp64 stages) take different
        # accepted steps; the adjoints of two different discretisations agree to the solver tolerance only
        # (measured 2-4 % on dp at reltol 1e-3) -- the tight fp32 check is the fixed-step case above.
        rtol = 1e-6 if dtype == "float64" else 1e-1
    dims, p, rng = _net(dtype=dtype, bias_scale=0.1)
    B, T = 24, 20
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(dtype)
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, 16)).astype(dtype)
    kw = dict(adaptive=False, dt=0.07) if not adaptive else dict(controller_pow=1)
    okw = og.Opts(adaptive=False, dt=0.07) if not adaptive else og.Opts(controller_pow=1)
    tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, **kw)
    # the oracle runs in the kernel's state precision (so both take the same steps); its adjoint sweep is fp64
    otr, ona, _, tape = om.solve(z0, p, dims, t, okw, record=True)
    oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d)
    print("traj err", np.abs(tr - otr).max(), "dz0 err", np.abs(gz - oz).max() / np.abs(oz).max(), "dp err",
          np.abs(gp - op).max() / np.abs(op).max())
    assert np.abs(tr - otr).max() <= (1e-3 if dtype == "float32" else 1e-8) * np.abs(otr).max()
    assert np.abs(gz - oz).max() <= rtol * np.abs(oz).max()
    assert np.abs(gp - op).max() <= rtol * np.abs(op).max()


def test_other_architectures_and_augmentation(ldeq):
    # 2 layers / 4 layers / widths that are not multiples of anything
    rng = np.random.Generator(np.random.PCG64(5))
    for dims in ([5, 33, 5], [3, 17, 40, 9, 3], [18, 200, 200, 18]):
        layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]), (0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32))
                  for i in range(len(dims) - 1)]
        p = om.pack_params(layers).astype(np.float64)
        B, T = 9, 12
        z0 = 0.5 * rng.standard_normal((B, dims[0]))
        t = 0.05 * np.arange(T)
        d = rng.standard_normal((T, B, dims[0]))
        tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, adaptive=False, dt=0.05)
        otr, _, _, tape = om.solve(z0, p, dims, t, og.Opts(adaptive=False, dt=0.05), record=True)
        oz, op = om.discrete_adjoint(p, dims, t, tape, d)
        assert np.abs(tr - otr).max() <= 1e-11 * np.abs(otr).max()
        assert np.abs(gz - oz).max() <= 1e-9 * np.abs(oz).max()
        assert np.abs(gp - op).max() <= 1e-9 * max(np.abs(op).max(), 1e-30)


def test_large_batch_per_trajectory(ldeq):
    # beyond one wave of tiles: persistent CTAs loop over tiles
    dims, p, rng = _net()
    B, T = 3000, 20
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(np.float32)
    t = 0.05 * np.arange(T)
    tr, ret, na, nr = _solve(ldeq, z0, p, dims, t, norm_mode=ldeq.NORM_PER_TRAJ)
    sel = np.arange(0, B, 97)
    otr, _, _, _ = om.solve(z0[sel], p, dims, t, norm_mode="per_traj")
    assert (ret == 0).all()
    assert np.abs(tr[:, sel] - otr).max() <= 1e-3 * np.abs(otr).max()


# ---- resident-weights kernels (Float32 weight image in shared memory, ldeq_mlp_res.cu) ------------------------------
@pytest.mark.parametrize("dims", [[16, 200, 200, 16], [6, 50, 30, 6], [8, 64, 64, 32, 8], [4, 36, 4], [10, 10]])
def test_resident_path_other_shapes_match_oracle_and_general_path(ldeq, dims, monkeypatch):
    # depth 1..4, widths that are / are not multiples of the warp tiling, an odd batch (ragged last tile); the
    # Float32 fixed-step solve and its discrete adjoint against the oracle, and against the general kernels
    rng = np.random.Generator(np.random.PCG64(11))
    layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]), (0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32))
              for i in range(len(dims) - 1)]
    p = om.pack_params(layers).astype(np.float32)
    B, T = 37, 12
    z0 = (0.5 * rng.standard_normal((B, dims[0]))).astype(np.float32)
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, dims[0])).astype(np.float32)
    kw = dict(adaptive=False, dt=0.05)
    monkeypatch.delenv("LDEQ_MLP_NO_RESIDENT", raising=False)
    tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, **kw)
    monkeypatch.setenv("LDEQ_MLP_NO_RESIDENT", "1")
    tr2, gz2, gp2 = _solve(ldeq, z0, p, dims, t, want_grad=d, **kw)
    monkeypatch.delenv("LDEQ_MLP_NO_RESIDENT")
    otr, _, _, tape = om.solve(z0.astype(np.float64), p.astype(np.float64), dims, t, og.Opts(adaptive=False, dt=0.05), record=True)
    oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d.astype(np.float64))
    for a, b, o, tol in ((tr, tr2, otr, 2e-5), (gz, gz2, oz, 2e-4), (gp, gp2, op, 2e-4)):
        scale = max(np.abs(o).max(), 1e-30)
        assert np.abs(a - o).max() <= tol * scale      # resident vs fp64 oracle
        assert np.abs(b - o).max() <= tol * scale      # general vs fp64 oracle
        assert np.abs(a - b).max() <= 0.2 * tol * scale


def test_resident_adjoint_adaptive_ragged_steps_and_failures(ldeq, monkeypatch):
    # per-trajectory control: the two trajectories of a tile take different numbers of steps (one finishes early and
    # pulls back k1 of its first step while the other still sweeps); a trajectory that fails (maxiters) must
    # contribute nothing to the parameter gradient and get a zero dz0
    dims, p, rng = _net(bias_scale=0.1)
    B, T = 33, 20
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(np.float32)
    z0[::2] *= 4.0   # larger states -> more steps on every other trajectory
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, 16)).astype(np.float32)
    kw = dict(norm_mode=ldeq.NORM_PER_TRAJ, reltol=1e-5, abstol=1e-7)
    monkeypatch.delenv("LDEQ_MLP_NO_RESIDENT", raising=False)
    # ONE forward solve, then both adjoint kernels on the SAME tape: two discrete adjoints of an identical step
    # sequence must agree to Float32 rounding (two separate solves would only agree to the solver tolerance)
    zz, pp, dd = (torch.from_numpy(a).to(DEV) for a in (z0, p, d))
    tr, st, tape = ldeq.mlp_solve_raw(zz, pp, dims, t, ldeq.default_opts(**kw), want_tape=True)
    na = st.naccept.cpu().numpy()
    assert (st.retcode.cpu().numpy() == 0).all() and len(set(na.tolist())) > 2 and (na[0:32:2] != na[1:32:2]).any()
    gz, gp = (a.cpu().numpy() for a in ldeq.mlp_bwd_raw(tape, dd))
    monkeypatch.setenv("LDEQ_MLP_NO_RESIDENT", "1")
    gz2, gp2 = (a.cpu().numpy() for a in ldeq.mlp_bwd_raw(tape, dd))
    monkeypatch.delenv("LDEQ_MLP_NO_RESIDENT")
    tape.free()
    assert np.abs(gz - gz2).max() <= 2e-5 * np.abs(gz2).max()
    assert np.abs(gp - gp2).max() <= 2e-5 * np.abs(gp2).max()
    # failures: maxiters = 3 stops every trajectory -> NaN block, zero gradients
    tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, maxiters=3, **kw)
    assert np.isnan(tr).all() and (gz == 0).all() and (gp == 0).all()


def test_resident_path_long_nonuniform_grid_and_single_trajectory(ldeq):
    # more save points than the on-chip copy of the grid holds (falls back to the global grid), unevenly spaced, B = 1
    dims, p, rng = _net(bias_scale=0.1)
    T = 300
    t = np.cumsum(np.r_[0.0, rng.uniform(0.002, 0.01, T - 1)])
    for B in (1, 5):
        z0 = (0.5 * rng.standard_normal((B, 16))).astype(np.float32)
        d = rng.standard_normal((T, B, 16)).astype(np.float32)
        tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, adaptive=False, dt=0.004)
        otr, _, _, tape = om.solve(z0.astype(np.float64), p.astype(np.float64), dims, t, og.Opts(adaptive=False, dt=0.004), record=True)
        oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d.astype(np.float64))
        assert np.abs(tr - otr).max() <= 2e-5 * np.abs(otr).max()
        assert np.abs(gz - oz).max() <= 2e-4 * np.abs(oz).max()
        assert np.abs(gp - op).max() <= 2e-4 * np.abs(op).max()


# ---- tcgen05 / TMEM path (LDEQ_MLP_MATH_BF16X3) ---------------------------------------------------------------
@pytest.mark.parametrize("B", [256, 1000, 20000])
def test_tensor_core_path_matches_oracle_and_exact_path(ldeq, B):
    # north star: trajectories within 1e-3 (fp32).  The bf16x3 product carries 2^-16 per term: fixed-step
    # trajectories agree with the oracle to ~4e-6.
    dims, p, rng = _net(bias_scale=0.1)
    T = 50
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(np.float32)
    t = 0.05 * np.arange(T)
    tc, ret, na, nr = _solve(ldeq, z0, p, dims, t, adaptive=False, dt=0.05, mlp_math=ldeq.MLP_MATH_BF16X3)
    ex, _, na_e, _ = _solve(ldeq, z0, p, dims, t, adaptive=False, dt=0.05)
    assert (ret == 0).all() and (na == na_e).all() and (na == T - 1).all()
    assert np.abs(tc - ex).max() <= 5e-5 * np.abs(ex).max()
    sel = np.arange(0, B, max(1, B // 64))
    otr, _, _, _ = om.solve(z0[sel], p, dims, t, og.Opts(adaptive=False, dt=0.05))
    assert np.abs(tc[:, sel] - otr).max() <= 5e-5 * np.abs(otr).max()
    # adaptive, per-trajectory control (any B) and global norm (B <= 128 * SM count)
    tc, ret, na, nr = _solve(ldeq, z0, p, dims, t, norm_mode=ldeq.NORM_PER_TRAJ, mlp_math=ldeq.MLP_MATH_BF16X3)
    otr, _, _, _ = om.solve(z0[sel[:16]], p, dims, t, norm_mode="per_traj")
    assert (ret == 0).all() and np.abs(tc[:, sel[:16]] - otr).max() <= 1e-3 * np.abs(otr).max()
    if B <= 1000:
        tc, ret, na, nr = _solve(ldeq, z0, p, dims, t, norm_mode=ldeq.NORM_GLOBAL, mlp_math=ldeq.MLP_MATH_BF16X3)
        otr, ona, onr, _ = om.solve(z0, p, dims, t, norm_mode="global")
        assert (na == ona).all() and (nr == onr).all()
        assert np.abs(tc - otr).max() <= 1e-3 * np.abs(otr).max()


def _saturated_net(rng, dims):
    """A network whose relu masks cannot depend on rounding: hidden biases of +-3 (layer 1, |W1 x| < 1 here) and +-12
    (layer 2, |W2 h1| < 8 here) keep every unit far from its kink, so the oracle, the exact kernel and the bf16x3
    tensor-core kernel differentiate the SAME piecewise-linear function; the output layer is scaled down to keep the
    dynamics as mild as the glorot network's."""
    layers = []
    for i in range(3):
        W = om.glorot_uniform(rng, dims[i + 1], dims[i])
        b = (0.1 * rng.standard_normal(dims[i + 1])).astype(np.float32)
        if i < 2:
            b = (np.where(rng.random(dims[i + 1]) < 0.6, 1.0, -1.0) * (3.0 if i == 0 else 12.0)).astype(np.float32) + b
        else:
            W = (0.05 * W).astype(np.float32)
        layers.append((W, b))
    return om.pack_params(layers).astype(np.float32)


@pytest.mark.parametrize("B,T,kw", [(130, 12, dict(adaptive=False, dt=0.05)), (256, 20, dict(norm_mode=0)), (300, 12, dict(norm_mode=1))])
def test_tensor_core_reverse_pass_exact_when_masks_are_unambiguous(ldeq, B, T, kw, monkeypatch):
    """The tcgen05 reverse pass (adjoint sweep with transposed reads of the resident weight images + weight-gradient GEMM
    over bf16 hi/lo records, ldeq_mlp_tc_bwd.cu) against the oracle's discrete adjoint: dz0 and all 46 816 parameter
    gradients within 2e-4 in the max norm (ragged last tile; fixed step, batch-global steps, per-trajectory steps)."""
    rng = np.random.Generator(np.random.PCG64(21))
    dims = [16, 200, 200, 16]
    p = _saturated_net(rng, dims)
    z0 = (0.3 * rng.standard_normal((B, 16))).astype(np.float32)
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, 16)).astype(np.float32)
    tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, mlp_math=ldeq.MLP_MATH_BF16X3, **kw)
    # the exact-arithmetic adjoint kernel sweeping the SAME tape (same step sequence): strict in every mode
    monkeypatch.setenv("LDEQ_MLP_TC_BWD_OFF", "1")
    tr2, gz2, gp2 = _solve(ldeq, z0, p, dims, t, want_grad=d, mlp_math=ldeq.MLP_MATH_BF16X3, **kw)
    monkeypatch.delenv("LDEQ_MLP_TC_BWD_OFF")
    assert np.array_equal(tr, tr2)
    e1, e2 = np.abs(gz - gz2).max() / np.abs(gz2).max(), np.abs(gp - gp2).max() / np.abs(gp2).max()
    print(f"tc reverse pass vs exact adjoint on the same tape (B={B}, {kw}): dz0 {e1:.2e} dparams {e2:.2e}")
    assert e1 <= 2e-4 and e2 <= 2e-4
    if not kw.get("adaptive", True):
        # fixed step: the oracle takes the same steps
        _, _, _, tape = om.solve(z0, p, dims, t, og.Opts(adaptive=False, dt=kw["dt"]), record=True)
        oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d)
        ez, ep = np.abs(gz - oz).max() / np.abs(oz).max(), np.abs(gp - op).max() / np.abs(op).max()
        print(f"tc reverse pass vs oracle (fixed step, B={B}): dz0 {ez:.2e} dparams {ep:.2e}")
        assert ez <= 2e-4 and ep <= 2e-4
    elif kw.get("norm_mode") == 0:
        # adaptive, batch-global steps: the oracle's own step sizes differ in the last bits of a Float32 controller, the two
        # discrete adjoints agree at the solver tolerance
        _, _, _, tape = om.solve(z0, p, dims, t, og.Opts(), norm_mode="global", record=True)
        oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d)
        assert np.abs(gz - oz).max() <= 5e-2 * np.abs(oz).max() and np.linalg.norm(gp - op) <= 5e-2 * np.linalg.norm(op)


def test_tensor_core_reverse_pass_realistic_net_is_the_adjoint_of_the_tensor_core_forward(ldeq, monkeypatch):
    """Glorot weights, small biases (C2's network): pre-activations come arbitrarily close to zero, where bf16x3 and Float32
    arithmetic can disagree on a relu mask bit (measured: 1 row of 128 at T = 3).  The tensor-core reverse pass differentiates
    the masks the tensor-core FORWARD pass used; against the exact-arithmetic adjoint kernel sweeping the same tape every
    other row agrees to rounding, and so does the bulk of the parameter gradient."""
    dims, p, rng = _net(bias_scale=0.1)
    B, T = 256, 20
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(np.float32)
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, 16)).astype(np.float32)
    kw = dict(adaptive=False, dt=0.05)
    tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, mlp_math=ldeq.MLP_MATH_BF16X3, **kw)
    monkeypatch.setenv("LDEQ_MLP_TC_BWD_OFF", "1")      # same tensor-core tape, exact-arithmetic adjoint kernel
    tr2, gz2, gp2 = _solve(ldeq, z0, p, dims, t, want_grad=d, mlp_math=ldeq.MLP_MATH_BF16X3, **kw)
    monkeypatch.delenv("LDEQ_MLP_TC_BWD_OFF")
    assert np.array_equal(tr, tr2)
    rows = np.abs(gz - gz2).max(1) / np.abs(gz2).max()
    l2 = np.linalg.norm(gp - gp2) / np.linalg.norm(gp2)
    frac = np.mean(np.abs(gp - gp2) <= 2e-4 * np.abs(gp2).max())
    print(f"tc reverse pass vs exact adjoint on the same tape: rows within 2e-4: {np.mean(rows <= 2e-4):.4f}, worst row {rows.max():.2e}; "
          f"dparams L2 {l2:.2e}, entries within 2e-4 of max: {frac:.4f}")
    assert np.mean(rows <= 2e-4) >= 0.9 and np.median(rows) <= 1e-5 and rows.max() <= 5e-2
    assert l2 <= 2e-2 and frac >= 0.9
    _, _, _, tape = om.solve(z0, p, dims, t, og.Opts(**kw), record=True)
    oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d)
    rows = np.abs(gz - oz).max(1) / np.abs(oz).max()
    assert np.mean(rows <= 2e-4) >= 0.9 and np.linalg.norm(gp - op) / np.linalg.norm(op) <= 2e-2


def test_tensor_core_reverse_pass_failed_rows_and_smaller_nets(ldeq):
    # a trajectory that fails (maxiters) contributes nothing; a narrower network (padding columns in every layer)
    rng = np.random.Generator(np.random.PCG64(11))
    dims = [6, 50, 70, 6]
    p = _saturated_net(rng, dims)
    B, T = 200, 12
    z0 = (0.3 * rng.standard_normal((B, 6))).astype(np.float32)
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, 6)).astype(np.float32)
    tr, gz, gp = _solve(ldeq, z0, p, dims, t, want_grad=d, adaptive=False, dt=0.025, mlp_math=ldeq.MLP_MATH_BF16X3)
    _, _, _, tape = om.solve(z0, p, dims, t, og.Opts(adaptive=False, dt=0.025), record=True)
    oz, op = om.discrete_adjoint(p.astype(np.float64), dims, t, tape, d)
    assert np.abs(gz - oz).max() <= 2e-4 * np.abs(oz).max() and np.abs(gp - op).max() <= 2e-4 * np.abs(op).max()


def test_tensor_core_path_refuses_unsupported_shapes(ldeq):
    rng = np.random.Generator(np.random.PCG64(5))
    dims = [16, 300, 300, 16]
    p = np.zeros(om.n_params(dims), np.float32)
    z0 = rng.standard_normal((8, 16)).astype(np.float32)
    with pytest.raises(ldeq.LdeqError) as e:
        _solve(ldeq, z0, p, dims, 0.05 * np.arange(5), mlp_math=ldeq.MLP_MATH_BF16X3)
    assert e.value.code == -3     # LDEQ_ERR_UNSUPPORTED, never a silent fallback
    with pytest.raises(ldeq.LdeqError):
        _solve(ldeq, z0.astype(np.float64), np.zeros(om.n_params([16, 200, 200, 16])), [16, 200, 200, 16], 0.05 * np.arange(5),
               mlp_math=ldeq.MLP_MATH_BF16X3)


# ---- row a13: the reference's own reverse pass (InterpolatingAdjoint) ------------------------------------------------
def _cadj(ldeq, z0, p, dims, t, d, trace=0, **kw):
    import ctypes
    opts = ldeq.default_opts(sensealg=ldeq.SENSE_INTERPOLATING_ADJOINT, **kw)
    z = torch.from_numpy(z0).to(DEV)
    pp = torch.from_numpy(p).to(DEV)
    traj, st, tape = ldeq.mlp_solve_raw(z, pp, dims, t, opts, want_tape=True)
    gz, gp = ldeq.mlp_bwd_raw(tape, torch.from_numpy(d).to(DEV))
    stats = ldeq.mlp_bwd_stats(tape)
    tr = None
    if trace:
        buf = (ctypes.c_double * (4 * trace))()
        tape.h.check(tape.h._lib.ldeq_debug_cadj_trace(tape.h.ptr, tape.ptr, buf, trace))
        tr = np.array(buf).reshape(trace, 4)
    tape.free()
    return traj.cpu().numpy(), gz.cpu().numpy(), gp.cpu().numpy(), stats, tr


@pytest.mark.parametrize("adaptive", [True, False])
def test_interpolating_adjoint_fp64_follows_the_oracle_step_for_step(ldeq, adaptive, monkeypatch):
    # fp64: kernel and oracle interpolate the same forward steps and integrate [lambda; mu] backwards with the same
    # controller.  Fixed step: identical steps, gradients equal to rounding.  Adaptive: the device controller proposes dt
    # through Float32 arithmetic (1e-7 relative), and the relu right-hand side makes the error estimate jump wherever a mask
    # flips inside a step, so the sequences agree attempt by attempt at first (asserted on the first 16: time, dt, EEst,
    # accept/reject) and drift apart at borderline decisions later; the gradients then agree to the solver tolerance.
    monkeypatch.setenv("LDEQ_CADJ_TRACE", "1")
    dims, p, rng = _net(dtype="float64", bias_scale=0.1)
    B, T = 24, 20
    z0 = 0.5 * rng.standard_normal((B, 16))
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, 16))
    kw = dict(controller_pow=1) if adaptive else dict(adaptive=False, dt=0.02)
    okw = og.Opts(controller_pow=1) if adaptive else og.Opts(adaptive=False, dt=0.02)
    tr, gz, gp, (na, nr, ret), trace = _cadj(ldeq, z0, p, dims, t, d, trace=16 if adaptive else 0, **kw)
    otr, _, _, tape = om.solve(z0, p, dims, t, okw, record=True)
    st = {}
    oz, op = om.interpolating_adjoint(z0, p, dims, t, d, okw, tape=tape, stats=st)
    ez, ep = np.abs(gz - oz).max() / np.abs(oz).max(), np.abs(gp - op).max() / np.abs(op).max()
    print("backward solve", na, nr, ret, "oracle", st["naccept"], st["nreject"], "dz0", ez, "dp", ep)
    assert ret == ldeq.RET_SUCCESS
    assert np.abs(tr - otr).max() <= 1e-8 * np.abs(otr).max()
    if adaptive:
        otrace = np.array(st["trace"][:16])
        assert np.array_equal(trace[:, 3], otrace[:, 3])
        assert np.allclose(trace[:, :2], otrace[:, :2], rtol=1e-5, atol=0)
        assert np.allclose(trace[:, 2], otrace[:, 2], rtol=1e-3, atol=1e-12)
        assert abs(na - st["naccept"]) <= 0.1 * st["naccept"] and ez <= 5e-3 and ep <= 1e-2
        # and it is NOT the discrete adjoint: the two differ by the discretisation error (row a13's point)
        dz, dp = om.discrete_adjoint(p, dims, t, tape, d)
        assert np.abs(gp - dp).max() > 10 * np.abs(gp - op).max()
    else:
        assert (na, nr) == (st["naccept"], st["nreject"]) and ez <= 1e-9 and ep <= 1e-9


def test_interpolating_adjoint_fp32_c2_shape(ldeq):
    # C2 (B = 256, T = 50, glorot weights): fp32 kernel vs the fp64 oracle on the kernel's own forward steps
    dims, p, rng = _net()
    B, T = 256, 50
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(np.float32)
    t = 0.05 * np.arange(T)
    d = rng.standard_normal((T, B, 16)).astype(np.float32)
    tr, gz, gp, (na, nr, ret), _ = _cadj(ldeq, z0, p, dims, t, d)
    otr, _, _, tape = om.solve(z0, p, dims, t, og.Opts(), record=True)
    st = {}
    oz, op = om.interpolating_adjoint(z0, p, dims, t, d, og.Opts(), tape=tape, stats=st)
    ez, ep = np.abs(gz - oz).max() / np.abs(oz).max(), np.abs(gp - op).max() / np.abs(op).max()
    print("backward solve", na, nr, "oracle", st["naccept"], st["nreject"], "dz0", ez, "dp", ep)
    # the count of backward steps is a chaotic function of the Float32 summation order (the relu right-hand side makes the
    # error estimate jump wherever a mask flips inside a step): within 10 % of the Float64 oracle's; the gradients are the test
    assert abs(na - st["naccept"]) <= 0.10 * st["naccept"] + 2
    assert ez <= 5e-3 and ep <= 5e-3


def test_interpolating_adjoint_needs_the_reference_configuration(ldeq):
    dims, p, rng = _net()
    z = torch.from_numpy((0.5 * rng.standard_normal((8, 16))).astype(np.float32)).to(DEV)
    pp = torch.from_numpy(p).to(DEV)
    t = 0.05 * np.arange(10)
    d = torch.zeros(10, 8, 16, device=DEV)
    for kw in (dict(norm_mode=1), dict(mlp_math=1)):
        _, _, tape = ldeq.mlp_solve_raw(z, pp, dims, t, ldeq.default_opts(sensealg=ldeq.SENSE_INTERPOLATING_ADJOINT, **kw),
                                        want_tape=True)
        with pytest.raises(ldeq.LdeqError, match="interpolating adjoint"):
            ldeq.mlp_bwd_raw(tape, d)
        tape.free()
    # and the GOKU entry points refuse it
    with pytest.raises(ldeq.LdeqError):
        ldeq.goku_solve_raw(torch.zeros(4, 2, device=DEV), torch.ones(4, 1, device=DEV), t, 0,
                            ldeq.default_opts(sensealg=ldeq.SENSE_INTERPOLATING_ADJOINT), want_tape=True)
