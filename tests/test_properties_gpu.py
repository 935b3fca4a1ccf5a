"""Size-independent properties at BASELINE.json's full sizes (the oracle only finishes samples of them in seconds):
trajectories are independent problems (GOKU.jl:111: `prob_func` indexes column i), so a solve commutes with any split or
permutation of the batch, bit for bit; a pullback is linear in its cotangent; the host-buffer entry points return what the
device-pointer ones return.  C4 at 2^20 x 200 (GOKU), C5's per-GPU share 8192 x 50 (pattern extractor)."""
import numpy as np
import pytest
import torch

from conftest import pendulum_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _solve_grad(ldeq, z0, th, t, d, **kw):
    z = z0.clone().requires_grad_(True)
    p = th.clone().requires_grad_(True)
    st = []
    tr = ldeq.goku_solve(z, p, t, 0, ldeq.default_opts(**kw), st)
    tr.backward(d)
    return tr.detach(), z.grad, p.grad, st[0]


@pytest.mark.parametrize("sense", ["forward_dual", "discrete_adjoint"])
def test_c4_full_size_split_and_permutation_invariance(ldeq, sense):
    B, T = 1 << 20, 200
    code = ldeq.SENSE_FORWARD_DUAL if sense == "forward_dual" else ldeq.SENSE_DISCRETE_ADJOINT
    z0n, thn = pendulum_inputs(B)
    z0, th = torch.from_numpy(z0n).to(DEV), torch.from_numpy(thn).to(DEV)
    t = 0.05 * np.arange(T)
    g = torch.Generator(device=DEV)
    g.manual_seed(7)
    d = torch.randn(T, B, 2, device=DEV, generator=g)
    tr, gz, gp, st = _solve_grad(ldeq, z0, th, t, d, sensealg=code)
    assert int((st.retcode != 0).sum()) == 0
    # a ragged split: the same columns solved as three separate batches are bit-identical
    cuts = [0, 333_333, 700_001, B]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        tr2, gz2, gp2, st2 = _solve_grad(ldeq, z0[lo:hi], th[lo:hi], t, d[:, lo:hi].contiguous(), sensealg=code)
        assert torch.equal(tr2, tr[:, lo:hi]) and torch.equal(gz2, gz[lo:hi]) and torch.equal(gp2, gp[lo:hi])
        assert torch.equal(st2.naccept, st.naccept[lo:hi])
    # a permutation of the batch permutes the results (other CTAs, other lanes, other warps' step-count mix)
    perm = torch.randperm(B, device=DEV, generator=g)
    tr3, gz3, gp3, _ = _solve_grad(ldeq, z0[perm], th[perm], t, d[:, perm].contiguous(), sensealg=code)
    assert torch.equal(tr3, tr[:, perm]) and torch.equal(gz3, gz[perm]) and torch.equal(gp3, gp[perm])


@pytest.mark.parametrize("sense", ["forward_dual", "discrete_adjoint"])
def test_c4_full_size_pullback_is_linear_in_the_cotangent(ldeq, sense):
    B, T = 1 << 20, 200
    code = ldeq.SENSE_FORWARD_DUAL if sense == "forward_dual" else ldeq.SENSE_DISCRETE_ADJOINT
    z0n, thn = pendulum_inputs(B)
    z0, th = torch.from_numpy(z0n).to(DEV), torch.from_numpy(thn).to(DEV)
    t = 0.05 * np.arange(T)
    g = torch.Generator(device=DEV)
    g.manual_seed(8)
    d1, d2 = torch.randn(T, B, 2, device=DEV, generator=g), torch.randn(T, B, 2, device=DEV, generator=g)
    _, z1, p1, _ = _solve_grad(ldeq, z0, th, t, d1, sensealg=code)
    _, z2, p2, _ = _solve_grad(ldeq, z0, th, t, d2, sensealg=code)
    _, z3, p3, _ = _solve_grad(ldeq, z0, th, t, 0.5 * d1 - 2.0 * d2, sensealg=code)
    # Float32 sums over 200 save points: 1e-5 of the gradient scale on every trajectory
    for a, b, c in ((z1, z2, z3), (p1, p2, p3)):
        want = 0.5 * a - 2.0 * b
        assert float((c - want).abs().max()) <= 1e-5 * float(want.abs().max())
    # the zero cotangent gives exactly zero
    _, z4, p4, _ = _solve_grad(ldeq, z0, th, t, torch.zeros_like(d1), sensealg=code)
    assert float(z4.abs().max()) == 0.0 and float(p4.abs().max()) == 0.0


def test_c4_full_size_host_entry_points_equal_the_device_path(ldeq):
    B, T = 1 << 20, 200
    z0n, thn = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    g = torch.Generator(device=DEV)
    g.manual_seed(9)
    d = torch.randn(T, B, 2, device=DEV, generator=g)
    tr, gz, gp, _ = _solve_grad(ldeq, torch.from_numpy(z0n).to(DEV), torch.from_numpy(thn).to(DEV), t, d)
    hz0, hth, hd = torch.from_numpy(z0n).pin_memory(), torch.from_numpy(thn).pin_memory(), d.cpu().pin_memory()
    out = ldeq.goku_fwd_bwd_host(hz0, hth, t, hd, 0, ldeq.default_opts(), device=0)       # eight column slabs, overlapped copies
    htr, hgz, hgp = out[0], out[1], out[2]
    assert torch.equal(htr, tr.cpu()) and torch.equal(hgz, gz.cpu()) and torch.equal(hgp, gp.cpu())


def test_c5_share_pattern_extractor_split_and_linearity(ldeq):
    from latentdiffeq_jl_b200.solve import _PatternExtractor
    from oracle import recurrent as orr
    T, B, F = 50, 8192, 32
    rng = np.random.default_rng(3)
    ps = [torch.from_numpy(orr.init_params(l, F, rng)).to(DEV) for l in (False, True, True)]
    g = torch.Generator(device=DEV)
    g.manual_seed(10)
    x = torch.randn(T, B, F, device=DEV, generator=g)
    w1, w2 = torch.randn(B, 48, device=DEV, generator=g), torch.randn(B, 48, device=DEV, generator=g)

    def run(xs, w):
        xs = xs.clone().requires_grad_(True)
        pp = [p.clone().requires_grad_(True) for p in ps]
        z0, th = _PatternExtractor.apply(xs, *pp)
        (torch.cat([z0, th], 1) * w).sum().backward()
        return z0.detach(), th.detach(), xs.grad, [p.grad for p in pp]
    z0, th, dx, dps = run(x, w1)
    # sequences are independent: a ragged split gives the same final states and frame cotangents bit for bit, and the
    # parameter gradients add up (other CTA partition, hence rounding-level differences only)
    acc = [torch.zeros_like(p) for p in ps]
    for lo, hi in ((0, 3000), (3000, 8192)):
        a, b, c, dd = run(x[:, lo:hi].contiguous(), w1[lo:hi])
        assert torch.equal(a, z0[lo:hi]) and torch.equal(b, th[lo:hi]) and torch.equal(c, dx[:, lo:hi])
        for s_, d_ in zip(acc, dd):
            s_ += d_
    for s_, d_ in zip(acc, dps):
        assert float((s_ - d_).abs().max()) <= 2e-5 * float(d_.abs().max())
    # linear in the cotangent
    _, _, dx2, dps2 = run(x, w2)
    _, _, dx3, dps3 = run(x, 0.5 * w1 - 2.0 * w2)
    assert float((dx3 - (0.5 * dx - 2.0 * dx2)).abs().max()) <= 1e-5 * float(dx3.abs().max())
    for a, b, c in zip(dps, dps2, dps3):
        assert float((c - (0.5 * a - 2.0 * b)).abs().max()) <= 2e-5 * float(c.abs().max())


@pytest.mark.parametrize("path", ["resident_fp32", "tensor_core_bf16x3", "general_fp64"])
def test_latentode_per_trajectory_mode_split_invariance(ldeq, path):
    # LDEQ_NORM_PER_TRAJ (the documented deviation for sharded batches): every row has its own controller, so the rows of a
    # solve are independent problems here too -- trajectories and dz0 commute bit for bit with a ragged split of the batch
    # on every kernel family (rows share a tile / an MMA with other rows, never an arithmetic result)
    from oracle import mlp as om
    rng = np.random.Generator(np.random.PCG64(1))
    dims = [16, 200, 200, 16]
    dt_ = np.float64 if path == "general_fp64" else np.float32
    p = om.pack_params([(om.glorot_uniform(rng, dims[i + 1], dims[i]), np.zeros(dims[i + 1], np.float32)) for i in range(3)]).astype(dt_)
    B, T = (300, 50) if path != "tensor_core_bf16x3" else (1000, 50)
    z0 = (0.5 * rng.standard_normal((B, 16))).astype(dt_)
    d = rng.standard_normal((T, B, 16)).astype(dt_)
    t = 0.05 * np.arange(T)
    kw = dict(norm_mode=ldeq.NORM_PER_TRAJ, sensealg=ldeq.SENSE_DISCRETE_ADJOINT)
    if path == "tensor_core_bf16x3":
        kw["mlp_math"] = ldeq.MLP_MATH_BF16X3

    def run(lo, hi):
        z = torch.from_numpy(z0[lo:hi]).to(DEV).requires_grad_(True)
        q = torch.from_numpy(p).to(DEV).requires_grad_(True)
        tr = ldeq.mlp_solve(z, q, dims, t, ldeq.default_opts(**kw))
        tr.backward(torch.from_numpy(np.ascontiguousarray(d[:, lo:hi])).to(DEV))
        return tr.detach(), z.grad, q.grad
    tr, gz, gq = run(0, B)
    acc = torch.zeros_like(gq)
    for lo, hi in ((0, 101), (101, B)):
        tr2, gz2, gq2 = run(lo, hi)
        assert torch.equal(tr2, tr[:, lo:hi]), path
        assert torch.equal(gz2, gz[lo:hi]), path
        acc += gq2
    assert float((acc - gq).abs().max()) <= (1e-12 if path == "general_fp64" else 2e-5) * float(gq.abs().max())
