"""CPU tier: host-side mirror of the reference API (layers, model container, utilities, data-parallel helpers)."""
import numpy as np
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch


def test_default_goku_architecture(ldeq):
    mt = ldeq.GOKU_basic()
    enc, dec = ldeq.default_layers(mt, 784, ldeq.Pendulum())
    model = ldeq.LatentDiffEqModel(mt, enc, dec)
    # 503 387 trainable parameters; the GOKU ODE has none of its own
    assert sum(p.numel() for p in model.parameters()) == 503387
    x = torch.rand(50, 8, 784)
    mu, lv = model.encoder(x)
    assert [tuple(m.shape) for m in mu] == [(8, 16), (8, 16)] and [tuple(l.shape) for l in lv] == [(8, 16), (8, 16)]
    z0h, thh = ldeq.apply_latent_out(model.decoder, mu)
    assert z0h.shape == (8, 2) and thh.shape == (8, 1) and (thh > 0).all()      # softplus
    xh = ldeq.apply_reconstructor(model.decoder, torch.rand(50, 8, 2))
    assert xh.shape == (50, 8, 784) and (xh >= 0).all() and (xh <= 1).all()      # sigmoid output
    # hidden state restarts from state0 on every call (Flux.reset!)
    a = model.encoder(x)[0][0]
    b = model.encoder(x)[0][0]
    assert torch.equal(a, b)
    # Flux kaiming_uniform(gain = 1/sqrt(3)) = U(+-1/sqrt(fan_in)); zero bias; LSTM forget-gate bias 1
    w = enc[0][0].weight
    assert w.abs().max() <= 1 / np.sqrt(784) + 1e-7 and enc[0][0].bias.abs().sum() == 0
    lstm = enc[1][1][0]
    assert torch.equal(lstm.b[16:32], torch.ones(16)) and lstm.b[:16].abs().sum() == 0


def test_default_latentode_architecture(ldeq):
    mt = ldeq.LatentODE()
    node = ldeq.NODE(16)
    enc, dec = ldeq.default_layers(mt, 784, node)
    model = ldeq.LatentDiffEqModel(mt, enc, dec)
    assert node.flat_params().numel() == 46816 and node.dims == [16, 200, 200, 16]
    assert sum(p.numel() for p in model.parameters()) == 537312
    mu, lv = model.encoder(torch.rand(20, 4, 784))
    assert mu.shape == (4, 16) and lv.shape == (4, 16)
    # destructure order: vec(W) column-major (W is (out, in)) then b
    fp = node.flat_params()
    assert fp[1] == node.weights[0][1, 0] and fp[200] == node.weights[0][0, 1] and fp[3200] == node.biases[0][0]
    assert ldeq.NODE(16, augment_dim=2).latent_dim_out == 18


def test_diffeq_struct_contract(ldeq):
    p = ldeq.Pendulum(abstol=1e-8)
    assert len(p.prob.u0) == 2 and len(p.prob.p) == 1 and p.kwargs == {"abstol": 1e-8}
    assert repr(p.solver) == "Tsit5()" and p.prob.f == ldeq.RHS_PENDULUM
    assert ldeq.Pendulum_friction().prob.f == ldeq.RHS_PENDULUM_FRICTION


def test_sensealg_and_solver_map_to_the_solver_options(ldeq):
    # the diffeq struct's sensealg / solver fields (pendulum.jl:11) are read, not ignored
    from importlib import import_module
    model = import_module(ldeq.__name__ + ".model") if hasattr(ldeq, "__path__") else ldeq.model
    p = ldeq.Pendulum()
    assert repr(p.sensealg) == "ForwardDiffSensitivity()" and repr(p.solver) == "Tsit5()"
    o = model._opts_from_kwargs(p.kwargs, p.sensealg, p.solver)
    assert o.sensealg == ldeq.SENSE_FORWARD_DUAL == 1 and o.solver == 0      # the reference's own algorithm by default
    q = ldeq.Pendulum(sensalg=ldeq.DiscreteAdjoint(), reltol=1e-5)            # the cheaper discrete adjoint: explicit opt-in
    assert repr(q.sensealg) == "DiscreteAdjoint()"
    o = model._opts_from_kwargs(q.kwargs, q.sensealg, q.solver)
    assert o.sensealg == ldeq.SENSE_DISCRETE_ADJOINT == 0 and o.reltol == 1e-5
    o = model._opts_from_kwargs({"saveat": [0.0, 1.0], "abstol": 1e-9}, None)      # saveat is the layer's own argument
    assert o.sensealg == ldeq.SENSE_FORWARD_DUAL and o.abstol == 1e-9
    with pytest.raises(TypeError):
        model._opts_from_kwargs({}, object())
    with pytest.raises(NotImplementedError):
        model._opts_from_kwargs({}, None, "Rodas5()")
    # the `solver` field is the user's (GOKU.jl:121): DP5 / BS3 / RK4 carry OrdinaryDiffEq's controller defaults
    r = ldeq.Pendulum(solver=ldeq.DP5(), reltol=1e-4)
    o = model._opts_from_kwargs(r.kwargs, r.sensealg, r.solver)
    assert repr(r.solver) == "DP5()" and o.solver == ldeq.SOLVER_DP5 == 1 and o.beta2 == 0.04 and o.reltol == 1e-4
    o = model._opts_from_kwargs({"adaptive": False, "dt": 0.01}, None, ldeq.RK4())
    assert o.solver == ldeq.SOLVER_RK4 == 3 and o.adaptive == 0 and o.beta2 == pytest.approx(0.1)


@pytest.mark.parametrize("method", ["Tsit5M|Tsit5Dual", "DP5M|ErkDual<TabDP5>", "BS3M|ErkDual<TabBS3>", "RK4M|ErkDual<TabRK4>"])
def test_user_rhs_translation_unit_compiles_with_nvrtc(method):
    """The text ldeq_rhs_from_source hands to NVRTC (the integrator headers embedded in the library + a user function +
    the kernel entry points) compiles for sm_100a -- one variant per value of the diffeq struct's `solver` field; NVRTC
    needs no GPU."""
    import re
    nvrtc = pytest.importorskip("cuda.bindings.nvrtc")
    csrc = os.path.join(ROOT, "latentdiffeq.jl_b200", "csrc")

    def strip(fn):
        txt = open(os.path.join(csrc, fn)).read()
        txt = re.sub(r'^\s*#pragma once\s*$', '', txt, flags=re.M)
        return re.sub(r'^\s*#include "ldeq_[a-z0-9_]+\.cuh"\s*$', '', txt, flags=re.M)
    wrapper = re.search(r'kWrapper = R"LDEQ\((.*?)\)LDEQ";', open(os.path.join(csrc, "ldeq_user_rhs.cu")).read(), re.S).group(1)
    user = ("template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t) { du[0] = u[1]; "
            "du[1] = -p[0] * sin(u[0]) - S(0.1) * u[1]; du[2] = p[1] * u[0] - u[2] + exp(-t); }")
    m, md = method.split("|")
    assert f"#define LDEQ_USER_M {m}\\n#define LDEQ_USER_MD {md}\\n" in open(os.path.join(csrc, "ldeq_user_rhs.cu")).read()
    src = (f"#define LDEQ_USER_ZD 3\n#define LDEQ_USER_PD 2\n#define LDEQ_USER_M {m}\n#define LDEQ_USER_MD {md}\n" + "".join(strip(f) for f in (
        "ldeq_common.cuh", "ldeq_erk.cuh", "ldeq_tsit5.cuh", "ldeq_julia_trig.cuh", "ldeq_dual.cuh", "ldeq_fwdsens.cuh")) + "\n" + user + "\n" + wrapper)
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"u.cu", 0, [], [])
    opts = [b"--gpu-architecture=sm_100a", b"--std=c++17", b"-default-device", b"--fmad=true", b"-DLDEQ_NVRTC=1"]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    assert int(err) == 0, log.decode()[:2000]


def test_utils(ldeq):
    mu, lv = torch.randn(5, 3), torch.randn(5, 3)
    assert torch.allclose(ldeq.kl(mu, lv), (lv.exp() + mu ** 2 - lv - 1) / 2)
    assert torch.allclose(ldeq.vector_kl((mu, mu), (lv, lv)), 2 * ldeq.kl(mu, lv).sum() / 5)
    assert torch.allclose(ldeq.vector_kl(mu, lv), ldeq.kl(mu, lv).sum() / 5)
    L = ldeq.frange_cycle_linear(20, 0.0, 1.0, 4, 0.5)
    assert L.dtype == np.float32 and np.allclose(L[:5], [0, 0.4, 0.8, 1, 1]) and np.allclose(L[5:10], L[:5])
    L = ldeq.frange_cycle_linear(1500, 0.0, 1.0, 4, 0.9)       # the example's schedule (model_train.jl:50-54)
    assert L[0] == 0 and L.max() == 1 and L[374] == 1 and L[375] == 0
    x = torch.arange(100 * 2 * 3, dtype=torch.float64).reshape(100, 2, 3)
    w = ldeq.time_loader(x, 100, 50, np.random.default_rng(0))
    assert w.shape == (50, 2, 3) and w.dtype == torch.float32 and torch.equal(w[1:, 0, 0] - w[:-1, 0, 0], torch.full((49,), 6.0))
    starts = {ldeq.rand_time(100, 50, np.random.default_rng(s)).start for s in range(400)}
    assert min(starts) == 0 and max(starts) == 49     # rand(1:50): the window starting at the 51st frame is never drawn
    Xn, mn, mx = ldeq.normalize_to_unit_segment(torch.tensor([2.0, 4.0, 6.0]))
    assert torch.equal(Xn, torch.tensor([0.0, 0.5, 1.0])) and torch.equal(ldeq.denormalize_unit_segment(Xn, mn, mx), torch.tensor([2.0, 4.0, 6.0]))


def test_shard_bounds_and_flat_params(ldeq):
    for n, w in [(64, 2), (65536, 8), (10, 3), (5, 8)]:
        cuts = [ldeq.shard_bounds(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1
    lin = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Linear(5, 2))
    ref = [p.detach().clone() for p in lin.parameters()]
    flat = ldeq.FlatParams(lin)
    assert flat.n == 3 * 5 + 5 + 5 * 2 + 2 and flat.flat.numel() % 4 == 0
    for p, r in zip(lin.parameters(), ref):
        assert torch.equal(p, r) and p.data_ptr() >= flat.flat.data_ptr()
    flat.zero_grad()
    lin(torch.ones(4, 3)).sum().backward()
    assert flat.grad[:flat.n].abs().sum() > 0 and torch.equal(lin[0].weight.grad.reshape(-1), flat.grad[:15])


def test_hot_path_refuses_cpu_tensors(ldeq):
    mt = ldeq.GOKU_basic()
    enc, dec = ldeq.default_layers(mt, 784, ldeq.Pendulum())
    model = ldeq.LatentDiffEqModel(mt, enc, dec)
    with pytest.raises(RuntimeError):
        model(torch.rand(10, 2, 784), np.arange(10) * 0.05)
