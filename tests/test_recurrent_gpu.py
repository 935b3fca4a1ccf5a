"""The recurrent pattern extractor as persistent kernels (GOKU.jl:30-49, 224-234) through the C ABI,
against the float64 per-step restatement of the Flux cells in ``oracle/recurrent.py``; Float32 kernels: 2e-5 relative on
the final states, 2e-4 on the gradients (sums over B x T products)."""
import numpy as np
import pytest
import torch

from oracle import recurrent as orr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _raw(ldeq, x, rnn, lf, lb, dz0=None, dth=None):
    t = lambda a: None if a is None else torch.from_numpy(a).to(DEV).requires_grad_(dz0 is not None)
    xs, ps = t(x), [t(rnn), t(lf), t(lb)]
    from latentdiffeq_jl_b200.solve import _PatternExtractor
    z0, th = _PatternExtractor.apply(xs, *ps)
    if dz0 is None:
        return z0.cpu().numpy(), th.cpu().numpy() if lf is not None else None
    loss = (z0 * torch.from_numpy(dz0).to(DEV)).sum()
    if lf is not None:
        loss = loss + (th * torch.from_numpy(dth).to(DEV)).sum()
    loss.backward()
    torch.cuda.synchronize()
    return (z0.detach().cpu().numpy(), th.detach().cpu().numpy() if lf is not None else None,
            [xs.grad.cpu().numpy()] + [None if p is None else p.grad.cpu().numpy() for p in ps])


@pytest.mark.parametrize("F", [32, 16, 64])
@pytest.mark.parametrize("B,T", [(300, 50), (37, 7), (1, 1)])
def test_forward_and_gradients_match_the_oracle(ldeq, F, B, T):
    rng = np.random.default_rng(F * 1000 + B)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    rnn, lf, lb = orr.init_params(False, F, rng), orr.init_params(True, F, rng), orr.init_params(True, F, rng)
    assert rnn.size == orr.param_count(False, F) and lf.size == orr.param_count(True, F)
    dz0 = rng.standard_normal((B, 16)).astype(np.float32)
    dth = rng.standard_normal((B, 32)).astype(np.float32)
    z0, th, g = _raw(ldeq, x, rnn, lf, lb, dz0, dth)
    oz0, oth, og_ = orr.pattern_extractor(x, rnn, lf, lb, dz0, dth)
    assert _rel(z0, oz0) < 2e-5 and _rel(th, oth) < 2e-5, (_rel(z0, oz0), _rel(th, oth))
    names = ["dx", "d_rnn", "d_lstm_f", "d_lstm_b"]
    for n, a, b in zip(names, g, og_):
        assert a.shape == b.shape and _rel(a, b) < 2e-4, (n, _rel(a, b))
    # forward only (no tape) gives the same numbers
    z0b, thb = _raw(ldeq, x, rnn, lf, lb)
    assert np.array_equal(z0, z0b) and np.array_equal(th, thb)


def test_reverse_pass_is_deterministic(ldeq):
    # no floating-point atomics: per-CTA partial gradients and the three stacks' dx buffers are added in a fixed order
    # (the multi-rank self-test of the training bench compares two steps from identical state)
    rng = np.random.default_rng(11)
    T, B, F = 50, 2000, 32
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    rnn, lf, lb = orr.init_params(False, F, rng), orr.init_params(True, F, rng), orr.init_params(True, F, rng)
    dz0 = rng.standard_normal((B, 16)).astype(np.float32)
    dth = rng.standard_normal((B, 32)).astype(np.float32)
    a = _raw(ldeq, x, rnn, lf, lb, dz0, dth)
    b = _raw(ldeq, x, rnn, lf, lb, dz0, dth)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for ga, gb in zip(a[2], b[2]):
        assert np.array_equal(ga, gb)


def test_latentode_rnn_only(ldeq):
    rng = np.random.default_rng(5)
    T, B, F = 50, 256, 32
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    rnn = orr.init_params(False, F, rng)
    dz0 = rng.standard_normal((B, 16)).astype(np.float32)
    z0, th, g = _raw(ldeq, x, rnn, None, None, dz0)
    oz0, _, og_ = orr.pattern_extractor(x, rnn, None, None, dz0)
    assert th is None and _rel(z0, oz0) < 2e-5
    assert _rel(g[0], og_[0]) < 2e-4 and _rel(g[1], og_[1]) < 2e-4


@pytest.mark.parametrize("F", [32, 64])
@pytest.mark.parametrize("B,T", [(300, 50), (19, 5)])
def test_latentode_default_rnn_stack_32_hidden_units(ldeq, F, B, T):
    # LatentODE's default pattern extractor: Chain(RNN(32,32,relu), RNN(32,32,relu)) on the reversed sequence (LatentODE.jl:100-124)
    rng = np.random.default_rng(F + B)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    rnn = orr.init_params(False, F, rng, H=32)
    assert rnn.size == orr.param_count(False, F, 32)
    dz0 = rng.standard_normal((B, 32)).astype(np.float32)
    z0, th, g = _raw(ldeq, x, rnn, None, None, dz0)
    oz0, _, og_ = orr.pattern_extractor(x, rnn, None, None, dz0, H=32)
    assert z0.shape == (B, 32) and _rel(z0, oz0) < 2e-5
    assert _rel(g[0], og_[0]) < 2e-4 and _rel(g[1], og_[1]) < 2e-4
    a = _raw(ldeq, x, rnn, None, None, dz0)
    assert np.array_equal(a[2][0], g[0]) and np.array_equal(a[2][1], g[1])      # deterministic
    # the LSTM stacks are built for 16 hidden units only: refused, not replaced
    from latentdiffeq_jl_b200.solve import _PatternExtractor
    with pytest.raises(ldeq.LdeqError):
        _PatternExtractor.apply(torch.from_numpy(x).to(DEV), torch.from_numpy(rnn).to(DEV),
                                torch.zeros(orr.param_count(True, F, 32), device=DEV), torch.zeros(orr.param_count(True, F, 32), device=DEV))


def test_latentode_model_route(ldeq):
    # the LatentODE encoder of default_layers goes through the kernels (no cuDNN call) and matches the per-step route
    import latentdiffeq_jl_b200 as L
    model_mod = __import__(L.__name__ + ".model", fromlist=["x"]) if hasattr(L, "__path__") else L.model
    torch.manual_seed(0)
    enc, dec = ldeq.default_layers(ldeq.LatentODE(), 28 * 28, ldeq.NODE(16), device=DEV)
    model = ldeq.LatentDiffEqModel(ldeq.LatentODE(), enc, dec)
    fe = torch.randn(50, 64, 32, device=DEV, requires_grad=True)
    assert model_mod._pe_kernel_ok(fe, model.encoder.pattern_extractor)
    w = torch.randn(64, 32, device=DEV)
    outs = {}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    for flag in (True, False):
        model_mod.PERSISTENT_RECURRENT = flag
        model.zero_grad(); fe.grad = None
        out = ldeq.apply_pattern_extractor(model.encoder, fe)
        (out * w).sum().backward()
        outs[flag] = (out.detach().clone(), fe.grad.clone(), [p.grad.clone() for p in model.encoder.pattern_extractor.parameters()])
    model_mod.PERSISTENT_RECURRENT = True
    torch.backends.cudnn.allow_tf32 = tf32
    a, b = outs[True], outs[False]
    assert torch.allclose(a[0], b[0], rtol=1e-4, atol=1e-5) and (a[1] - b[1]).abs().max() <= 2e-4 * b[1].abs().max()
    for ga, gb in zip(a[2], b[2]):
        assert (ga - gb).abs().max() <= 5e-4 * max(gb.abs().max().item(), 1e-3)


def test_model_route_equals_the_cudnn_route(ldeq):
    # the default GOKU encoder: apply_pattern_extractor through the kernels vs the cuDNN / per-step route, values and
    # gradients of every recurrent parameter and of the frames
    import latentdiffeq_jl_b200 as L
    model_mod = __import__(L.__name__ + ".model", fromlist=["x"]) if hasattr(L, "__path__") else L.model
    torch.manual_seed(0)
    enc, dec = ldeq.default_layers(ldeq.GOKU(), 28 * 28, ldeq.Pendulum(), device=DEV)
    model = ldeq.LatentDiffEqModel(ldeq.GOKU(), enc, dec)
    for p in model.encoder.pattern_extractor.parameters():      # state0 / h0 / c0 are zeros by default: make them count
        if p.dim() == 1 and p.abs().sum() == 0:
            torch.nn.init.uniform_(p, -0.3, 0.3)
    fe = torch.randn(50, 200, 32, device=DEV, requires_grad=True)
    w = torch.randn(200, 48, device=DEV)
    outs = {}
    # cuDNN's RNN kernels run on TF32 tensor cores by default (the two routes then differ by 5e-3); compare in Float32
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    for flag in (True, False):
        model_mod.PERSISTENT_RECURRENT = flag
        model.zero_grad()
        fe.grad = None
        z0, th = ldeq.apply_pattern_extractor(model.encoder, fe)
        (torch.cat([z0, th], 1) * w).sum().backward()
        outs[flag] = (z0.detach().clone(), th.detach().clone(), fe.grad.clone(),
                      [p.grad.clone() for p in model.encoder.pattern_extractor.parameters()])
    model_mod.PERSISTENT_RECURRENT = True
    torch.backends.cudnn.allow_tf32 = tf32
    a, b = outs[True], outs[False]
    assert torch.allclose(a[0], b[0], rtol=1e-4, atol=1e-5) and torch.allclose(a[1], b[1], rtol=1e-4, atol=1e-5)
    assert (a[2] - b[2]).abs().max() <= 2e-4 * b[2].abs().max()
    for ga, gb in zip(a[3], b[3]):
        assert (ga - gb).abs().max() <= 5e-4 * max(gb.abs().max().item(), 1e-3), (ga - gb).abs().max()


def test_unsupported_shapes_are_refused_not_replaced(ldeq):
    x = torch.zeros(5, 8, 24, device=DEV)      # F = 24 is not built
    from latentdiffeq_jl_b200.solve import _PatternExtractor
    with pytest.raises(ldeq.LdeqError) as e:
        _PatternExtractor.apply(x, torch.zeros(10, device=DEV), None, None)
    assert e.value.code == -3
