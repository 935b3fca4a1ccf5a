"""GPU parity tests of the ELBO reduction, fused AdamW and the reparameterised sample."""
import numpy as np
import pytest
import torch

from oracle import loss as ol

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,T,P,heads", [(64, 50, 784, (16, 16)), (7, 3, 5, (4,)), (33, 9, 101, (16, 16))])
def test_elbo_matches_oracle(ldeq, B, T, P, heads):
    rng = np.random.default_rng(3)
    x = rng.random((T, B, P), dtype=np.float32)
    xh = rng.random((T, B, P), dtype=np.float32)
    mu = [rng.standard_normal((B, d)).astype(np.float32) for d in heads]
    lv = [(0.3 * rng.standard_normal((B, d))).astype(np.float32) for d in heads]
    beta = 0.37
    tx, txh = torch.from_numpy(x).to(DEV), torch.from_numpy(xh).to(DEV)
    tmu, tlv = [torch.from_numpy(a).to(DEV) for a in mu], [torch.from_numpy(a).to(DEV) for a in lv]
    loss, dxh, dmu, dlv = ldeq.elbo_raw(tx, txh, tmu, tlv, beta)
    torch.cuda.synchronize()
    arg_mu, arg_lv = (tuple(mu), tuple(lv)) if len(heads) > 1 else (mu[0], lv[0])
    tot, rec, k = ol.loss_batch(x, xh, arg_mu, arg_lv, beta)
    l3 = loss.cpu().numpy()
    assert abs(l3[1] - rec) <= 1e-5 * abs(rec) and abs(l3[2] - k) <= 1e-5 * abs(k) and abs(l3[0] - tot) <= 1e-5 * abs(tot)
    odx, odmu, odlv = ol.loss_batch_grads(x, xh, arg_mu, arg_lv, beta)
    assert np.abs(dxh.cpu().numpy() - odx).max() <= 1e-6 * np.abs(odx).max()
    for a, b in zip(dmu, odmu):
        assert np.abs(a.cpu().numpy() - b).max() <= 1e-6 * np.abs(b).max()
    for a, b in zip(dlv, odlv):
        assert np.abs(a.cpu().numpy() - b).max() <= 2e-6 * np.abs(b).max()
    # the reduction is deterministic run to run
    loss2, *_ = ldeq.elbo_raw(tx, txh, tmu, tlv, beta)
    assert torch.equal(loss, loss2)
    # autograd wrapper == torch reference
    txh2 = txh.clone().requires_grad_(True)
    l = ldeq.elbo_loss(tx, txh2, tuple(tmu), tuple(tlv), beta)
    l.backward()
    ref = ((tx - txh) ** 2).mean(dim=(0, 1)).sum()
    assert abs(float(l) - float(tot)) <= 1e-5 * abs(float(tot))
    assert torch.allclose(txh2.grad, 2 * (txh - tx) / (B * T), rtol=1e-5, atol=1e-9) and ref > 0


@pytest.mark.parametrize("B,T,P", [(64, 50, 784), (7, 3, 5)])
def test_elbo_with_the_sigmoid_output_layer_folded_in(ldeq, B, T, P):
    # ldeq_elbo_logits_fwd_bwd: x-hat = sigmoid(a) formed inside the loss kernel, gradient with respect to the pre-activations a
    # (GOKU.jl:265-268: the reconstructor's last Dense layer has output_activation = sigmoid); oracle: the loss of sigmoid(a)
    # and its gradient chained through sigmoid' in float64
    rng = np.random.default_rng(5)
    x = rng.random((T, B, P), dtype=np.float32)
    a = (3.0 * rng.standard_normal((T, B, P))).astype(np.float32)
    mu = [rng.standard_normal((B, 16)).astype(np.float32), rng.standard_normal((B, 16)).astype(np.float32)]
    lv = [(0.3 * rng.standard_normal((B, 16))).astype(np.float32) for _ in range(2)]
    beta, gs = 0.37, 0.125
    tx, ta = torch.from_numpy(x).to(DEV), torch.from_numpy(a).to(DEV)
    tmu, tlv = [torch.from_numpy(v).to(DEV) for v in mu], [torch.from_numpy(v).to(DEV) for v in lv]
    loss, da, dmu, dlv = ldeq.elbo_raw(tx, ta, tmu, tlv, beta, logits=True, grad_scale=gs)
    sig = 1.0 / (1.0 + np.exp(-a.astype(np.float64)))
    tot, rec, k = ol.loss_batch(x, sig.astype(np.float32), tuple(mu), tuple(lv), beta)
    l3 = loss.cpu().numpy()
    assert abs(l3[1] - rec) <= 2e-5 * abs(rec) and abs(l3[0] - tot) <= 2e-5 * abs(tot)
    oda = gs * 2.0 * (sig - x) / (B * T) * sig * (1.0 - sig)
    assert np.abs(da.cpu().numpy() - oda).max() <= 2e-6 * np.abs(oda).max()
    odx, odmu, odlv = ol.loss_batch_grads(x, sig.astype(np.float32), tuple(mu), tuple(lv), beta)
    for got, want in zip(dmu + dlv, odmu + odlv):
        assert np.abs(got.cpu().numpy() - gs * want).max() <= 2e-6 * np.abs(gs * want).max()
    # the autograd routes agree: logits + unit cotangent + grad_scale  ==  sigmoid in torch, then the plain loss, scaled
    a1 = ta.clone().requires_grad_(True)
    ldeq.elbo_loss(tx, a1, tuple(tmu), tuple(tlv), beta, logits=True, unit_cotangent=True, grad_scale=gs).backward()
    a2 = ta.clone().requires_grad_(True)
    l2 = ldeq.elbo_loss(tx, torch.sigmoid(a2), tuple(tmu), tuple(tlv), beta)
    (l2 * gs).backward()
    assert torch.allclose(a1.grad, a2.grad, rtol=1e-4, atol=1e-12)


def test_loss_batch_fused_output_equals_the_plain_route(ldeq):
    # train.loss_batch on the default GOKU model: sigmoid folded into the loss kernel vs model(x) -> x-hat -> loss;
    # same loss, same parameter gradients
    torch.manual_seed(0)
    enc, dec = ldeq.default_layers(ldeq.GOKU(), 28 * 28, ldeq.Pendulum(), device=DEV)
    model = ldeq.LatentDiffEqModel(ldeq.GOKU(), enc, dec)
    x = torch.rand(20, 32, 784, device=DEV)
    t = 0.05 * np.arange(20)
    grads = {}
    for fused in (True, False):
        model.zero_grad()
        loss = ldeq.loss_batch(model, x, t, 0.5, False, fused_output=fused, unit_cotangent=fused, grad_scale=0.25 if fused else 1.0)
        (loss if fused else loss * 0.25).backward()
        grads[fused] = (float(loss), [p.grad.clone() for p in model.parameters() if p.grad is not None])
    assert abs(grads[True][0] - grads[False][0]) <= 1e-5 * abs(grads[False][0])
    assert len(grads[True][1]) == len(grads[False][1])
    for ga, gb in zip(grads[True][1], grads[False][1]):
        assert (ga - gb).abs().max() <= 2e-4 * max(float(gb.abs().max()), 1e-6)


def test_adamw_matches_flux_semantics(ldeq):
    rng = np.random.default_rng(0)
    n = 503387 + 46816  # default GOKU + NODE parameter count (not a multiple of 4)
    x = rng.standard_normal(n).astype(np.float32)
    opt = ol.ADAMW(1e-3, (0.9, 0.999), np.float32(0.001))
    n_pad = (n + 3) // 4 * 4
    tx = torch.zeros(n_pad, device=DEV); tx[:n] = torch.from_numpy(x).to(DEV)
    m = torch.zeros(n_pad, device=DEV); v = torch.zeros(n_pad, device=DEV)
    xo = x.copy()
    for step in range(1, 6):
        g = rng.standard_normal(n).astype(np.float32) * (10.0 ** rng.integers(-3, 2))
        tg = torch.zeros(n_pad, device=DEV); tg[:n] = torch.from_numpy(g).to(DEV)
        ldeq.adamw_step(tx, tg, m, v, step, 1e-3, (0.9, 0.999), 1e-8, 1e-3)
        xo = opt.update("w", xo, g)
        got = tx[:n].cpu().numpy()
        assert np.abs(got - xo).max() <= 2e-7 * np.abs(xo).max()
    # unaligned tail path
    ldeq.adamw_step(tx[:n - 4 + 0], tg[:n - 4], m[:n - 4], v[:n - 4], 6)


def test_sample_statistics_and_formula(ldeq):
    n = 1 << 20
    mu = torch.randn(n // 16, 16, device=DEV)
    lv = 0.5 * torch.randn(n // 16, 16, device=DEV)
    z, eps = ldeq.sample_raw(mu, lv, seed=1234, offset=0)
    z2, eps2 = ldeq.sample_raw(mu, lv, seed=1234, offset=0)
    z3, eps3 = ldeq.sample_raw(mu, lv, seed=1235, offset=0)
    torch.cuda.synchronize()
    assert torch.equal(eps, eps2) and not torch.equal(eps, eps3)          # counter-based, reproducible
    e = eps.double().cpu().numpy().ravel()
    assert abs(e.mean()) < 5e-3 and abs(e.std() - 1) < 5e-3
    assert abs((e ** 3).mean()) < 2e-2 and abs((e ** 4).mean() - 3) < 5e-2
    assert abs(np.corrcoef(e[:-1], e[1:])[0, 1]) < 5e-3
    ref = ol.sample(mu.cpu().numpy(), lv.cpu().numpy(), eps.cpu().numpy())
    assert np.abs(z.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
    # gradient of the differentiable wrapper
    mu.requires_grad_(True); lv.requires_grad_(True)
    zz = ldeq.sample_reparam(mu, lv, 7, 0)
    zz.sum().backward()
    assert torch.allclose(mu.grad, torch.ones_like(mu))
