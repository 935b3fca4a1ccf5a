"""Independent pins of ``oracle/recurrent.py`` (CPU tier): the Flux cells restated there against PyTorch's own
``nn.RNN`` / ``nn.LSTM`` modules (same gate order input, forget, cell, output; their second bias set to zero), the flat
``Flux.destructure`` layout against the host-side packer the product uses, and gradients against finite differences."""
import numpy as np
import torch

from oracle import recurrent as orr

H = 16


def _torch_stack(flat, lstm, F):
    layers = orr._split(torch.tensor(flat, dtype=torch.float64), lstm, F)
    mod = (torch.nn.LSTM if lstm else torch.nn.RNN)(F, H, num_layers=2, **({} if lstm else {"nonlinearity": "relu"})).double()
    with torch.no_grad():
        for li, l in enumerate(layers):
            getattr(mod, f"weight_ih_l{li}").copy_(l[0])
            getattr(mod, f"weight_hh_l{li}").copy_(l[1])
            getattr(mod, f"bias_ih_l{li}").copy_(l[2])
            getattr(mod, f"bias_hh_l{li}").zero_()
    return mod, layers


def test_cells_match_pytorch_modules():
    rng = np.random.default_rng(0)
    T, B, F = 9, 5, 32
    x = torch.tensor(rng.standard_normal((T, B, F)))
    for lstm in (False, True):
        flat = orr.init_params(lstm, F, rng)
        mod, layers = _torch_stack(flat, lstm, F)
        h0 = torch.stack([l[3].expand(B, H) for l in layers]).contiguous()
        for reverse in (False, True):
            xs = torch.flip(x, dims=[0]) if reverse else x
            if lstm:
                c0 = torch.stack([l[4].expand(B, H) for l in layers]).contiguous()
                out, _ = mod(xs, (h0, c0))
            else:
                out, _ = mod(xs, h0)
            got = orr.stack_final(x, torch.tensor(flat, dtype=torch.float64), lstm, reverse)
            assert torch.allclose(got, out[-1], rtol=1e-12, atol=1e-13)


def test_flat_layout_is_flux_destructure_order(ldeq):
    # the host-side packer (column-major Wi, Wh, b, state0 per layer) against the oracle's reader
    torch.manual_seed(1)
    layers = [ldeq.LSTM(32, H), ldeq.LSTM(H, H)]
    for l in layers:
        torch.nn.init.uniform_(l.h0, -1, 1)
        torch.nn.init.uniform_(l.c0, -1, 1)
    flat = ldeq.pe_flat_params(layers, True).detach()
    assert flat.numel() == orr.param_count(True, 32)
    back = orr._split(flat.double(), True, 32)
    for l, (wi, wh, b, h0, c0) in zip(layers, back):
        assert torch.equal(wi.float(), l.Wi.detach()) and torch.equal(wh.float(), l.Wh.detach()) and torch.equal(b.float(), l.b.detach())
        assert torch.equal(h0.float(), l.h0.detach()) and torch.equal(c0.float(), l.c0.detach())
    # column-major: element (row, j) of Wi sits at j * rows + row
    assert flat[3 * 64 + 7] == layers[0].Wi[7, 3]


def test_gradients_match_finite_differences():
    rng = np.random.default_rng(2)
    T, B, F = 6, 3, 16
    x = rng.standard_normal((T, B, F))
    rnn, lf, lb = (orr.init_params(False, F, rng).astype(np.float64), orr.init_params(True, F, rng).astype(np.float64),
                   orr.init_params(True, F, rng).astype(np.float64))
    dz0, dth = rng.standard_normal((B, H)), rng.standard_normal((B, 2 * H))
    _, _, g = orr.pattern_extractor(x, rnn, lf, lb, dz0, dth)

    def loss(xx, a, b, c):
        z0, th = orr.pattern_extractor(xx, a, b, c)
        return (z0 * dz0).sum() + (th * dth).sum()
    eps = 1e-6
    for which, arr, grad in ((0, x, g[0]), (1, rnn, g[1]), (2, lf, g[2]), (3, lb, g[3])):
        flat = arr.reshape(-1)
        for idx in rng.choice(flat.size, 12, replace=False):
            args = [x.copy(), rnn.copy(), lf.copy(), lb.copy()]
            args[which].reshape(-1)[idx] += eps
            lp = loss(*args)
            args[which].reshape(-1)[idx] -= 2 * eps
            lm = loss(*args)
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - grad.reshape(-1)[idx]) < 1e-6 * max(1.0, abs(fd)), (which, idx, fd, grad.reshape(-1)[idx])
