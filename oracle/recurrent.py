"""CPU oracle of the recurrent pattern extractor (TEST INFRASTRUCTURE -- see ``oracle/__init__.py``).

Restates ``apply_pattern_extractor`` (reference ``src/models/GOKU.jl:30-49``, ``src/models/LatentODE.jl:20-34``) for the
default stacks (``GOKU.jl:224-234``): ``Chain(RNN(F,H,relu), RNN(H,H,relu))`` applied frame by frame to the REVERSED
sequence, ``Chain(LSTM(F,H), LSTM(H,H))`` on the sequence and a second one on the reversed sequence; only the last output
of each is kept and the hidden states restart from the trainable ``state0`` (``Flux.reset!``).  The cells follow Flux 0.13.6
``recurrent.jl`` [3P, restated]: ``RNNCell: h' = act.(Wi*x .+ Wh*h .+ b)``; ``LSTMCell: g = Wi*x .+ Wh*h .+ b``, gates
``input, forget, cell, output = sigm(g[1:o]), sigm(g[o+1:2o]), tanh(g[2o+1:3o]), sigm(g[3o+1:4o])``, ``c' = forget.*c .+
input.*cell``, ``h' = output.*tanh.(c')``.

Plain per-step loops in float64 torch on the CPU (the "plain reference" of a floating-point kernel); gradients come from
torch autograd through those loops.  Parameters are the flat ``Flux.destructure`` vectors the C ABI reads:
per layer ``Wi`` (rows x in, column-major), ``Wh`` (rows x H), ``b`` (rows), ``state0`` (H) [LSTM: ``h0``, ``c0``].
"""
from __future__ import annotations

import numpy as np
import torch

H = 16     # rnn_output_dim of GOKU's default stacks (GOKU.jl:201); LatentODE's RNN stack has 32 (LatentODE.jl:102)


def layer_sizes(lstm: bool, fan_in: int, H: int = H):
    rows = 4 * H if lstm else H
    return rows, [rows * fan_in, rows * H, rows, H] + ([H] if lstm else [])


def param_count(lstm: bool, F: int, H: int = H) -> int:
    return sum(layer_sizes(lstm, F, H)[1]) + sum(layer_sizes(lstm, H, H)[1])


def init_params(lstm: bool, F: int, rng: np.random.Generator, scale: float = 0.3, H: int = H) -> np.ndarray:
    """Random parameters in flat ``Flux.destructure`` order (forget-gate bias 1 like Flux's LSTMCell; random state0 so
    that their gradients are exercised)."""
    out = []
    for fan_in in (F, H):
        rows, sizes = layer_sizes(lstm, fan_in, H)
        wi = rng.uniform(-1, 1, sizes[0]) * scale
        wh = rng.uniform(-1, 1, sizes[1]) * scale
        b = rng.uniform(-0.1, 0.1, rows)
        if lstm:
            b[H:2 * H] += 1.0
        out += [wi, wh, b] + [rng.uniform(-0.5, 0.5, H) for _ in sizes[3:]]
    return np.concatenate(out).astype(np.float32)


def _split(flat: torch.Tensor, lstm: bool, F: int, H: int = H):
    layers, off = [], 0
    for fan_in in (F, H):
        rows, sizes = layer_sizes(lstm, fan_in, H)
        parts = []
        for n in sizes:
            parts.append(flat[off:off + n])
            off += n
        wi = parts[0].reshape(fan_in, rows).t()     # column-major (rows x in)
        wh = parts[1].reshape(H, rows).t()
        layers.append((wi, wh, parts[2]) + tuple(parts[3:]))
    assert off == flat.numel()
    return layers


def stack_final(x: torch.Tensor, flat: torch.Tensor, lstm: bool, reverse: bool, H: int = H) -> torch.Tensor:
    """Final hidden state ``[B, H]`` of one two-layer stack over ``x [T, B, F]``."""
    T, B, F = x.shape
    layers = _split(flat, lstm, F, H)
    hs = [l[3].expand(B, H) for l in layers]
    cs = [l[4].expand(B, H) for l in layers] if lstm else None
    order = range(T - 1, -1, -1) if reverse else range(T)
    for k in order:
        inp = x[k]
        for li, l in enumerate(layers):
            g = inp @ l[0].t() + hs[li] @ l[1].t() + l[2]
            if lstm:
                i, f = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H])
                cell, o = torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
                cs[li] = f * cs[li] + i * cell
                hs[li] = o * torch.tanh(cs[li])
            else:
                hs[li] = torch.relu(g)
            inp = hs[li]
    return hs[-1]


def pattern_extractor(x: np.ndarray, rnn: np.ndarray, lstm_f: np.ndarray | None = None, lstm_b: np.ndarray | None = None,
                      dz0: np.ndarray | None = None, dth: np.ndarray | None = None, H: int = H):
    """``(z0_out [B,H], theta_out [B,2H] or None)``; with cotangents also ``(dx, d_rnn, d_lstm_f, d_lstm_b)``."""
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=dz0 is not None)
    ps = [None if p is None else torch.tensor(p, dtype=torch.float64, requires_grad=dz0 is not None) for p in (rnn, lstm_f, lstm_b)]
    z0 = stack_final(xt, ps[0], False, True, H)
    th = None
    if lstm_f is not None:
        th = torch.cat([stack_final(xt, ps[1], True, False, H), stack_final(xt, ps[2], True, True, H)], dim=1)
    if dz0 is None:
        return z0.detach().numpy(), None if th is None else th.detach().numpy()
    loss = (z0 * torch.tensor(dz0, dtype=torch.float64)).sum()
    if th is not None:
        loss = loss + (th * torch.tensor(dth, dtype=torch.float64)).sum()
    loss.backward()
    grads = [xt.grad.numpy()] + [None if p is None else p.grad.numpy() for p in ps]
    return z0.detach().numpy(), None if th is None else th.detach().numpy(), grads
