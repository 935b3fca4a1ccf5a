"""numpy restatement of the loss reduction, reparameterised sample and optimiser step -- TEST
INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows the reference:
  * ``kl`` / ``vector_kl``           ``src/utils/utils.jl:16-49``
  * ``loss_batch``                   ``examples/pendulum_friction-less/model_train.jl:225-238``
  * ``sample``                       ``src/models/GOKU.jl:155-163``
  * ``ADAMW(eta, beta, decay)``      ``model_train.jl:138`` = Flux 0.13 ``Optimiser(ADAM, WeightDecay)``
Arrays use the torch/numpy order ``[T, B, P]`` for Julia's ``(P, B, T)``; heads are ``[B, d]``.
"""
from __future__ import annotations

import numpy as np


def kl(mu, logvar):
    """utils.jl:16"""
    return (np.exp(logvar) + mu ** 2 - logvar - 1) / 2


def vector_kl(mu, logvar):
    """utils.jl:18-32 (tuple of heads) and :34-44 (single matrix): per head sum / batch size."""
    if isinstance(mu, (tuple, list)):
        return sum(np.float32(kl(m, lv).astype(np.float32).sum(dtype=np.float32) / np.float32(m.shape[0]))
                   for m, lv in zip(mu, logvar))
    return np.float32(kl(mu, logvar).astype(np.float32).sum(dtype=np.float32) / np.float32(mu.shape[0]))


def vector_mse(x, xhat):
    """utils.jl:5-13: x is a vector over time of (features, batch) matrices -> here [T, B, P]."""
    res = ((x - xhat) ** 2).sum()
    return res / (x.shape[0] * x.shape[1])


def loss_batch(x, xhat, mu, logvar, beta):
    """model_train.jl:231-237: sum(mean((x - xhat)^2, dims=(2,3))) + beta * vector_kl."""
    rec = ((x.astype(np.float64) - xhat.astype(np.float64)) ** 2).mean(axis=(0, 1)).sum()
    k = vector_kl(mu, logvar)
    return np.float32(rec) + np.float32(beta) * k, np.float32(rec), np.float32(k)


def loss_batch_grads(x, xhat, mu, logvar, beta):
    """Analytic gradients of ``loss_batch`` wrt xhat, mu heads and logvar heads."""
    T, B, _ = x.shape
    dxhat = (2.0 * (xhat.astype(np.float64) - x.astype(np.float64)) / (B * T)).astype(np.float32)
    mus = list(mu) if isinstance(mu, (tuple, list)) else [mu]
    lvs = list(logvar) if isinstance(logvar, (tuple, list)) else [logvar]
    dmu = [(beta * m / m.shape[0]).astype(np.float32) for m in mus]
    dlv = [(beta * 0.5 * (np.exp(lv) - 1) / lv.shape[0]).astype(np.float32) for lv in lvs]
    return dxhat, dmu, dlv


def sample(mu, logvar, eps):
    """GOKU.jl:159: mu + eps .* exp.(logvar / 2f0) in Float32."""
    return (mu + eps * np.exp(logvar / np.float32(2))).astype(np.float32)


class ADAMW:
    """Flux 0.13 ``ADAMW(eta, (b1, b2), decay)``: ``Optimiser(ADAM(eta, beta), WeightDecay(decay))``.

    ADAM keeps (b1, b2) and their running powers in Float64; moments are arrays of the parameter type
    (Float32); ``WeightDecay`` adds ``decay * x`` to the step (not scaled by eta); ``x .-= step``."""

    def __init__(self, eta=1e-3, beta=(0.9, 0.999), decay=np.float32(0.001), eps=1e-8):
        self.eta, self.beta, self.decay, self.eps = eta, beta, np.float32(decay), eps
        self.state = {}

    def update(self, key, x, g):
        if key not in self.state:
            self.state[key] = [np.zeros_like(x), np.zeros_like(x), np.array(self.beta, dtype=np.float64)]
        mt, vt, bp = self.state[key]
        b1, b2 = self.beta
        mt[...] = (b1 * mt.astype(np.float64) + (1 - b1) * g.astype(np.float64)).astype(np.float32)
        vt[...] = (b2 * vt.astype(np.float64) + (1 - b2) * g.astype(np.float64) ** 2).astype(np.float32)
        d = (mt.astype(np.float64) / (1 - bp[0]) / (np.sqrt(vt.astype(np.float64) / (1 - bp[1])) + self.eps) * self.eta)
        d = d.astype(np.float32)
        bp *= np.array(self.beta)
        d = d + self.decay * x
        x -= d
        return x
