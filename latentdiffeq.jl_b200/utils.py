"""The reference's exported utilities (``src/LatentDiffEq.jl:21-22``): ``vector_mse``, ``kl``,
``vector_kl``, ``frange_cycle_linear``, ``normalize_to_unit_segment``, ``time_loader``, ``rand_time``
(``src/utils/utils.jl``).  ``kl`` / ``vector_kl`` are plain torch here for API parity; the training
step uses the fused CUDA reduction ``elbo_loss`` instead (same value)."""
from __future__ import annotations

import numpy as np
import torch


def vector_mse(x, xhat):
    """utils.jl:5-13: sum of squared errors divided by (time steps * batch size); arrays ``[T, B, P]``."""
    return ((x - xhat) ** 2).sum() / (x.shape[0] * x.shape[1])


def kl(mu, logvar):
    """utils.jl:16"""
    return (torch.exp(logvar) + mu ** 2 - logvar - 1) / 2


def vector_kl(mu, logvar):
    """utils.jl:18-49: sum over heads of (sum of kl over the head) / batch size; heads are ``[B, d]``."""
    if isinstance(mu, (tuple, list)):
        return sum(kl(m, lv).sum() / m.shape[0] for m, lv in zip(mu, logvar))
    return kl(mu, logvar).sum() / mu.shape[0]


def frange_cycle_linear(n_iter, start=0.0, stop=1.0, n_cycle=4, ratio=0.5):
    """Cyclical KL-annealing schedule (utils.jl:53-67; 1-based indices of the reference mapped to 0-based)."""
    L = np.ones(n_iter, dtype=np.float64) * stop
    period = n_iter / n_cycle
    step = np.float32((stop - start) / (period * ratio))
    for c in range(n_cycle):
        v, i = np.float32(start), 1
        while v <= stop and _jl_round(i + c * period) < n_iter:
            L[_jl_round(i + c * period) - 1] = v
            v = np.float32(v + step)
            i += 1
    return L.astype(np.float32)


def _jl_round(x):
    """Julia's ``round`` (ties to even), as ``Int(round(x))``."""
    return int(np.round(x))


def normalize_to_unit_segment(X):
    """utils.jl:72-78"""
    mn, mx = X.min(), X.max()
    return (X - mn) / (mx - mn), mn, mx


def denormalize_unit_segment(Xh, mn, mx):
    return Xh * (mx - mn) + mn


def rand_time(full_seq_len, seq_len, rng=None):
    """utils.jl:96-100: ``rand(1:full_seq_len - seq_len)`` (the last possible window start is never drawn,
    SURVEY.md Appendix C.6); returns a 0-based slice."""
    rng = rng or np.random.default_rng()
    start = int(rng.integers(1, full_seq_len - seq_len + 1))
    return slice(start - 1, start - 1 + seq_len)


def time_loader(x, full_seq_len, seq_len, rng=None):
    """utils.jl:86-94: one random window of ``seq_len`` frames for the whole minibatch; ``x`` is ``[T, B, P]``."""
    return x[rand_time(full_seq_len, seq_len, rng)].float()
