import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ldeq():
    import latentdiffeq_jl_b200 as m
    return m


def pendulum_inputs(B, seed=333, dtype="float32"):
    """Synthetic pendulum inputs of the reference's data distribution
    (examples/pendulum_friction-less/create_data.jl:19-22): x0~U(+-pi/6), y0~U(+-pi/3), L~U(1,2)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    z0 = np.stack([rng.uniform(-np.pi / 6, np.pi / 6, B), rng.uniform(-np.pi / 3, np.pi / 3, B)], 1).astype(dtype)
    th = rng.uniform(1.0, 2.0, (B, 1)).astype(dtype)
    return z0, th
