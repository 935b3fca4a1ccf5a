"""Writes tests/golden/julia_inputs.bson: the seeded inputs of SURVEY.md 8(d) in a file BSON.jl can `BSON.load`, so that
julia/make_golden.jl (run ONCE by a maintainer who has Julia + the reference's environment) evaluates the REFERENCE on
exactly the arrays the oracle and the CUDA path are tested on.  numpy's PCG64 streams cannot be reproduced from Julia,
hence a file instead of seeds.

    python tests/golden/make_julia_inputs.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import pendulum_inputs  # noqa: E402

import latentdiffeq_jl_b200 as ldeq  # noqa: E402
from importlib import import_module  # noqa: E402

bson_io = import_module(ldeq.__name__ + ".bson_io")


def main():
    t = 0.05 * np.arange(50)
    # Julia arrays: (z, B), (p, B), (z, B, T) -- numpy arrays saved with those shapes (the writer stores column-major bytes)
    z0, th = pendulum_inputs(64, seed=333, dtype="float32")
    d = np.random.default_rng(334).standard_normal((50, 64, 2)).astype(np.float32)
    c1 = dict(z0=np.ascontiguousarray(z0.T), theta=np.ascontiguousarray(th.T), t=t, dtraj=np.ascontiguousarray(d.transpose(2, 1, 0)))
    z0, th = pendulum_inputs(1024, seed=333, dtype="float64")
    z0, th = z0[:64], th[:64]
    d = np.random.default_rng(335).standard_normal((50, 64, 2))
    c3 = dict(z0=np.ascontiguousarray(z0.T), theta=np.ascontiguousarray(th.T), t=t, dtraj=np.ascontiguousarray(d.transpose(2, 1, 0)))
    rng = np.random.default_rng(7)
    trig = np.concatenate([rng.uniform(-1, 1, 1024), rng.uniform(-8, 8, 2048), rng.uniform(-1e3, 1e3, 1024)]).astype(np.float32)
    fp_x = np.concatenate([10.0 ** rng.uniform(-8, 2, 512), [1e-4, 1.0, 0.5]])
    # the pattern-extractor fixture (tests/golden/pattern_extractor.npz): frames as the (F,B,T) array the feature extractor
    # returns, the three stacks' parameters as Flux.destructure vectors, the cotangents of the final states as (H,B) arrays
    g = np.load(os.path.join(HERE, "pattern_extractor.npz"))
    pe = dict(x=np.ascontiguousarray(g["x"].transpose(2, 1, 0)), rnn=g["rnn"], lstm_f=g["lstm_f"], lstm_b=g["lstm_b"], rnn32=g["rnn32"],
              dz0=np.ascontiguousarray(g["dz0"].T), dtheta=np.ascontiguousarray(g["dtheta"].T), dz0_32=np.ascontiguousarray(g["dz0_32"].T))
    bson_io.save(os.path.join(HERE, "julia_inputs.bson"), c1=c1, c3=c3, trig_x=trig, fastpow_x=fp_x,
                 fastpow_y=np.array([7 / 50, 2 / 25]), pe=pe)
    print(os.path.getsize(os.path.join(HERE, "julia_inputs.bson")), "bytes")


if __name__ == "__main__":
    main()
