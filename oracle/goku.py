"""ctypes front-end of the C++/OpenMP GOKU oracle (``oracle/ldeq_oracle.cpp``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates reference
``diffeq_layer(::Decoder{<:GOKU}, ...)`` (``src/models/GOKU.jl:98-130``) on the CPU.

Array layout matches the product C ABI: ``z0`` is ``[B, z]`` (Julia ``(z,B)`` column-major),
``theta`` is ``[B, p]``, trajectories are ``[T, B, z]`` (Julia ``(z,B,T)``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, asdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libldeq_oracle.so")

PENDULUM = 0            # examples/pendulum_friction-less/pendulum.jl:19-26
PENDULUM_FRICTION = 1   # examples/pendulum_friction-less/pendulum.jl:65-74

RET_SUCCESS, RET_MAXITERS, RET_DTLESSTHANMIN, RET_UNSTABLE = 0, 1, 2, 3

# the diffeq struct's `solver` field (pendulum.jl:11,58 use Tsit5(); SURVEY.md 8(f)4 names DP5 / BS3 / RK4)
TSIT5, DP5, BS3, RK4 = 0, 1, 2, 3
# OrdinaryDiffEq's PI-controller defaults per algorithm (beta2_default / beta1_default in alg_utils.jl [3P]):
# beta2 = 2/(5 order), beta1 = 7/(10 order), except DP5: beta2 = 4/100, beta1 = 1/order - 3 beta2/4
CONTROLLER_DEFAULTS = {TSIT5: (7.0 / 50.0, 2.0 / 25.0), DP5: (1.0 / 5.0 - 3.0 * 0.04 / 4.0, 0.04), BS3: (7.0 / 30.0, 2.0 / 15.0),
                       RK4: (7.0 / 40.0, 2.0 / 20.0)}


class _COpts(ctypes.Structure):
    _fields_ = [
        ("abstol", ctypes.c_double), ("reltol", ctypes.c_double), ("adaptive", ctypes.c_int),
        ("dt", ctypes.c_double), ("dtmax", ctypes.c_double), ("dtmin", ctypes.c_double),
        ("maxiters", ctypes.c_longlong), ("gamma", ctypes.c_double), ("qmin", ctypes.c_double),
        ("qmax", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double),
        ("qoldinit", ctypes.c_double), ("qsteady_min", ctypes.c_double), ("qsteady_max", ctypes.c_double),
        ("controller_pow", ctypes.c_int), ("solver", ctypes.c_int),
    ]


@dataclass
class Opts:
    """OrdinaryDiffEq ``solve`` keyword arguments that reach Tsit5 (SURVEY.md A.3 defaults)."""
    abstol: float = 1e-6
    reltol: float = 1e-3
    adaptive: bool = True
    dt: float = 0.0
    dtmax: float = 0.0
    dtmin: float = 0.0
    maxiters: int = 1_000_000
    gamma: float = 0.9
    qmin: float = 0.2
    qmax: float = 10.0
    beta1: float = 7.0 / 50.0
    beta2: float = 2.0 / 25.0
    qoldinit: float = 1e-4
    qsteady_min: float = 1.0
    qsteady_max: float = 1.0
    controller_pow: int = 0   # 0: DiffEqBase.fastpow (reference), 1: exact pow
    solver: int = 0           # TSIT5 / DP5 / BS3 / RK4

    @classmethod
    def for_solver(cls, solver: int, **kw) -> "Opts":
        """Options with the controller defaults OrdinaryDiffEq gives ``solver``."""
        b1, b2 = CONTROLLER_DEFAULTS[solver]
        return cls(solver=solver, **{"beta1": b1, "beta2": b2, **kw})

    def c(self) -> _COpts:
        d = asdict(self)
        d["adaptive"] = int(d["adaptive"])
        return _COpts(**d)


def build(force: bool = False) -> str:
    """Compile ``libldeq_oracle.so`` (g++); building the checker is not using it."""
    src = os.path.join(_HERE, "ldeq_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libldeq_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_fastpow.restype = ctypes.c_double
        _lib.oracle_fastpow.argtypes = [ctypes.c_double, ctypes.c_double]
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def fastpow(x: float, y: float) -> float:
    return float(lib().oracle_fastpow(x, y))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def solve(rhs: int, z0: np.ndarray, theta: np.ndarray, t: np.ndarray, opts: Opts | None = None,
          nthreads: int = 0):
    """B independent Tsit5 solves.  Returns ``(traj[T,B,z], retcode[B], naccept[B], nreject[B])``.

    dtype of ``z0`` selects the state precision (float32: Float32 state with Float64 time, the
    reference's mixed mode; float64: all Float64).
    """
    opts = opts or Opts()
    dt = z0.dtype
    assert dt in (np.float32, np.float64)
    z0 = np.ascontiguousarray(z0, dtype=dt)
    theta = np.ascontiguousarray(theta, dtype=dt)
    t = np.ascontiguousarray(t, dtype=np.float64)
    B, Z = z0.shape
    assert Z == 2 and theta.shape == (B, 1)
    T = t.shape[0]
    traj = np.empty((T, B, Z), dtype=dt)
    ret = np.empty(B, dtype=np.int32)
    na = np.empty(B, dtype=np.int32)
    nr = np.empty(B, dtype=np.int32)
    co = opts.c()
    fn = lib().oracle_goku_solve_f32 if dt == np.float32 else lib().oracle_goku_solve_f64
    fn(ctypes.c_int(rhs), _ptr(z0), _ptr(theta), _ptr(t), ctypes.c_int(B), ctypes.c_int(T), ctypes.byref(co),
       _ptr(traj), _ptr(ret), _ptr(na), _ptr(nr), ctypes.c_int(nthreads))
    return traj, ret, na, nr


def grad(rhs: int, z0: np.ndarray, theta: np.ndarray, t: np.ndarray, dtraj: np.ndarray,
         opts: Opts | None = None, norm_partials: bool = True, nthreads: int = 0):
    """Pullback of ``solve`` the way the reference computes it: ForwardDiffSensitivity, one dual
    solve seeded on ``p`` and one on ``u0`` per trajectory.

    ``norm_partials=True`` is ForwardDiff's behaviour (partials enter the error norm);
    ``False`` freezes the primal step sequence (exact derivative of the primal discretisation).
    Returns ``(dz0[B,z], dtheta[B,p])``.
    """
    opts = opts or Opts()
    dt = z0.dtype
    z0 = np.ascontiguousarray(z0, dtype=dt)
    theta = np.ascontiguousarray(theta, dtype=dt)
    t = np.ascontiguousarray(t, dtype=np.float64)
    dtraj = np.ascontiguousarray(dtraj, dtype=dt)
    B, Z = z0.shape
    T = t.shape[0]
    assert dtraj.shape == (T, B, Z)
    dz0 = np.empty((B, Z), dtype=dt)
    dth = np.empty((B, 1), dtype=dt)
    co = opts.c()
    fn = lib().oracle_goku_grad_f32 if dt == np.float32 else lib().oracle_goku_grad_f64
    fn(ctypes.c_int(rhs), _ptr(z0), _ptr(theta), _ptr(t), ctypes.c_int(B), ctypes.c_int(T), ctypes.byref(co),
       ctypes.c_int(int(norm_partials)), _ptr(dtraj), _ptr(dz0), _ptr(dth), ctypes.c_int(nthreads))
    return dz0, dth


def steps(rhs: int, z0, theta, t, opts: Opts | None = None, cap: int = 100000):
    """Accepted-step tape ``(t_n, dt_n)`` of one Float64 trajectory."""
    opts = opts or Opts()
    z0 = np.ascontiguousarray(z0, dtype=np.float64).reshape(2)
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(1)
    t = np.ascontiguousarray(t, dtype=np.float64)
    ts = np.empty(cap)
    dts = np.empty(cap)
    co = opts.c()
    lib().oracle_goku_steps_f64.restype = ctypes.c_int
    n = lib().oracle_goku_steps_f64(ctypes.c_int(rhs), _ptr(z0), _ptr(theta), _ptr(t), ctypes.c_int(t.shape[0]),
                                    ctypes.byref(co), _ptr(ts), _ptr(dts), ctypes.c_int(cap))
    return ts[:n].copy(), dts[:n].copy()


def method_table(solver: int):
    """``(ns, order, c[7], a[7,7], btilde[7])`` of a table-driven method as the oracle holds it."""
    c, a, bt = np.zeros(7), np.zeros((7, 7)), np.zeros(7)
    order = ctypes.c_int(0)
    ns = lib().oracle_method_table(ctypes.c_int(solver), _ptr(c), _ptr(a), _ptr(bt), ctypes.byref(order))
    return int(ns), int(order.value), c, a, bt


def dense_weights(solver: int, theta: float) -> np.ndarray:
    """Weights of ``k_1..k_ns`` in ``(u(theta) - u_n)/dt``, probed from the oracle's own dense-output routine."""
    w = np.zeros(7)
    lib().oracle_dense_weights(ctypes.c_int(solver), ctypes.c_double(theta), _ptr(w))
    return w


def jl_sincosf(x: np.ndarray):
    """``Base.sin`` / ``Base.cos`` of Float32 arguments as Julia 1.8 computes them (the restatement in
    ``ldeq_oracle.cpp``; every Float32 sine / cosine of the oracle goes through it)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().oracle_jl_sincosf(_ptr(x), _ptr(s), _ptr(c), ctypes.c_int(x.size))
    return s, c
