"""The user-defined ``diffeq`` structs of the reference's examples, same field names.

GOKU reads ``prob`` (``length(prob.u0)``, ``length(prob.p)``), ``solver``, ``sensealg`` and ``kwargs``
from the struct (reference ``src/models/GOKU.jl:105-108, 207-208``); LatentODE reads ``dudt``,
``solver``, ``neural_model``, ``augment_dim``, ``kwargs`` and ``latent_dim_in/out``
(``src/models/LatentODE.jl:62-66, 105-106``).  The structs here carry the same fields; ``prob.f``
names the right-hand side the CUDA integrator is instantiated with (a built-in kind, or CUDA C
source compiled by NVRTC for a user-defined system).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from . import _cabi


class Tsit5:
    """``OrdinaryDiffEq.Tsit5()`` -- the solver of the reference's diffeq structs (pendulum.jl:11,58).  The ``solver``
    field is the user's to set (it is handed to ``solve`` at GOKU.jl:121): :class:`DP5`, :class:`BS3` and :class:`RK4`
    are built for the GOKU path as well; any other object is refused, never silently replaced."""
    code = _cabi.SOLVER_TSIT5

    def __repr__(self):
        return "Tsit5()"


class DP5:
    """``OrdinaryDiffEq.DP5()``: Dormand-Prince 5(4), dense output in dopri5's ``contd5`` form (``csrc/ldeq_erk.cuh``)."""
    code = _cabi.SOLVER_DP5

    def __repr__(self):
        return "DP5()"


class BS3:
    """``OrdinaryDiffEq.BS3()``: Bogacki-Shampine 3(2), cubic Hermite dense output."""
    code = _cabi.SOLVER_BS3

    def __repr__(self):
        return "BS3()"


class RK4:
    """``OrdinaryDiffEq.RK4()``: the classical method, Hermite dense output.  Fixed step only (``adaptive=False, dt=...``):
    OrdinaryDiffEq's adaptive RK4 is a defect-control estimate that is not built (``LDEQ_ERR_UNSUPPORTED``)."""
    code = _cabi.SOLVER_RK4

    def __repr__(self):
        return "RK4()"


class ForwardDiffSensitivity:
    """Sensitivity request of the reference's diffeq structs (pendulum.jl:11): ``LDEQ_SENSE_FORWARD_DUAL`` -- the
    reference's own algorithm, two dual-number re-solves per trajectory whose error norm includes the partials
    (SciMLSensitivity 7.10 ``_concrete_solve_adjoint``).  Gradients equal the reference's to 1e-4."""
    code = _cabi.SENSE_FORWARD_DUAL

    def __repr__(self):
        return "ForwardDiffSensitivity()"


class DiscreteAdjoint:
    """Explicit opt-in (not a reference type): ``LDEQ_SENSE_DISCRETE_ADJOINT``, the reverse sweep over the taped accepted
    steps of the primal solve.  The exact derivative of the primal discretisation with frozen step sizes: identical to
    ``ForwardDiffSensitivity`` in fixed-step mode, within the solver tolerance of it otherwise (2e-2 at reltol 1e-3),
    at about a third of its cost."""
    code = _cabi.SENSE_DISCRETE_ADJOINT

    def __repr__(self):
        return "DiscreteAdjoint()"


class InterpolatingAdjoint:
    """DiffEqFlux's ``NeuralODE`` default for the LatentODE path: ``LDEQ_SENSE_INTERPOLATING_ADJOINT`` --
    the continuous adjoint ODE on ``[lambda; mu]`` solved backwards by adaptive Tsit5 with the forward dense output and a
    callback at every save time (``csrc/ldeq_mlp_cadj.cuh``).  The default of :class:`NODE` whenever the solve uses the
    reference's batch-global error norm and the exact arithmetic path; ``DiscreteAdjoint()`` is the ~30x cheaper opt-in
    (it agrees within the solver tolerance)."""
    code = _cabi.SENSE_INTERPOLATING_ADJOINT

    def __repr__(self):
        return "InterpolatingAdjoint()"


class ODEProblem:
    """``ODEProblem(f, u0, tspan, p)``.  ``f`` is a built-in RHS kind or a :class:`CudaRHS`."""

    def __init__(self, f, u0, tspan, p):
        self.f = f
        self.u0 = np.asarray(u0, dtype=np.float32)
        self.tspan = tuple(tspan)
        self.p = np.asarray(p, dtype=np.float32)


class CudaRHS:
    """A user-defined right-hand side as CUDA C source, compiled once per handle with NVRTC.

    ``source`` must define
    ``template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t);``
    """

    def __init__(self, source: str, z_dim: int, p_dim: int):
        self.source, self.z_dim, self.p_dim = source, int(z_dim), int(p_dim)

    def resolve(self, h: _cabi.Handle):
        return h.rhs_from_source(self.source, self.z_dim, self.p_dim)


class _GokuDiffEq:
    _kind = None

    def __init__(self, solver=None, sensalg=None, **kwargs):
        # Parameters and initial conditions only used to initialise the ODE problem (pendulum.jl:12-16)
        self.prob = ODEProblem(self._kind, [1.0, 1.0], (0.0, 1.0), [1.0])
        self.solver = solver or Tsit5()
        self.sensealg = sensalg or ForwardDiffSensitivity()
        self.kwargs = dict(kwargs)


class Pendulum(_GokuDiffEq):
    """Friction-less pendulum ``du = [y, -G/L sin x]``, G = 10 (pendulum.jl:4-46)."""
    _kind = _cabi.RHS_PENDULUM


class Pendulum_friction(_GokuDiffEq):
    """Pendulum with friction ``du = [y, -G/L sin x - (b/m) y]``, b = 0.7, m = 1 (pendulum.jl:51-91)."""
    _kind = _cabi.RHS_PENDULUM_FRICTION


class UserDiffEq(_GokuDiffEq):
    """A user-defined system for the GOKU path: same struct fields, RHS given as CUDA source."""

    def __init__(self, source: str, u0, p, solver=None, sensalg=None, **kwargs):
        self.prob = ODEProblem(CudaRHS(source, len(u0), len(p)), u0, (0.0, 1.0), p)
        self.solver = solver or Tsit5()
        self.sensealg = sensalg or ForwardDiffSensitivity()
        self.kwargs = dict(kwargs)


def glorot_uniform_(w: torch.Tensor):
    """Flux.glorot_uniform (the default ``Dense`` init, used by nODE.jl:14-16)."""
    fan_out, fan_in = w.shape
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        w.uniform_(-lim, lim)
    return w


class NODE(nn.Module):
    """Neural-ODE diffeq struct (nODE.jl:3-33): ``dudt = Chain(Dense(D+aug,H,relu), Dense(H,H,relu),
    Dense(H,D+aug))``, ``Tsit5()``, ``neural_model = NeuralODE``.

    Unlike the reference struct (which is not a ``@functor`` and therefore never trains, SURVEY.md
    Appendix C.2) the MLP weights are registered parameters here and receive gradients.
    """

    def __init__(self, latent_dim_in: int, hidden_dim: int = 200, augment_dim: int = 0, sensealg=None, **kwargs):
        super().__init__()
        d = latent_dim_in + augment_dim
        self.dims = [d, hidden_dim, hidden_dim, d]
        self.weights = nn.ParameterList()
        self.biases = nn.ParameterList()
        for i in range(3):
            w = torch.empty(self.dims[i + 1], self.dims[i])
            self.weights.append(nn.Parameter(glorot_uniform_(w)))
            self.biases.append(nn.Parameter(torch.zeros(self.dims[i + 1])))
        self.solver = Tsit5()
        self.neural_model = "NeuralODE"
        self.latent_dim_in = latent_dim_in
        self.latent_dim_out = latent_dim_in + augment_dim
        self.augment_dim = augment_dim
        self.kwargs = dict(kwargs)
        # NeuralODE's default sensitivity algorithm is InterpolatingAdjoint (DiffEqFlux 1.52); it needs the reference's
        # batch-global norm and exact arithmetic -- with the per-trajectory norm or the bf16x3 tensor-core path (both
        # documented performance deviations) the reverse pass is the discrete adjoint
        if sensealg is None:
            ref = self.kwargs.get("norm_mode", _cabi.NORM_GLOBAL) == _cabi.NORM_GLOBAL and self.kwargs.get("mlp_math", 0) == 0
            sensealg = InterpolatingAdjoint() if ref else DiscreteAdjoint()
        self.sensealg = sensealg

    @property
    def dudt(self):
        return list(zip(self.weights, self.biases))

    def flat_params(self) -> torch.Tensor:
        """``Flux.destructure(dudt)`` order: per layer ``vec(W)`` (column-major, W is (out,in)) then ``b``."""
        return torch.cat([torch.cat([w.t().reshape(-1), b.reshape(-1)]) for w, b in self.dudt])
