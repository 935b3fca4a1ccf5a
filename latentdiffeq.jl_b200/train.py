"""The training step of the reference's example script on the CUDA path.

``loss_batch`` (``examples/pendulum_friction-less/model_train.jl:225-238``), the gradient step
``Flux.pullback`` + ``update!(ADAMW(...))`` (``model_train.jl:138,195-201``) and, for several GPUs of
one box, the data-parallel layout the north star asks for: every rank integrates its own slice of
the batch, the parameter gradient lives in ONE flat fp32 bucket that is all-reduced once per step
over NCCL, then every rank applies the same fused AdamW.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .solve import adamw_step, allreduce_adamw_step, elbo_loss


def loss_batch(model, x, t, beta, variational, fused_output: bool = True, unit_cotangent: bool = False, grad_scale: float = 1.0):
    """model_train.jl:225-238.  ``x`` is ``[T, B, P]``.  With ``fused_output`` (default) the sigmoid of the reconstructor's
    output layer (GOKU.jl:265-268) is evaluated inside the loss kernel (``ldeq_elbo_logits_fwd_bwd``): same loss, same
    gradients, no x-hat array.  ``unit_cotangent``: the caller differentiates the returned loss directly (``grad_scale``: constant
    factor the kernel applies to the gradients, see ``elbo_loss``)."""
    if not x.is_cuda:
        raise RuntimeError("loss_batch runs in libldeq.so on a CUDA device (no CPU fallback)")
    out = model.forward_logits(x, t, variational) if fused_output and hasattr(model, "forward_logits") else None
    if out is not None:
        (logits, z_hat, l_hat), mu, logvar = out
        return elbo_loss(x, logits, mu, logvar, beta, logits=True, unit_cotangent=unit_cotangent, grad_scale=grad_scale)
    X_hat, mu, logvar = model(x, t, variational)
    x_hat, z_hat, l_hat = X_hat
    return elbo_loss(x, x_hat, mu, logvar, beta, unit_cotangent=unit_cotangent, grad_scale=grad_scale)


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous batch slice of ``rank``: trajectories are independent (GOKU.jl:111), so the batch is
    simply cut into ``world`` pieces; remainders go to the first ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatParams:
    """All trainable parameters of a model viewed through ONE contiguous fp32 buffer (and one gradient
    buffer): a single all-reduce and a single fused AdamW launch per step."""

    def __init__(self, module: torch.nn.Module, symmetric: bool = False):
        """``symmetric=True`` (several CUDA ranks of one box): the gradient bucket is NVLink symmetric memory, every
        rank can read every other rank's bucket, and the optimiser step fuses the all-reduce (``ADAMW.step``)."""
        self.params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        n_pad = (n + 3) // 4 * 4
        dev = self.params[0].device
        self.n = n
        self.symm = self.symm_p = None
        if symmetric and dev.type == "cuda" and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            import torch.distributed._symmetric_memory as symm_mem
            self.grad = symm_mem.empty(n_pad, dtype=torch.float32, device=dev)
            self.grad.zero_()
            self.symm = symm_mem.rendezvous(self.grad, dist.group.WORLD)
            self.flat = symm_mem.empty(n_pad, dtype=torch.float32, device=dev)   # replicas are peer-writable (two-shot step)
            self.flat.zero_()
            self.symm_p = symm_mem.rendezvous(self.flat, dist.group.WORLD)
        else:
            self.flat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            self.grad = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += k

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:  # autograd may have replaced .grad
            k = p.numel()
            p.grad = self.grad[off:off + k].view_as(p)
            off += k


class ADAMW:
    """``ADAMW(eta, (b1, b2), decay)`` of the reference (model_train.jl:138) on a :class:`FlatParams`."""

    def __init__(self, flat: FlatParams, eta=1e-3, beta=(0.9, 0.999), decay=1e-3, eps=1e-8):
        self.flat, self.eta, self.beta, self.decay, self.eps = flat, eta, beta, decay, eps
        self.step_count = 0

    def step(self, grad_scale: float = 1.0):
        """Local AdamW on the (already all-reduced) flat gradient."""
        self.step_count += 1
        f = self.flat
        if f.flat.is_cuda:
            adamw_step(f.flat, f.grad, f.m, f.v, self.step_count, self.eta, self.beta, self.eps, self.decay, grad_scale)
        else:
            raise RuntimeError("the optimiser step runs in libldeq.so on a CUDA device (no CPU fallback)")

    def fused_allreduce_step(self, grad_scale: float = 1.0):
        """All-reduce + AdamW in ONE kernel over NVLink peer memory (needs ``FlatParams(..., symmetric=True)``):
        barrier (all buckets written) -> sum in rank order + AdamW -> barrier (every rank has read every bucket / written
        every replica before the next pass).  Two ranks: one-shot (each rank reduces everything).  More: two-shot (each
        rank reduces and updates its 1/N slice and stores the new parameters into every replica)."""
        f = self.flat
        assert f.symm is not None, "FlatParams was not built with symmetric=True"
        self.step_count += 1
        two_shot = f.symm.world_size > 2 and f.symm_p is not None
        f.symm.barrier(channel=0)
        allreduce_adamw_step(f.flat, list(f.symm.buffer_ptrs), f.m, f.v, self.step_count, self.eta, self.beta, self.eps,
                             self.decay, grad_scale, peer_param_ptrs=list(f.symm_p.buffer_ptrs) if two_shot else None,
                             rank=f.symm.rank)
        f.symm.barrier(channel=1)


def allreduce_grads(flat: FlatParams):
    """One sum all-reduce of the flat gradient bucket (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)


def train_step(model, flat: FlatParams, opt: ADAMW, x_local, t, beta, variational=True, global_batch=None, timers=None):
    """One data-parallel training step on this rank's slice ``x_local`` ``[T, B_local, P]``.

    The loss is a mean over the GLOBAL batch: the local loss is scaled by ``B_local / B_global`` before the
    backward pass, gradients are summed across ranks, and every rank applies the same AdamW update.
    ``timers`` (a dict) receives CUDA events at the phase boundaries: start, loss, backward, update."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    B_local = x_local.shape[1]
    B_global = global_batch or B_local * world
    def mark(name):
        if timers is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timers[name] = ev
    mark("start")
    flat.zero_grad()
    # the loss is a mean over the GLOBAL batch: the factor B_local / B_global rides inside the loss kernel's gradients
    # (grad_scale), the loss itself is differentiated with a unit cotangent -- no extra pass over the (P,B,T) gradient
    loss = loss_batch(model, x_local, t, beta, variational, unit_cotangent=True, grad_scale=B_local / B_global)
    mark("loss")
    loss.backward()
    mark("backward")
    if flat.symm is not None:
        opt.fused_allreduce_step()
    else:
        allreduce_grads(flat)
        opt.step()
    mark("update")
    return loss.detach()
