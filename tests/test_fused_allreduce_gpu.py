"""Two-GPU test of the all-reduce fused with AdamW over NVLink peer memory (skipped on a one-GPU box)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import latentdiffeq_jl_b200 as ldeq
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    lin_a = torch.nn.Sequential(torch.nn.Linear(301, 257), torch.nn.Linear(257, 3)).to(dev)   # odd sizes: tail path
    lin_b = torch.nn.Sequential(torch.nn.Linear(301, 257), torch.nn.Linear(257, 3)).to(dev)
    lin_b.load_state_dict(lin_a.state_dict())
    fa, fb = ldeq.FlatParams(lin_a, symmetric=True), ldeq.FlatParams(lin_b, symmetric=False)
    oa, ob = ldeq.ADAMW(fa), ldeq.ADAMW(fb)
    assert fa.symm is not None and fb.symm is None
    for step in range(4):
        # gradients bounded away from zero: Adam's update is ~ lr * sign(g) early on, so a sum that is zero up to rounding
        # would make the comparison depend on the summation order of the two all-reduce algorithms
        g = 1.0 + 0.25 * torch.randn(fa.grad.numel(), device=dev, generator=torch.Generator(device=dev).manual_seed(100 * step + rank))
        g = g * (1 - 2 * (torch.arange(g.numel(), device=dev) % 2))
        fa.grad.copy_(g)
        fb.grad.copy_(g)
        oa.fused_allreduce_step(grad_scale=1.0 / world)                  # ONE kernel over peer memory
        dist.all_reduce(fb.grad)                                          # baseline: NCCL all-reduce + AdamW kernel
        ob.step(grad_scale=1.0 / world)
    torch.cuda.synchronize()
    err = float((fa.flat - fb.flat).abs().max() / fb.flat.abs().max())
    gathered = [torch.zeros_like(fa.flat) for _ in range(world)]
    dist.all_gather(gathered, fa.flat)
    same = all(torch.equal(gathered[0], x) for x in gathered)            # rank-ordered sum: bit-identical replicas
    if rank == 0:
        print(f"world {world}: fused vs NCCL+AdamW rel err {err:.3e}, replicas identical {same}", flush=True)
        q.put((err, same))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_allreduce_adamw_matches_nccl_plus_adamw(world):
    # 2 ranks: one-shot kernel; 4 / 8 ranks: two-shot kernel (slice reduce + update, parameters written to every replica)
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, same = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert err < 1e-5 and same
