"""NCCL route of the gradient all-reduce through the C ABI (ldeq_comm_*): no torch.distributed anywhere -- this is what a
Julia host (one process per GPU) calls.  One-rank case on any GPU box; two ranks when the box has two GPUs."""
import os
import tempfile
import time

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def test_single_rank_communicator_and_error_paths(ldeq):
    h = ldeq.Handle(0)   # a private handle: the process-wide one keeps no communicator
    g = torch.arange(1000, dtype=torch.float32, device="cuda:0")
    with pytest.raises(ldeq.LdeqError):
        h.allreduce_grads(g)                       # no communicator yet
    uid = h.comm_unique_id()
    assert len(uid) == 128
    with pytest.raises(ldeq.LdeqError):
        h.comm_init(uid, 1, 1)                     # rank out of range
    h.comm_init(uid, 0, 1)
    with pytest.raises(ldeq.LdeqError):
        h.comm_init(uid, 0, 1)                     # already initialised
    h.allreduce_grads(g)
    torch.cuda.synchronize()
    assert torch.equal(g.cpu(), torch.arange(1000, dtype=torch.float32))
    h.comm_destroy()
    h.close()


def _worker(rank, world, idfile, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import latentdiffeq_jl_b200 as ldeq
    torch.cuda.set_device(rank)
    h = ldeq.Handle(rank)
    if rank == 0:
        uid = h.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)        # the id travels out of band (a file here; MPI / sockets in practice)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            assert time.time() - t0 < 120
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    h.comm_init(uid, rank, world)
    n = 503387                                      # the default GOKU-net's flat gradient
    g = torch.full((n,), float(rank + 1), device=f"cuda:{rank}") + torch.arange(n, device=f"cuda:{rank}") * 1e-6
    for _ in range(3):
        x = g.clone()
        h.allreduce_grads(x)
    torch.cuda.synchronize()
    want = sum(float(r + 1) for r in range(world)) + world * torch.arange(n, device=f"cuda:{rank}") * 1e-6
    err = float((x - want).abs().max())
    h.comm_destroy()
    h.close()
    q.put((rank, err))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
def test_two_rank_allreduce_through_the_c_abi():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        idfile = os.path.join(d, "nccl_id")
        procs = [ctx.Process(target=_worker, args=(r, world, idfile, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    assert all(err < 1e-5 for _, err in res)
