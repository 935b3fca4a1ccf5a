"""CPU tier: the N > 1 host logic on world_size 2 with the gloo backend -- batch sharding, loss scaling, ONE flat
gradient bucket all-reduced once per step, identical parameters on every rank afterwards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import latentdiffeq_jl_b200 as ldeq
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # the encoder side of the default GOKU model is stock torch and runs on the CPU; the solve does not
    mt = ldeq.GOKU_basic()
    enc, dec = ldeq.default_layers(mt, 36, ldeq.Pendulum(), hidden_dim_resnet=12, rnn_input_dim=6, rnn_output_dim=4,
                                   latent_dim_z0=4, latent_dim_theta=4, latent_to_diffeq_dim=8)
    encoder = ldeq.Encoder(mt, enc)
    flat = ldeq.FlatParams(encoder)
    B, T = 10, 7
    x = torch.rand(T, B, 36, generator=torch.Generator().manual_seed(1))
    lo, hi = ldeq.shard_bounds(B, rank, world)
    flat.zero_grad()
    mu, lv = encoder(x[:, lo:hi])
    loss_local = ldeq.vector_kl(mu, lv) * 0 + sum((m ** 2).sum() for m in mu) / (hi - lo)   # a mean over the local batch
    (loss_local * ((hi - lo) / B)).backward()
    ldeq.allreduce_grads(flat)
    g_dp = flat.grad.clone()
    # full-batch gradient on every rank for comparison
    flat.zero_grad()
    mu, lv = encoder(x)
    (sum((m ** 2).sum() for m in mu) / B).backward()
    err = float((g_dp - flat.grad).abs().max() / flat.grad.abs().max())
    gathered = [torch.zeros_like(g_dp) for _ in range(world)]
    dist.all_gather(gathered, g_dp)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        out.put((err, same, (lo, hi)))
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_full_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, same, cut = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5 and same and cut == (0, 5)
