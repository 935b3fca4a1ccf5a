"""Launches of the ELBO / AdamW / sample kernels at C5 shapes, with CUDA-event timing and achieved bandwidth against
their algorithmic bytes (ELBO 2*784*B*T*4 B read + the same size written for the gradient; AdamW 28 B
per parameter; sample 16 B per element incl. eps).  Run plain for the timing JSON, under ncu for the counters."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import latentdiffeq_jl_b200 as ldeq

dev = torch.device("cuda:0")
peak = 6555.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res = {}


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


# ELBO at one GPU's share of C5 on 8 GPUs: B = 8192 sequences x 50 frames x 784 pixels (1.28 GB per tensor: > L2)
B, T, P = 8192, 50, 784
x = torch.rand(T, B, P, device=dev)
xh = torch.rand(T, B, P, device=dev)
mus = [torch.randn(B, 16, device=dev) for _ in range(2)]
lvs = [0.1 * torch.randn(B, 16, device=dev) for _ in range(2)]
ms = timed(lambda: ldeq.elbo_raw(x, xh, mus, lvs, 0.5, want_grad=True), n=10)
alg = 3 * P * B * T * 4   # read x, xhat; write dxhat
res["elbo_fwd_bwd"] = {"ms": ms, "algorithmic_bytes": alg, "GBps": alg / ms / 1e6, "frac_of_hbm": alg / ms / 1e6 / peak,
                       "shape": [T, B, P]}
ms = timed(lambda: ldeq.elbo_raw(x, xh, mus, lvs, 0.5, want_grad=False), n=10)
alg = 2 * P * B * T * 4
res["elbo_fwd_only"] = {"ms": ms, "algorithmic_bytes": alg, "GBps": alg / ms / 1e6, "frac_of_hbm": alg / ms / 1e6 / peak}
del x, xh
# AdamW: the model's 503 387 parameters (launch-latency bound) and a 64 Mi-parameter bucket (bandwidth)
for n in (503388, 64 << 20):
    p_, g_, m_, v_ = (torch.randn(n, device=dev) for _ in range(4))
    m_.abs_(); v_.abs_()
    step = [0]

    def f():
        step[0] += 1
        ldeq.adamw_step(p_, g_, m_, v_, step[0])
    ms = timed(f)
    alg = 28 * n
    res[f"adamw_{n}"] = {"ms": ms, "algorithmic_bytes": alg, "GBps": alg / ms / 1e6, "frac_of_hbm": alg / ms / 1e6 / peak}
    del p_, g_, m_, v_
# sample: 16 x 65536 latent heads (C5 global batch) and a 64 Mi-element tensor
for shape in ((65536, 16), (1 << 22, 16)):
    mu = torch.randn(*shape, device=dev)
    lv = torch.randn(*shape, device=dev) * 0.1
    ms = timed(lambda: ldeq.sample_raw(mu, lv, 1, 0))
    n = mu.numel()
    alg = 16 * n  # read mu, logvar; write z, eps
    res[f"sample_{n}"] = {"ms": ms, "algorithmic_bytes": alg, "GBps": alg / ms / 1e6, "frac_of_hbm": alg / ms / 1e6 / peak}
print(json.dumps(res))
