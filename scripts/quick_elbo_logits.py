"""ELBO forward + gradient at one GPU's C5 share (B = 8192, T = 50, P = 784): plain (x-hat given) vs the sigmoid output layer
folded in (pre-activations given); algorithmic bytes = read x, read x-hat / a, write the gradient = 3 x 4 P B T."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import latentdiffeq_jl_b200 as ldeq
DEV = "cuda:0"
T, B, P = 50, 8192, 784
x = torch.rand(T, B, P, device=DEV); a = torch.randn(T, B, P, device=DEV)
mu = [torch.randn(B, 16, device=DEV) for _ in range(2)]; lv = [0.1 * torch.randn(B, 16, device=DEV) for _ in range(2)]
peak = 6555.2
try:
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
res = {"bytes": 3 * 4 * T * B * P, "hbm_peak_gbs": peak}
for name, kw, inp in (("plain", {}, torch.sigmoid(a)), ("logits", {"logits": True}, a)):
    for _ in range(3): ldeq.elbo_raw(x, inp, mu, lv, 0.5, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ldeq.elbo_raw(x, inp, mu, lv, 0.5, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    res[name] = {"ms": ms, "gbs": res["bytes"] / (ms * 1e-3) / 1e9, "frac": res["bytes"] / (ms * 1e-3) / 1e9 / peak}
# what the folded kernel replaces in a training step: sigmoid forward + plain ELBO + sigmoid backward
ar = a.clone().requires_grad_(True)
def unfused():
    ar.grad = None
    ldeq.elbo_loss(x, torch.sigmoid(ar), tuple(mu), tuple(lv), 0.5).backward()
def fused():
    ar.grad = None
    ldeq.elbo_loss(x, ar, tuple(mu), tuple(lv), 0.5, logits=True, unit_cotangent=True).backward()
for name, fn in (("step_unfused_sigmoid_elbo_backward", unfused), ("step_folded", fused)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    res[name + "_ms"] = e0.elapsed_time(e1) / 10
print(json.dumps(res))
