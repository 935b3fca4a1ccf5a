"""Timing of the persistent pattern-extractor kernels against the cuDNN route at one GPU's C5 share (B = 8192, T = 50)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import latentdiffeq_jl_b200 as ldeq
model_mod = __import__(ldeq.__name__ + ".model", fromlist=["x"]) if hasattr(ldeq, "__path__") else ldeq.model
DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
enc, dec = ldeq.default_layers(ldeq.GOKU(), 784, ldeq.Pendulum(), device=DEV)
model = ldeq.LatentDiffEqModel(ldeq.GOKU(), enc, dec)
fe = torch.randn(50, B, 32, device=DEV, requires_grad=True)
w = torch.randn(B, 48, device=DEV)
res = {"B": B, "T": 50}
for name, flag in (("persistent_kernels", True), ("cudnn", False)):
    model_mod.PERSISTENT_RECURRENT = flag
    def fwd():
        return ldeq.apply_pattern_extractor(model.encoder, fe)
    def fb():
        z0, th = fwd()
        (torch.cat([z0, th], 1) * w).sum().backward()
    for f, key in ((fwd, "fwd_ms"), (fb, "fwd_bwd_ms")):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        res.setdefault(name, {})[key] = e0.elapsed_time(e1) / 10
print(json.dumps(res))
