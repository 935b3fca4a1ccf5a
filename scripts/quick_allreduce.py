"""Fused all-reduce+AdamW over peer memory vs NCCL all-reduce + AdamW kernel on the default GOKU gradient bucket.
   torchrun --nproc-per-node N scripts/quick_allreduce.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import latentdiffeq_jl_b200 as ldeq
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
class M(torch.nn.Module):
    def __init__(s): super().__init__(); s.w = torch.nn.Parameter(torch.randn(503387))
ma, mb = M().to(dev), M().to(dev)
fa, fb = ldeq.FlatParams(ma, symmetric=True), ldeq.FlatParams(mb)
oa, ob = ldeq.ADAMW(fa), ldeq.ADAMW(fb)
fa.grad.normal_(); fb.grad.normal_()
def fused(): oa.fused_allreduce_step(1.0 / world)
def nccl():
    dist.all_reduce(fb.grad); ob.step(1.0 / world)
for name, fn in (("fused peer-memory all-reduce+AdamW", fused), ("NCCL all-reduce + AdamW kernel", nccl)):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 200], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{world} GPUs, 503 387 fp32 grads: {name}: {t.item()*1e3:.1f} us/step")
dist.destroy_process_group()
