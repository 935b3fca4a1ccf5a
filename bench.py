#!/usr/bin/env python
"""bench.py -- the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c1|c3]

Workload (``config.workload``): BASELINE.json configs[3], the C4 sweep at its largest size -- 2^20
independent pendulum trajectories x 200 save points, forward + adjoint, Float32 state / Float64 time --
per GPU (weak scaling: every rank integrates its own slice, no data-path collective).  A "step" is
one pass of the hot path over that batch: ``ldeq_solve_fwd`` (recording the tape) followed by
``ldeq_solve_bwd``.  ``value`` = trajectory-steps/s (B*(T-1) saved grid intervals per solve) with the
inputs resident in HBM; ``e2e`` = the same through the host-buffer C-ABI entry points
(``ldeq_solve_fwd_host`` / ``ldeq_solve_bwd_host``) with pinned host buffers, copies inside the
timed region.  ``--impl reference`` times the CPU oracle (the reference is pure Julia, which this
image does not have: the oracle port is the reference arm) on the host cores.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "latent trajectory-steps/sec (forward + adjoint)"
UNIT = "traj-steps/s"
WORKLOADS = {
    # name: (B per GPU, T, rhs kind, description)
    "c4": (1 << 20, 200, 0, "C4 sweep: 2^20 pendulum trajectories x 200 save points (t = 0:0.05:9.95), forward + adjoint, "
                            "adaptive Tsit5 abstol 1e-6 reltol 1e-3, fp32 state / fp64 time"),
    "c3": (1024, 50, 1, "C3: pendulum with friction, B = 1024, T = 50, forward + adjoint"),
    "c1": (64, 50, 0, "C1: friction-less pendulum tutorial shape, B = 64, T = 50, forward + adjoint"),
}
CPU_SAMPLE_B = 1 << 18  # bounded CPU sample of the c4 workload (same T, same distribution)


def pendulum_inputs(B, seed=333):
    """create_data.jl:19-22 distribution; PCG64 seed 333 (model_train.jl:42)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    z0 = np.stack([rng.uniform(-np.pi / 6, np.pi / 6, B), rng.uniform(-np.pi / 3, np.pi / 3, B)], 1).astype(np.float32)
    th = rng.uniform(1.0, 2.0, (B, 1)).astype(np.float32)
    return z0, th


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_cores() -> int:
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle_throughput(B, T, rhs, steps=1, warmup=0, nthreads=0):
    """The reference arm / cpu_baseline: the CPU oracle (C++/OpenMP restatement of the reference algorithm:
    per-trajectory Tsit5 under EnsembleThreads, gradients as ForwardDiffSensitivity computes them = 1 primal
    + 2 dual solves per trajectory) timed on the host cores.  Returns (traj-steps/s, ms per step, cores)."""
    from oracle import goku as og
    og.build()
    nthreads = nthreads or host_cores()
    z0, th = pendulum_inputs(B)
    t = 0.05 * np.arange(T)
    d = np.random.default_rng(334).standard_normal((T, B, 2)).astype(np.float32)
    for _ in range(warmup):
        og.solve(rhs, z0[:4096], th[:4096], t, nthreads=nthreads)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        og.solve(rhs, z0, th, t, nthreads=nthreads)
        og.grad(rhs, z0, th, t, d, norm_partials=True, nthreads=nthreads)
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    return B * (T - 1) / dt, dt * 1e3, nthreads


def run_reference(args):
    """`--impl reference`: rank 0 only; each step is a bounded sample of the workload on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, T, rhs, desc = WORKLOADS[args.workload]
    Bs = min(B, CPU_SAMPLE_B)
    steps = max(1, min(args.steps, 5))
    val, ms, cores = cpu_oracle_throughput(Bs, T, rhs, steps=steps, warmup=1)
    sample = f"{Bs} of {B} trajectories per step (same T={T}, same input distribution), forward + ForwardDiff-style gradient"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "trajectories_per_step": Bs, "save_points": T},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference algorithm (Julia unavailable in this image); "
                                 "optimistic stand-in for Julia+Zygote (no per-trajectory allocation or AD overhead)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import latentdiffeq_jl_b200 as ldeq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, rhs, desc = WORKLOADS[args.workload]
    K, W = args.steps, max(args.warmup, 3)

    z0n, thn = pendulum_inputs(B, seed=333 + rank)
    t = 0.05 * np.arange(T)
    z0 = torch.from_numpy(z0n).to(dev)
    th = torch.from_numpy(thn).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(334 + rank)
    dtraj = torch.randn(T, B, 2, device=dev, generator=gen)
    opts = ldeq.default_opts()
    h = ldeq.handle(local)

    def step(ev=None):
        if ev:
            ev[0].record()
        traj, _, tape = ldeq.goku_solve_raw(z0, th, t, rhs, opts, want_tape=True, want_stats=False)
        tape.p_dim = 1
        if ev:
            ev[1].record()
        g = ldeq.goku_bwd_raw(tape, dtraj)
        if ev:
            ev[2].record()
        tape.free()
        return traj, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = h.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(evs[i])
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = h.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    bwd_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / K
    value = world * B * (T - 1) / (ms_per_step * 1e-3)

    # the same step with the reference's OWN sensitivity algorithm (two dual-number re-solves per trajectory,
    # LDEQ_SENSE_FORWARD_DUAL) instead of the discrete adjoint: the parity mode, reported next to the headline
    ref_sem = None
    if rhs in (ldeq.RHS_PENDULUM, ldeq.RHS_PENDULUM_FRICTION):
        import copy
        opts_fd = copy.copy(opts)
        opts_fd.sensealg = ldeq.SENSE_FORWARD_DUAL

        def step_fd():
            _, _, tape = ldeq.goku_solve_raw(z0, th, t, rhs, opts_fd, want_tape=True, want_stats=False)
            ldeq.goku_bwd_raw(tape, dtraj)
            tape.free()
        Kf = max(2, min(K, 5))
        for _ in range(2):
            step_fd()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(Kf):
            step_fd()
        f1.record()
        barrier()
        fd_ms = f0.elapsed_time(f1) / Kf
        if world > 1:
            tt = torch.tensor([fd_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fd_ms = float(tt.item())
        ref_sem = {"sensealg": "LDEQ_SENSE_FORWARD_DUAL (ForwardDiffSensitivity restated: dual-number re-solves, partials in the "
                               "error norm)", "ms_per_step": fd_ms, "value": world * B * (T - 1) / (fd_ms * 1e-3), "unit": UNIT, "steps": Kf}

    # ---- end to end through the host-buffer C-ABI entry points (pinned host memory) -------------------
    hz0 = torch.from_numpy(z0n).pin_memory()
    hth = torch.from_numpy(thn).pin_memory()
    hd = torch.empty(T, B, 2, dtype=torch.float32).pin_memory()
    hd.copy_(dtraj)
    htraj = torch.empty(T, B, 2, dtype=torch.float32).pin_memory()
    hdz0 = torch.empty(B, 2, dtype=torch.float32).pin_memory()
    hdth = torch.empty(B, 1, dtype=torch.float32).pin_memory()

    def e2e_step(bufs=None, handle=None):
        z_, th_, d_, tr_, gz_, gth_ = bufs or (hz0, hth, hd, htraj, hdz0, hdth)
        _, tape = ldeq.goku_solve_host(z_, th_, t, rhs, opts, device=local, want_tape=True, out=tr_, handle=handle)
        ldeq.goku_bwd_host(tape, d_, gz_, gth_)
        tape.free()
        return float(gz_[0, 0])  # the step's result is read on the host

    Ke = max(4, min(K, 10))
    Ke += Ke % 2
    for _ in range(2):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    e2e_ms_single = (time.perf_counter() - w0) * 1e3 / Ke

    # The entry points return when their result is on the host, so one host thread keeps only one direction of the
    # PCIe link busy at a time (1.69 GB down after the forward kernel, 1.69 GB up before the adjoint).  Independent
    # batches are served by two host threads, each with its own handle, stream and pinned buffers: the download of
    # one overlaps the upload of the other.  Same calls, same work per step; the steps are split between the threads.
    import threading
    h2 = ldeq.Handle(local)
    bufs2 = tuple(torch.empty_like(b).pin_memory() for b in (hz0, hth, hd, htraj, hdz0, hdth))
    for dst, src in zip(bufs2[:3], (hz0, hth, hd)):
        dst.copy_(src)
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

    errors = []

    def worker(i, n, go):
        # thread 1 starts half a step late, so that one thread downloads while the other uploads
        try:
            if i:
                go.wait(120)
            with torch.cuda.stream(streams[i]):
                for k in range(n):
                    z_, th_, d_, tr_, gz_, gth_ = bufs2 if i else (hz0, hth, hd, htraj, hdz0, hdth)
                    _, tape = ldeq.goku_solve_host(z_, th_, t, rhs, opts, device=local, want_tape=True, out=tr_, handle=h2 if i else None)
                    if not i and k == 0:
                        go.set()
                    ldeq.goku_bwd_host(tape, d_, gz_, gth_)
                    tape.free()
                    float(gz_[0, 0])
        except BaseException as e:  # noqa: BLE001 -- re-raised on the main thread
            errors.append(e)
        finally:
            go.set()

    def run_pair(n):
        go = threading.Event()
        th = [threading.Thread(target=worker, args=(i, n, go)) for i in range(2)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if errors:
            raise errors[0]
    run_pair(1)
    barrier()
    w0 = time.perf_counter()
    run_pair(Ke // 2)
    barrier()
    e2e_ms_dual = (time.perf_counter() - w0) * 1e3 / Ke
    e2e_streams = 2 if e2e_ms_dual < e2e_ms_single else 1
    e2e_ms = min(e2e_ms_dual, e2e_ms_single)
    if world > 1:
        tt = torch.tensor([e2e_ms, e2e_ms_single], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_ms_single = float(tt[0].item()), float(tt[1].item())
    e2e_value = world * B * (T - 1) / (e2e_ms * 1e-3)
    h2d = (hz0.numel() + hth.numel() + hd.numel()) * 4
    d2h = (htraj.numel() + hdz0.numel() + hdth.numel()) * 4

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes: forward reads (z+p) s and writes T z s per trajectory; the adjoint
        # reads the cotangent T z s and writes (z+p) s.  Tape traffic is implementation overhead, not counted.
        alg_fwd = B * (12 + 8 * T)
        alg_bwd = B * (8 * T + 12)
        fwd_gbs = alg_fwd / (fwd_ms * 1e-3) / 1e9
        bwd_gbs = alg_bwd / (bwd_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "trajectories_per_gpu": B, "save_points": T,
                       "sharding": f"independent trajectories, {B} per GPU, no data-path collective",
                       "l2": "inputs larger than L2 (1.68 GB cotangent + 1.68 GB trajectories per step vs 126 MB L2)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": Ke, "host_threads": e2e_streams,
                    "single_thread": {"value": world * B * (T - 1) / (e2e_ms_single * 1e-3), "ms_per_step": e2e_ms_single},
                    "path": "ldeq_solve_fwd_host + ldeq_solve_bwd_host, pinned host buffers"},
            "gpu_launches": int(launches),
            "reference_semantics_gradient": ref_sem,
            "roofline": {"bound": "hbm", "kernel": "tsit5_fwd_kernel<PendulumRHS<float,0>,float,TAPE=1>",
                         "achieved": fwd_gbs, "peak": peak, "unit": "GB/s", "frac": fwd_gbs / peak,
                         "peak_source": peak_src, "launch_ms": fwd_ms, "algorithmic_bytes_per_launch": alg_fwd,
                         "traffic": profile_traffic("tsit5_fwd_kernel_tape"),
                         "note": "per-trajectory work is ~37 k issued instructions for 1.6 kB of output: the kernel is "
                                 "instruction-issue bound (ncu: 79% issue-active), not HBM bound"},
            "roofline_bwd": {"bound": "hbm", "kernel": "tsit5_bwd_kernel<PendulumRHS<float,0>,float>",
                             "achieved": bwd_gbs, "peak": peak, "unit": "GB/s", "frac": bwd_gbs / peak,
                             "launch_ms": bwd_ms, "algorithmic_bytes_per_launch": alg_bwd,
                             "traffic": profile_traffic("tsit5_bwd_kernel")},
        }
        if world == 1 and not args.no_cpu:
            Bs = min(B, CPU_SAMPLE_B)
            val, ms, cores = cpu_oracle_throughput(Bs, T, rhs, steps=1, warmup=1)
            line["cpu_baseline"] = {
                "value": val, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{Bs} of {B} trajectories (same T={T}), forward + ForwardDiff-style gradient "
                          f"(1 primal + 2 dual solves per trajectory), {ms:.0f} ms"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def synthetic_frames(theta_angle, device):
    """Synthetic 28x28 pendulum frames [T, B, 784] in [0, 1]: a Gaussian blob at the bob position
    19*(cos(pi/2 + x), sin(pi/2 + x)) px from the pivot (geometry of create_data.jl:27,67-101; Luxor is unavailable)."""
    import torch
    T, B = theta_angle.shape
    ys, xs = torch.meshgrid(torch.arange(28, device=device, dtype=torch.float32),
                            torch.arange(28, device=device, dtype=torch.float32), indexing="ij")
    cx = 13.5 + 9.5 * torch.cos(torch.pi / 2 + theta_angle)
    cy = 5.0 + 9.5 * torch.sin(torch.pi / 2 + theta_angle)
    d2 = (xs.reshape(1, 1, -1) - cx.unsqueeze(-1)) ** 2 + (ys.reshape(1, 1, -1) - cy.unsqueeze(-1)) ** 2
    return torch.exp(-d2 / (2 * 1.5 ** 2))


def run_training(args):
    """`--workload c5`: GOKU-net data-parallel training (BASELINE.json configs[4]): default GOKU architecture
    (GOKU.jl:199-274), global batch 65 536 pendulum sequences of 50 frames split over the ranks (strong scaling), one
    flat-bucket NCCL gradient all-reduce + fused AdamW per step.  Metric: training samples/s."""
    import torch
    import torch.distributed as dist

    import latentdiffeq_jl_b200 as ldeq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    GB, T = args.global_batch, 50
    lo, hi = ldeq.shard_bounds(GB, rank, world)
    B = hi - lo
    K, W = args.steps, max(args.warmup, 3)
    torch.manual_seed(333)                                   # same initial weights on every rank
    mt = ldeq.GOKU_basic()
    enc, dec = ldeq.default_layers(mt, 784, ldeq.Pendulum(), device=dev)
    model = ldeq.LatentDiffEqModel(mt, enc, dec)
    flat = ldeq.FlatParams(model, symmetric=(world > 1 and not args.nccl_allreduce))
    opt = ldeq.ADAMW(flat, 1e-3, (0.9, 0.999), 1e-3)
    # synthetic data: true pendulum angles from the hot path itself, rasterised on the device
    z0n, thn = pendulum_inputs(GB, seed=1)
    t = 0.05 * np.arange(T)
    with torch.no_grad():
        ang, _, _ = ldeq.goku_solve_raw(torch.from_numpy(z0n[lo:hi]).to(dev), torch.from_numpy(thn[lo:hi]).to(dev), t, 0)
        x = synthetic_frames(ang[..., 0], dev)
    h = ldeq.handle(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        ldeq.train_step(model, flat, opt, x, t, beta=0.5, variational=True, global_batch=GB)
    barrier()
    l0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = ldeq.train_step(model, flat, opt, x, t, beta=0.5, variational=True, global_batch=GB)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / K
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    if rank == 0:
        print(json.dumps({
            "metric": "GOKU-net training samples/sec", "value": GB / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5: GOKU-net pendulum data-parallel training, default architecture (503 387 params), "
                                   f"global batch {GB} x 50 frames of 28x28, {B} per GPU, encoder/decoder layers stock PyTorch fp32, "
                                   "solve + sample + ELBO + AdamW in libldeq.so; gradient: " +
                                   ("all-reduce fused with AdamW over NVLink peer memory (one kernel)" if flat.symm is not None
                                    else "one NCCL all-reduce of the 2.0 MB flat bucket + AdamW kernel" if world > 1 else "local")},
            "gpu_launches": int(h.launch_count() - l0), "loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


def _latentode_inputs(workload):
    from oracle import mlp as om   # weights only (glorot init of nODE.jl:14-16)

    B = 256 if workload == "c2" else 18944
    T, dims = 50, [16, 200, 200, 16]
    rng = np.random.Generator(np.random.PCG64(1))
    layers = [(om.glorot_uniform(rng, dims[i + 1], dims[i]), np.zeros(dims[i + 1], np.float32)) for i in range(3)]
    p = om.pack_params(layers).astype(np.float32)
    z = (0.5 * rng.standard_normal((B, 16))).astype(np.float32)
    d = rng.standard_normal((T, B, 16)).astype(np.float32)
    return B, T, dims, p, z, d, 0.05 * np.arange(T)


def latentode_cpu_oracle(workload, reps=3, max_B=2048):
    """The numpy restatement of the reference's LatentODE path (oracle/mlp.py: BLAS matmuls on the (D,B) state, the way
    Flux Dense layers run on the CPU), forward solve with the batch-global norm + discrete adjoint, on the host cores.
    Bounded sample: at most max_B trajectories of the workload."""
    from oracle import mlp as om

    B, T, dims, p, z, d, t = _latentode_inputs(workload)
    Bs = min(B, max_B)
    z, d = z[:Bs], d[:, :Bs]
    try:  # numpy's BLAS reads OMP_NUM_THREADS (1 under torchrun) at import: give it the host cores
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=host_cores())
    except Exception:
        pass
    om.solve(z, p, dims, t, norm_mode="global")  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        om.solve(z, p, dims, t, norm_mode="global")
    fwd = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        _, _, _, tape = om.solve(z, p, dims, t, norm_mode="global", record=True)
        om.discrete_adjoint(p, dims, t, tape, d)
    both = (time.perf_counter() - t0) / reps
    return {"B": Bs, "fwd_ms": fwd * 1e3, "fwd_bwd_ms": both * 1e3, "traj_steps_per_s": Bs * (T - 1) / fwd,
            "traj_steps_per_s_fwd_bwd": Bs * (T - 1) / both, "cores": host_cores()}


def run_latentode_reference(args):
    """`--impl reference --workload c2|mlp`: the CPU oracle of the LatentODE path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    r = latentode_cpu_oracle(args.workload, reps=max(1, min(args.steps, 3)))
    B, T = (256 if args.workload == "c2" else 18944), 50
    line = {"impl": "reference", "metric": "latent trajectory-steps/sec (LatentODE forward solve)", "value": r["traj_steps_per_s"],
            "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": r["fwd_ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{'C2' if args.workload == 'c2' else 'MLP sweep'}: LatentODE, MLP RHS 16-200-200-16, batch {B}, T = {T}, "
                                   "adaptive Tsit5 abstol 1e-6 reltol 1e-3, forward solve", "trajectories_per_step": r["B"]},
            "cpu_baseline": {"value": r["traj_steps_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": f"{r['B']} of {B} trajectories per step, numpy/BLAS restatement of the reference algorithm "
                                       "(Julia unavailable in this image)", "fwd_bwd_ms": r["fwd_bwd_ms"]},
            "e2e": {"value": r["traj_steps_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_latentode(args):
    """`--workload c2` (BASELINE.json configs[1]: LatentODE, small-MLP RHS 16-200-200-16, batch 256, T = 50) and
    `--workload mlp` (the same network at batch 18 944 = 128 trajectories per SM: where batch x hidden is a genuine
    dense contraction).  Adaptive Tsit5, per-trajectory and global (reference) error norm, exact CUDA-core path vs the
    tcgen05 path; forward solve and forward + adjoint.  Unit: trajectory-steps/s (and accepted RK steps per trajectory
    per second)."""
    import torch

    import latentdiffeq_jl_b200 as ldeq

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B, T, dims, p_np, z_np, d_np, t = _latentode_inputs(args.workload)
    p, z, d = (torch.from_numpy(a).to(dev) for a in (p_np, z_np, d_np))
    K, W = args.steps, max(args.warmup, 3)
    peak_bf16 = 1634.1
    try:
        peak_bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    h = ldeq.handle(dev.index or 0)
    res = {}
    launches = 0
    for name, kw in (("tcgen05_bf16x3_global", dict(norm_mode=0, mlp_math=1)), ("tcgen05_bf16x3_per_traj", dict(norm_mode=1, mlp_math=1)),
                     ("exact_fp32_global", dict(norm_mode=0)), ("exact_fp32_per_traj", dict(norm_mode=1))):
        o = ldeq.default_opts(**kw)

        def fwd_bwd():
            tr, st, tape = ldeq.mlp_solve_raw(z, p, dims, t, o, want_tape=True)
            g = ldeq.mlp_bwd_raw(tape, d)
            tape.free()
            return st
        try:
            for _ in range(W):
                tr, st, _ = ldeq.mlp_solve_raw(z, p, dims, t, o)
                fwd_bwd()
        except ldeq.LdeqError as e:
            res[name] = {"unavailable": str(e)}
            continue
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        l0 = h.launch_count()
        e0.record()
        for _ in range(K):
            tr, st, _ = ldeq.mlp_solve_raw(z, p, dims, t, o)
        e1.record()
        for _ in range(K):
            fwd_bwd()
        e2.record()
        torch.cuda.synchronize()
        launches += h.launch_count() - l0
        ms, ms2 = e0.elapsed_time(e1) / K, e1.elapsed_time(e2) / K
        att = float((st.naccept + st.nreject).float().mean())
        rhs = B * (6 * att + 2)
        res[name] = {"ms": ms, "fwd_bwd_ms": ms2, "traj_steps_per_s": B * (T - 1) / (ms * 1e-3),
                     "traj_steps_per_s_fwd_bwd": B * (T - 1) / (ms2 * 1e-3), "rk_steps_per_traj_per_s": B * att / (ms * 1e-3),
                     "naccept_mean": float(st.naccept.float().mean()), "algorithmic_tflops": rhs * 92800 / (ms * 1e-3) / 1e12}
        if "tcgen05" in name:
            # tensor-pipe work actually issued: padded widths (208) and three bf16 passes per product
            issued = rhs * 3 * 2 * (16 * 208 + 208 * 208 + 208 * 16) / (ms * 1e-3) / 1e12
            res[name]["roofline"] = {"bound": "tensor", "achieved": issued, "peak": peak_bf16, "unit": "TFLOP/s",
                                     "frac": issued / peak_bf16, "traffic": None,
                                     "note": "bf16 flops issued to tcgen05 (3 passes, padded widths) / measured cuBLAS bf16 peak"}
    best = max((v["traj_steps_per_s"], k) for k, v in res.items() if "ms" in v)
    line = {"metric": "latent trajectory-steps/sec (LatentODE forward solve)", "value": best[0], "unit": UNIT, "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": res[best[1]]["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{'C2' if args.workload == 'c2' else 'MLP sweep'}: LatentODE, MLP RHS 16-200-200-16, batch {B}, T = 50, "
                                   "adaptive Tsit5 abstol 1e-6 reltol 1e-3, forward solve", "best": best[1]},
            "gpu_launches": launches, "variants": res}
    tc = [v["roofline"] for k, v in res.items() if "roofline" in v]
    if tc:
        line["roofline"] = max(tc, key=lambda r: r["frac"])
    if not args.no_cpu:
        r = latentode_cpu_oracle(args.workload)
        line["cpu_baseline"] = {"value": r["traj_steps_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": f"{r['B']} of {B} trajectories, numpy/BLAS restatement of the reference algorithm, forward solve "
                                          "with the batch-global norm", "fwd_ms": r["fwd_ms"], "fwd_bwd_ms": r["fwd_bwd_ms"]}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["c5", "c2", "mlp"])
    ap.add_argument("--global-batch", type=int, default=65536, help="c5 only")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--nccl-allreduce", action="store_true", help="c5: NCCL all-reduce + AdamW instead of the fused peer-memory kernel")
    args = ap.parse_args()
    if args.workload in ("c2", "mlp"):
        if args.impl == "reference":
            run_latentode_reference(args)
        else:
            run_latentode(args)
    elif args.workload == "c5":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "C5 is a training-loop workload: the reference arm "
                              "(CPU oracle) covers the hot path only (c4/c3/c1)"}))
        else:
            run_training(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
