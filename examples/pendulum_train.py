"""Training loop of the reference's pendulum examples on the B200 path.

    python examples/pendulum_train.py --model goku      [--epochs 3] [--batch-size 64]
    python examples/pendulum_train.py --model latentode

Mirrors examples/pendulum_friction-less/model_train.jl (GOKU-net, `Args` :27-62, loop :186-217) and
model_train_LatentODE.jl of the reference: 450 pendulum sequences of 100 frames (create_data.jl:17-23), 90 / 10 split, minibatches of
64 windows of 50 frames (`time_loader`), ADAMW(1e-3, (0.9, 0.999), 1e-3), cyclical KL annealing (`frange_cycle_linear`), validation
loss on the held-out sequences after every minibatch.  The dataset is synthetic (no Luxor rasteriser here): true pendulum angles from the
hot path itself, drawn as 28 x 28 Gaussian-blob frames (bench.py::synthetic_frames).  Everything on the hot path -- sample, the latent
ODE solve and its reverse pass, the ELBO and the AdamW step -- runs in libldeq.so; the encoder / decoder layers are stock PyTorch."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latentdiffeq_jl_b200 as ldeq  # noqa: E402
from bench import pendulum_inputs, synthetic_frames  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="goku", choices=["goku", "latentode"])
    ap.add_argument("--epochs", type=int, default=3)            # the reference trains 1500
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--seq-len", type=int, default=50)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--decay", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=333)
    ap.add_argument("--dt", type=float, default=0.05)
    ap.add_argument("--n-cycle", type=int, default=4)
    ap.add_argument("--ratio", type=float, default=0.9)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("this example needs a CUDA device: the hot path has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.manual_seed(args.seed)
    rng = np.random.default_rng(args.seed)

    # ---- data: 450 sequences x 100 frames (create_data.jl:17-23), 90 / 10 split (model_train.jl:115) ----
    n_obs, full_seq_len = 450, 100
    z0, th = pendulum_inputs(n_obs, seed=1)
    t_full = args.dt * np.arange(full_seq_len)
    with torch.no_grad():
        ang, _, _ = ldeq.goku_solve_raw(torch.from_numpy(z0).to(dev), torch.from_numpy(th).to(dev), t_full, ldeq.RHS_PENDULUM)
        frames = synthetic_frames(ang[..., 0], dev)              # [100, 450, 784]
    n_train = int(round(0.9 * n_obs))
    train_set, val_set = frames[:, :n_train], frames[:, n_train:]
    t_val = args.dt * np.arange(full_seq_len)

    # ---- model (default_layers, GOKU.jl:199-274 / LatentODE.jl:100-140) ----
    if args.model == "goku":
        mt, diffeq = ldeq.GOKU_basic(), ldeq.Pendulum()
    else:
        mt, diffeq = ldeq.LatentODE(), ldeq.NODE(16)
    enc, dec = ldeq.default_layers(mt, 784, diffeq, device=dev)
    model = ldeq.LatentDiffEqModel(mt, enc, dec)
    flat = ldeq.FlatParams(model)
    opt = ldeq.ADAMW(flat, args.lr, (0.9, 0.999), args.decay)
    schedule = ldeq.frange_cycle_linear(args.epochs, 0.0, 1.0, args.n_cycle, args.ratio)
    t = args.dt * np.arange(args.seq_len)
    n_batches = n_train // args.batch_size                        # partial = false

    best_val = float("inf")
    for epoch in range(args.epochs):
        beta = float(schedule[epoch])
        perm = rng.permutation(n_train)
        tic = time.perf_counter()
        for i in range(n_batches):
            idx = torch.from_numpy(perm[i * args.batch_size:(i + 1) * args.batch_size]).to(dev)
            x = ldeq.time_loader(train_set[:, idx], full_seq_len, args.seq_len, rng)
            loss = ldeq.train_step(model, flat, opt, x, t, beta, variational=True)
            with torch.no_grad():
                val_loss = float(ldeq.loss_batch(model, val_set, t_val, beta, False))
        torch.cuda.synchronize()
        dt_epoch = time.perf_counter() - tic
        best_val = min(best_val, val_loss)
        print(f"epoch {epoch + 1}/{args.epochs}  beta {beta:.3f}  loss {float(loss):.4f}  val_loss {val_loss:.4f}  "
              f"{n_batches * args.batch_size / dt_epoch:.0f} samples/s", flush=True)
    print(f"best validation loss {best_val:.4f}")


if __name__ == "__main__":
    main()
