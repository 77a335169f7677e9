#!/usr/bin/env python
"""End-to-end run of the path this repository replaces, on synthetic columns (no dataset offline):

    python examples/train_mlp_v1.py [--columns 200000] [--batch 3072] [--epochs 3] [--workdir /tmp/climsim_b200_demo]

1. writes `train_input.npy` / `train_target.npy` / `val_*.npy` the way `data_utils.save_as_npy` does (climsim_utils/data_utils.py:884-921);
   the targets are a fixed smooth function of the inputs plus noise, so that there is something to learn;
2. trains MLP_v1 (baseline_models/MLP/.../hpo_baseline_v1.py:75-129: LeakyReLU(0.15), MSE, Keras Adam, tfa cyclical learning rate) with
   `NpyColumnStream` -> `Trainer.step` -- the role of `model.fit(tds, ...)` (step2_retrain.py:280-285);
3. predicts the validation split in slabs (`model.predict`, step3_inference.ipynb cell 2) and reports the MSE before / after.

Everything between the .npy files and the numbers printed runs on the GPU (one B200) through the C ABI."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from climsim_b200 import MLPEngine, NpyColumnStream
from climsim_b200.trainer import Trainer, cyclical_lr, glorot_uniform_flat


def make_split(path: str, name: str, n: int, seed: int, teacher) -> None:
    rng = np.random.default_rng(seed)
    x = (0.3 * rng.standard_normal((n, 124))).astype(np.float32)
    y = teacher(x) + (0.02 * rng.standard_normal((n, 128))).astype(np.float32)
    y[:, 120:] = np.abs(y[:, 120:])                           # the eight scalar targets are non-negative
    np.save(os.path.join(path, f"{name}_input.npy"), x)
    np.save(os.path.join(path, f"{name}_target.npy"), y.astype(np.float32))


def run(columns: int = 200_000, batch: int = 3072, epochs: int = 3, workdir: str = "/tmp/climsim_b200_demo", dtype: str = "bf16",
        verbose: bool = True) -> dict:
    os.makedirs(workdir, exist_ok=True)
    rng = np.random.default_rng(0)
    a, b = (rng.standard_normal((124, 64)) / np.sqrt(124)).astype(np.float32), (rng.standard_normal((64, 128)) / 8).astype(np.float32)
    teacher = lambda x: (np.tanh(x @ a) @ b).astype(np.float32)
    make_split(workdir, "train", columns, 1, teacher)
    make_split(workdir, "val", max(columns // 10, batch), 2, teacher)

    eng = MLPEngine.mlp_v1(dtype=dtype, max_batch=max(batch, 65536))
    eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
    stream = NpyColumnStream(os.path.join(workdir, "train_input.npy"), os.path.join(workdir, "train_target.npy"), batch, seed=0,
                             drop_last=True)
    steps_per_epoch = len(stream)
    trainer = Trainer(eng, rule="adam_keras", lr=lambda it: cyclical_lr(it, 2.5e-4, 2.5e-3, 2 * steps_per_epoch))

    xv = torch.from_numpy(np.load(os.path.join(workdir, "val_input.npy"))).cuda()
    yv = torch.from_numpy(np.load(os.path.join(workdir, "val_target.npy"))).cuda()

    def val_mse() -> float:
        se, n = 0.0, 0
        for i in range(0, xv.shape[0], 65536):                # model.predict in slabs
            p = eng.forward(xv[i:i + 65536])
            se += float(((p - yv[i:i + 65536]) ** 2).sum().item())
            n += p.numel()
        return se / n

    before = val_mse()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = []
    for e in range(epochs):
        run_loss = torch.zeros((), device="cuda")
        nb = 0
        for x, y in stream.epoch(e):
            run_loss += trainer.step(x, y, return_loss=False).reshape(())
            nb += 1
        losses.append(float(run_loss.item()) / max(nb, 1))
        if verbose:
            print(f"epoch {e}: mean training loss {losses[-1]:.6f}  ({nb} steps of {batch} columns)", flush=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    after = val_mse()
    res = {"val_mse_before": before, "val_mse_after": after, "epoch_losses": losses, "train_seconds": dt,
           "columns_per_s_incl_streaming": epochs * steps_per_epoch * batch / dt}
    if verbose:
        print(json.dumps(res))
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=200_000)
    ap.add_argument("--batch", type=int, default=3072)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--workdir", default="/tmp/climsim_b200_demo")
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    run(a.columns, a.batch, a.epochs, a.workdir, a.dtype)
