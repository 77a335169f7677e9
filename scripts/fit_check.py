"""Trainer.fit on a B200 with the column stream for training and validation data, checkpoints, CSV log (GPU box):
    python scripts/fit_check.py [workdir]
Also run by tests/test_stream_gpu.py::test_fit_with_streams."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def run(d: str, verbose: int = 2) -> dict:
    from climsim_b200 import MLPEngine, NpyColumnStream
    from climsim_b200.trainer import Trainer, glorot_uniform_flat
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((124, 128)) / 11).astype(np.float32)
    for name, n in (("train", 20000), ("val", 4000)):
        x = (0.3 * rng.standard_normal((n, 124))).astype(np.float32)
        np.save(f"{d}/{name}_input.npy", x)
        np.save(f"{d}/{name}_target.npy", np.tanh(x @ a).astype(np.float32))
    eng = MLPEngine.mlp_v1(dtype="bf16", max_batch=4096)
    eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
    tr = Trainer(eng, lr=1e-3)
    h = tr.fit(NpyColumnStream(f"{d}/train_input.npy", f"{d}/train_target.npy", 2048, drop_last=True), epochs=3,
               validation_data=NpyColumnStream(f"{d}/val_input.npy", f"{d}/val_target.npy", 4096, shuffle=False),
               checkpoint_best=f"{d}/best.ckpt", checkpoint_last=f"{d}/last.ckpt", csv_log=f"{d}/log.csv", early_stopping_patience=8,
               verbose=verbose)
    tr.load_checkpoint(f"{d}/last.ckpt")
    h["iteration"] = tr.iteration
    h["log_rows"] = len(open(f"{d}/log.csv").read().strip().splitlines())
    return h


if __name__ == "__main__":
    out = run(sys.argv[1] if len(sys.argv) > 1 else "/tmp/fitchk")
    print(out)
    assert out["val_loss"][-1] < out["val_loss"][0] and out["loss"][-1] < out["loss"][0]
    print("FIT_OK")
