"""Secondary benchmark (BASELINE.json configs[2]): CNN (baseline_models/CNN, 12 x (k3, 406) ResNet-1D) bf16 training step on
one B200, synthetic (B,60,6) -> (B,60,10) columns.  Prints one JSON line; FLOP model: SURVEY.md section 8d (4.75 GFLOP/column)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from climsim_b200 import CNNEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps, warmup = 20, 5
eng = CNNEngine(max_batch=B)
rng = np.random.default_rng(0)
parts = []
for shp in eng.shapes():
    if len(shp) == 1:
        parts.append(np.zeros(shp, np.float32))
    else:
        fan_in, fan_out = int(np.prod(shp[:-1])), int(shp[0] * shp[-1]) if len(shp) == 3 else int(shp[-1])
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        parts.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
eng.set_params_flat(np.concatenate([p.reshape(-1) for p in parts]))
xs = [torch.randn(B, 60, 6, device="cuda") * 0.5 for _ in range(2)]
ys = [torch.randn(B, 60, 10, device="cuda") * 0.3 for _ in range(2)]
for i in range(warmup):
    eng.train_step(xs[i % 2], ys[i % 2]); eng.apply_opt("adam_keras", lr=1e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = eng.launch_count
e0.record()
for i in range(steps):
    loss = eng.train_step(xs[i % 2], ys[i % 2]); eng.apply_opt("adam_keras", lr=1e-4)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
# algorithmic FLOPs / column: forward MACs of the unpadded network x2, x3 for fwd + dgrad + wgrad (first-layer dgrad skipped)
mac = 0
c = 6
for _ in range(12):
    mac += 60 * (3 * c * 406 + 3 * 406 * 406 + c * 406); c = 406
mac += 60 * (406 * 10 + 10 * 10)
flop_train = 3 * 2 * mac - 2 * 60 * (3 * 6 * 406 + 6 * 406)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1400.0
tf = flop_train * B / (ms * 1e-3) / 1e12
print(json.dumps({"metric": "columns/sec", "workload": "CNN ResNet-1D 12x(k3,406) bf16 train step (fwd + mae_adjusted + bwd + Keras-Adam)", "value": B / (ms * 1e-3),
                  "unit": "columns/s", "batch": B, "ms_per_step": ms, "flop_per_column_train": flop_train, "tflops": tf, "frac_of_sustained_peak": tf / peak,
                  "gpu_launches_per_step": (eng.launch_count - l0) / steps, "loss": float(loss.item())}))
