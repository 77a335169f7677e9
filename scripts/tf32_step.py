"""A few MLP_v1 training steps in one arithmetic mode at the benchmark batch (for ncu launch lists):
    python scripts/tf32_step.py [tf32|fp32|bf16] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climsim_b200 import MLPEngine                      # noqa: E402
from climsim_b200.synthetic import synthetic_batch      # noqa: E402
from climsim_b200.trainer import Trainer, glorot_uniform_flat  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "tf32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = 65536
eng = MLPEngine.mlp_v1(dtype=dtype, max_batch=B)
eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
tr = Trainer(eng, rule="adam_keras", lr=1e-3)
x, y = synthetic_batch(B, seed=1, device="cuda")
for _ in range(steps):
    loss = tr.step(x, y)
torch.cuda.synchronize()
print(dtype, "loss", loss)
