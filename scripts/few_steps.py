"""A few eager training steps at the benchmark size (for ncu captures): python scripts/few_steps.py [steps]"""
import os, sys
os.environ["CSB_NO_GRAPHS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climsim_b200 import MLPEngine
from climsim_b200.synthetic import synthetic_batch
from climsim_b200.trainer import Trainer, glorot_uniform_flat
B = 65536
eng = MLPEngine.mlp_v1(dtype="bf16", max_batch=B)
eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
tr = Trainer(eng, lr=1e-3)
x, y = synthetic_batch(B, seed=0, device="cuda")
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    tr.step(x, y, return_loss=False)
torch.cuda.synchronize()
