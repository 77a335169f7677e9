"""Time the graph-replayed training step at the benchmark size and print ms/step (GPU box).  A/B of two builds on one box:
    for i in 1 2 3; do CSB_LIB_PATH=$PWD/climsim_b200/libclimsim_b200_old.so python scripts/time_step.py; python scripts/time_step.py; done"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climsim_b200 import MLPEngine
from climsim_b200.synthetic import synthetic_batch
from climsim_b200.trainer import Trainer, glorot_uniform_flat

B, STEPS = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 300
batches = [synthetic_batch(B, seed=i, device="cuda") for i in range(4)]
eng = MLPEngine.mlp_v1(dtype="bf16", max_batch=B)
eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
tr = Trainer(eng, lr=1e-3)
for it in range(20):
    tr.step(*batches[it % 4], return_loss=False)
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(STEPS):
        tr.step(*batches[it % 4], return_loss=False)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / STEPS)
print(os.environ.get("CSB_LIB_PATH", "current build"), " ".join(f"{t:.4f}" for t in res), "ms/step")
