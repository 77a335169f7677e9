"""A/B timing of the full training step under environment switches, interleaved in one process (GPU box):
    python scripts/ab_step.py CSB_NO_PDL CSB_NO_PAIRS ...      (each named variable is one variant, plus the baseline)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climsim_b200 import MLPEngine
from climsim_b200.synthetic import synthetic_batch
from climsim_b200.trainer import Trainer, glorot_uniform_flat

B, STEPS = 65536, 200
variants = ["baseline"] + sys.argv[1:]
batches = [synthetic_batch(B, seed=i, device="cuda") for i in range(4)]


def run(var):
    for v in sys.argv[1:]:
        os.environ.pop(v, None)
    if var != "baseline":
        os.environ[var] = "1"
    eng = MLPEngine.mlp_v1(dtype="bf16", max_batch=B)
    eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
    tr = Trainer(eng, lr=1e-3)
    for it in range(12):
        tr.step(*batches[it % 4], return_loss=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(STEPS):
        tr.step(*batches[it % 4], return_loss=False)
    e1.record()
    torch.cuda.synchronize()
    eng.close()
    return e0.elapsed_time(e1) / STEPS


os.system("nvidia-smi --query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm --format=csv")
res = {v: [] for v in variants}
for rep in range(4):
    for v in variants:
        res[v].append(run(v))
for v in variants:
    print(f"{v:16s} " + "  ".join(f"{t:.4f}" for t in res[v]) + f"   min {min(res[v]):.4f} ms", flush=True)
