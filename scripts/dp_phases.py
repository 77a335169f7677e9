"""Phase timing of the peer-memory data-parallel kernel (csb_mlp_dp_step), run under torchrun:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dp_phases.py
Prints, per rank, the median over 40 steps of the kernel's phases (globaltimer stamps written by block 0)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from climsim_b200 import MLPEngine
from climsim_b200.synthetic import synthetic_batch
from climsim_b200.trainer import Trainer, glorot_uniform_flat

rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 65536
eng = MLPEngine.mlp_v1(dtype="bf16", max_batch=B)
eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
tr = Trainer(eng, lr=1e-3)
assert tr._peer, "peer exchange not active"
x, y = synthetic_batch(B, seed=rank, device="cuda")
rows = []
for it in range(50):
    tr.step(x, y, return_loss=False)
    if it >= 10:
        st = (C.c_uint64 * 6)()
        eng.lib.csb_mlp_dp_debug(eng._h, st)
        rows.append([st[i + 1] - st[i] for i in range(5)] + [st[5] - st[0]])
med = np.median(np.array(rows, dtype=np.float64), axis=0) / 1e3
names = ["partials->grad + grid barrier", "wait all ranks ready", "slice exchange + grid barrier", "wait all ranks done", "optimizer", "total"]
torch.distributed.barrier()
for r in range(torch.distributed.get_world_size()):
    if r == rank:
        print(f"rank {rank}: " + "  ".join(f"{n} {v:.1f} us" for n, v in zip(names, med)), flush=True)
    torch.distributed.barrier()
torch.distributed.destroy_process_group()
