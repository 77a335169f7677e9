"""Secondary benchmark (BASELINE.json configs[4]): the deeper MLP variants on one B200, bf16, 131 072 columns per GPU (= 2^20 / 8):
HSR mean network (124 -> 4 x [1024, LayerNorm, ReLU] -> 128, torch-Adam), ED (124 -> 463 -> ... -> 5 -> ... -> 463 -> 128, Keras-Adam),
online MLP (557 -> 384 -> 1024 -> 640 -> 368, Huber).  One JSON line per variant; FLOPs = 2 * MACs of the unpadded layers x (fwd + dgrad
[all but the first layer] + wgrad)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from climsim_b200 import MLPEngine
from climsim_b200.trainer import Trainer, glorot_uniform_flat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
STEPS, WARM = 60, 10
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["bf16_tflops_sustained"] if os.path.exists(pk) else 1400.0


def run(name, in_dim, layers, rule, layernorm=None, loss="mse", head_relu_from=-1):
    eng = MLPEngine(in_dim, layers, head_relu_from=head_relu_from, dtype="bf16", loss=loss, max_batch=B, layernorm=layernorm)
    eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0, layernorm=eng.layernorm))
    tr = Trainer(eng, rule=rule, lr=1e-4)
    out_dim = layers[-1][0]
    xs = [0.3 * torch.randn(B, in_dim, device="cuda") for _ in range(2)]
    ys = [0.1 * torch.randn(B, out_dim, device="cuda") for _ in range(2)]
    for i in range(WARM):
        tr.step(xs[i % 2], ys[i % 2], return_loss=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count
    e0.record()
    for i in range(STEPS):
        tr.step(xs[i % 2], ys[i % 2], return_loss=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    macs = [k * n for k, n in eng.layer_dims]
    flop = 2 * (3 * sum(macs) - macs[0])
    tf = flop * B / (ms * 1e-3) / 1e12
    print(json.dumps({"variant": name, "metric": "columns/sec", "value": B / (ms * 1e-3), "unit": "columns/s", "batch": B, "ms_per_step": ms,
                      "flop_per_column_train": flop, "tflops": tf, "frac_of_sustained_peak": tf / PEAK,
                      "gpu_launches_per_step": (eng.launch_count - l0) / STEPS}), flush=True)
    eng.close()


if __name__ == "__main__":
    only = sys.argv[2] if len(sys.argv) > 2 else ""          # python scripts/bench_variants.py 131072 hsr   (one variant, e.g. under ncu)
    if only:
        STEPS, WARM = 4, 3
    if only in ("", "hsr"):
        run("HSR mean net 4x1024 LayerNorm", 124, [(1024, "relu", 0.0)] * 4 + [(128, "none", 0.0)], "adam_torch", layernorm=[True] * 4 + [False])
    d = 463
    w = [d, d, d // 2, d // 4, d // 8, d // 16, 5, d // 16, d // 8, d // 4, d // 2, d, d]
    if only in ("", "ed"):
        run("ED 463-5-463", 124, [(x, "relu", 0.0) for x in w] + [(128, "elu", 0.0)], "adam_keras")
    if only in ("", "online"):
        run("online MLP 557-[384,1024,640]-368 Huber", 557, [(384, "relu", 0.0), (1024, "relu", 0.0), (640, "relu", 0.0), (368, "none", 0.0)],
        "adam_torch", loss="huber", head_relu_from=360)
