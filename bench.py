#!/usr/bin/env python
"""bench.py -- columns/sec of the column-emulator training step on N B200s of one node, plus the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mlp_v1|cnn|hsr|ed] [--extras cnn,hsr,ed|none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --check            (parity of the benchmarked step against the oracle; under torchrun also the data-parallel check)

One JSON line on stdout (rank 0).  The headline (`--workload mlp_v1`, the default) is BASELINE.json configs[1] / [3]: MLP_v1 bf16,
65 536 columns per GPU per step, forward + MSE + backward + Keras-Adam.  `value` = whole-job columns/s with the inputs resident in
HBM (device-timed, CUDA events, max over ranks); `e2e` = the same step driven through the public Trainer.step() from pinned HOST
buffers (H2D of x,y and D2H of the loss inside the timed region) with the H2D-only rate of the same buffers beside it (the host-link
roofline); `roofline` = the dominant kernel kind; `cpu_baseline` = the CPU oracle on a bounded sample of the same workload on this
box's host cores.  The other BASELINE configurations ride in the same line under `workloads`: `cnn` (configs[2]: ResNet-1D 12 x
(k3, 406), B = 4096, Dropout 0.175, mae_adjusted), `hsr` and `ed` (configs[4]: 2^20 / 8 = 131 072 columns per GPU; HSR = BOTH
LayerNorm networks with the Gaussian-NLL loss and per-group L2 Adam as hpo.py:225-238 runs them), each with its own device-timed
value, step-level roofline, end-to-end number and CPU sample -- so that the driver's record covers every configuration.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNITS = (768, 640, 512, 640, 640)
IN_DIM, OUT_DIM = 124, 128
LAYER_DIMS = [(124, 768), (768, 640), (640, 512), (512, 640), (640, 640), (640, 128), (128, 128)]
# algorithmic FLOPs per column (SURVEY.md section 8d / BASELINE.md section 2): forward 3 500 032, training 10 309 632
FLOP_FWD_HIDDEN = 2 * sum(k * n for k, n in LAYER_DIMS[:-1])
FLOP_FWD_HEAD = 2 * 128 * 128
FLOP_DGRAD = 2 * sum(k * n for k, n in LAYER_DIMS[1:])
FLOP_WGRAD = 2 * sum(k * n for k, n in LAYER_DIMS)
FLOP_TRAIN = FLOP_FWD_HIDDEN + FLOP_FWD_HEAD + FLOP_DGRAD + FLOP_WGRAD
assert FLOP_FWD_HIDDEN + FLOP_FWD_HEAD == 3_500_032 and FLOP_TRAIN == 10_309_632
KIND_FLOPS = {"gemm_tn_fwd": FLOP_FWD_HIDDEN, "gemm_tn_head": FLOP_FWD_HEAD, "gemm_tn_dgrad": FLOP_DGRAD, "gemm_nt_wgrad": FLOP_WGRAD}

HSR_HIDDEN, HSR_LAYERS, HSR_GAMMA, HSR_LR = 1024, 4, 0.022, 7e-5        # baseline_models/HSR/training/hpo.py:225-238 (final configuration)
CNN_DEPTH, CNN_WIDTH, CNN_DROPOUT, CNN_BATCH = 12, 406, 0.175, 4096      # baseline_models/CNN/training/hpo_train.py:131-200,143
VARIANT_BATCH = (1 << 20) // 8                                           # BASELINE.json configs[4]: 2^20 columns over 8 GPUs


def dense_train_flops(dims) -> int:
    """2 x MACs x (forward + weight gradient + data gradient of every layer but the first), unpadded layer sizes."""
    macs = [k * n for k, n in dims]
    return 2 * (3 * sum(macs) - macs[0])


def ed_dims():
    d = 463
    w = [124, d, d, d // 2, d // 4, d // 8, d // 16, 5, d // 16, d // 8, d // 4, d // 2, d, d, 128]
    return list(zip(w[:-1], w[1:]))


def hsr_dims():
    w = [124] + [HSR_HIDDEN] * HSR_LAYERS + [128]
    return list(zip(w[:-1], w[1:]))


def cnn_train_flops() -> int:
    mac, c = 0, 6
    for _ in range(CNN_DEPTH):
        mac += 60 * (3 * c * CNN_WIDTH + 3 * CNN_WIDTH * CNN_WIDTH + c * CNN_WIDTH)
        c = CNN_WIDTH
    mac += 60 * (CNN_WIDTH * 10 + 10 * 10)
    return 3 * 2 * mac - 2 * 60 * (3 * 6 * CNN_WIDTH + 6 * CNN_WIDTH)      # no data gradient for the two convolutions that read x


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tf_sustained": d.get("bf16_tflops_sustained"), "tf_burst": d.get("bf16_tflops"), "hbm": d.get("hbm_gbs"), "src": "measured"}
    return {"tf_sustained": 1400.0, "tf_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


def choose_peak(peaks: dict, clocks: dict | None, timed_s: float):
    """Burst or sustained cuBLAS bf16 rate as the roofline denominator, decided by what THIS run saw: a timed region that ran under
    the power cap (sw_power_cap active, or the SM clock well below its maximum) is held to the sustained figure, a short or
    unthrottled one to the burst figure."""
    capped = False
    if clocks:
        capped = "sw_power_cap" in (clocks.get("reasons") or [])
        if clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] < 0.9 * clocks["sm_max_mhz"] and timed_s >= 0.4:
            capped = True
    if capped:
        return peaks["tf_sustained"], f"{peaks['src']} bf16_tflops_sustained (the timed region ran under the power cap)"
    return peaks["tf_burst"], f"{peaks['src']} bf16_tflops burst (no power capping seen during the timed region)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t_begin: float, t_end: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for ts, ln in self.lines if t_begin <= ts <= t_end + 0.25]
        window = "timed region"
        if len(inside) < 2:                       # timed region shorter than two sampling periods: include its own warm-up (the 2 s before)
            inside, window = [ln for ts, ln in self.lines if t_begin - 2.0 <= ts <= t_end + 0.25], "warm-up + timed region (timed region < 0.4 s)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()


def bind_to_gpu_cpus(local_rank: int) -> str:
    """Pin this process to the CPUs NVML reports as local to its GPU before any pinned buffer is allocated, so that the staging
    memory is first-touched on the GPU's NUMA node (the H2D copies then do not cross the socket interconnect)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs local to GPU {local_rank}"
    except Exception as e:                          # no NVML / not permitted: keep the inherited affinity
        return f"inherited ({type(e).__name__})"
    return "inherited"


# ------------------------------------------------------------------------------------------------------------------ CPU reference arm
def _time_cpu(step, steps: int, warmup: int) -> float:
    t0 = None
    for it in range(warmup + steps):
        if it == warmup:
            t0 = time.perf_counter()
        step(it)
    return time.perf_counter() - t0


def cpu_mlp_steps(batch: int, steps: int, warmup: int, threads: int, train: bool = True):
    """The CPU oracle (PyTorch fp32 restatement of the Keras MLP_v1 graph, Keras 'mse', Keras Adam) on `batch` synthetic columns per
    step.  Returns (columns/s, ms/step)."""
    from oracle import models as M          # bench.py may execute oracle/ only here: the CPU baseline / reference arm
    from climsim_b200.synthetic import synthetic_batch
    torch.set_num_threads(threads)
    ref = M.MLPRef(units=UNITS, seed=0)
    x, y = synthetic_batch(batch, 0)
    m = [torch.zeros_like(p) for p in ref.params]
    v = [torch.zeros_like(p) for p in ref.params]

    def step(it):
        if not train:
            with torch.no_grad():
                ref(x)
            return
        for p in ref.params:
            p.grad = None
        M.mse(y, ref(x)).backward()
        M.keras_adam_step(ref.params, [p.grad for p in ref.params], m, v, it + 1, M.cyclical_lr(it, step_size=2000))

    dt = _time_cpu(step, steps, warmup)
    return batch * steps / dt, 1e3 * dt / steps


def cpu_ed_steps(batch: int, steps: int, warmup: int, threads: int):
    from oracle import models as M
    from climsim_b200.synthetic import synthetic_batch
    torch.set_num_threads(threads)
    ref = M.EDRef(seed=0)
    x, y = synthetic_batch(batch, 0)
    m = [torch.zeros_like(p) for p in ref.params]
    v = [torch.zeros_like(p) for p in ref.params]

    def step(it):
        for p in ref.params:
            p.grad = None
        M.mse(y, ref(x)).backward()
        M.keras_adam_step(ref.params, [p.grad for p in ref.params], m, v, it + 1, 1e-4)

    dt = _time_cpu(step, steps, warmup)
    return batch * steps / dt, 1e3 * dt / steps


def cpu_hsr_steps(batch: int, steps: int, warmup: int, threads: int):
    """Both LayerNorm networks, Gaussian NLL, torch.optim.Adam with the per-group L2 decay: the reference's own loop body
    (hsr.py:100-140) on the oracle's restatement of its module (same state_dict, pinned against hsr.py)."""
    from oracle import models as M
    from climsim_b200.synthetic import synthetic_batch
    torch.set_num_threads(threads)
    net = M.HSRRef(124, 128, HSR_HIDDEN, HSR_LAYERS)
    alpha, beta = M.HSRRef.weight_decays(HSR_GAMMA)
    opt = torch.optim.Adam([{"params": net.mean.parameters(), "lr": HSR_LR, "weight_decay": alpha},
                            {"params": net.logprec.parameters(), "lr": HSR_LR, "weight_decay": beta}])
    x, y = synthetic_batch(batch, 0)

    def step(it):
        opt.zero_grad()
        mu, lp = net(x)
        M.HSRRef.loss(mu, lp, y, mle=True).backward()
        opt.step()

    dt = _time_cpu(step, steps, warmup)
    return batch * steps / dt, 1e3 * dt / steps


def cpu_cnn_steps(batch: int, steps: int, warmup: int, threads: int):
    from oracle import models as M
    torch.set_num_threads(threads)
    ref = M.CNNRef(depth=CNN_DEPTH, width=CNN_WIDTH, seed=0)
    g = torch.Generator().manual_seed(0)
    x, y = 0.5 * torch.randn(batch, 60, 6, generator=g), 0.3 * torch.randn(batch, 60, 10, generator=g)
    m = [torch.zeros_like(p) for p in ref.params]
    v = [torch.zeros_like(p) for p in ref.params]
    keep = 1.0 - CNN_DROPOUT

    def step(it):
        for p in ref.params:
            p.grad = None
        masks = [[(torch.rand(batch, 60, CNN_WIDTH, generator=g) < keep).float() / keep for _ in range(2)] for _ in range(CNN_DEPTH)]
        M.mae_adjusted(y, ref.forward(x, masks=masks)).backward()
        M.keras_adam_step(ref.params, [p.grad for p in ref.params], m, v, it + 1, 1e-4)

    dt = _time_cpu(step, steps, warmup)
    return batch * steps / dt, 1e3 * dt / steps


CPU_STEPS = {"mlp_v1": cpu_mlp_steps, "ed": cpu_ed_steps, "hsr": cpu_hsr_steps, "cnn": cpu_cnn_steps}
CPU_WHAT = {"mlp_v1": "PyTorch-CPU fp32 oracle of the Keras MLP_v1 train step (fwd + mse + bwd + Keras-Adam, cyclical LR)",
            "ed": "PyTorch-CPU fp32 oracle of the Keras encoder-decoder train step (fwd + mse + bwd + Keras-Adam 1e-4)",
            "hsr": "PyTorch-CPU fp32 restatement of hsr.py's two LayerNorm MLPs: fwd + Gaussian NLL + bwd + torch Adam with per-group L2",
            "cnn": "PyTorch-CPU fp32 oracle of the Keras ResNet-1D train step (Dropout .175, mae_adjusted, Keras-Adam 1e-4)"}
CPU_SAMPLE = {"mlp_v1": 4096, "ed": 4096, "hsr": 2048, "cnn": 32}           # columns per CPU step (bounded: seconds, not minutes)


def best_cpu_threads(workload: str, batch: int, max_threads: int) -> int:
    """PyTorch-CPU GEMMs of this size do not scale to every core of a big host (128 threads are ~10x slower than 16 on the GPU
    boxes): time one step at a few thread counts and keep the fastest, so the baseline is not sandbagged."""
    cands = sorted({c for c in (8, 16, 32, 64, max_threads) if c <= max_threads})
    best, best_cps = cands[0], 0.0
    for c in cands:
        cps, _ = CPU_STEPS[workload](batch, 2, 1, c)
        if cps > best_cps:
            best, best_cps = c, cps
    return best


def cpu_baseline(workload: str, sample: int, steps: int, warmup: int, max_threads: int) -> dict:
    threads = best_cpu_threads(workload, sample, max_threads)
    cps, ms = CPU_STEPS[workload](sample, steps, warmup, threads)
    return {"value": cps, "unit": "columns/s", "cores": threads, "kind": "port", "ms_per_step": ms,
            "sample": f"{steps} steps of {sample} columns of the same synthetic workload ({CPU_WHAT[workload]}), {ms:.0f} ms/step, "
                      f"{threads} of {max_threads} host threads (the fastest of 8/16/32/64/all)"}


def reference_arm(args, config_of) -> dict:
    """--impl reference: the CPU restatement of the reference's own path, timed on this box's host cores.  The Keras / TF original
    cannot run here (no tensorflow in the image, no network), so this is the oracle port; `config.columns_per_step` is what RAN."""
    threads_all = os.cpu_count() or 1
    wl = args.workload
    sample = args.cpu_sample or CPU_SAMPLE[wl]
    threads = best_cpu_threads(wl, sample, threads_all)
    cps, ms = CPU_STEPS[wl](sample, args.steps, args.warmup, threads)
    config = dict(config_of(wl, 1))
    config.update({"columns_per_step": sample, "columns_per_gpu_per_step": sample, "global_batch": sample, "parallelism": "cpu",
                   "note": f"the CPU arm steps on {sample} columns (a bounded sample of the GPU arm's per-step batch); one process, "
                           f"{threads} threads, whatever --gpus says"})
    line = {"impl": "reference", "metric": "columns/sec", "value": cps, "unit": "columns/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": cps, "unit": "columns/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps of {sample} columns ({CPU_WHAT[wl]})"},
            "e2e": {"value": cps, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if wl == "mlp_v1" and args.legs:
        # B = 1024 (configs[0], "1 CPU epoch" = 1000 steps; bounded here) and B = 3072 (the reference's best
        # batch size), forward-only and forward + backward + Adam, cores stated
        legs = {}
        for b in (1024, 3072):
            for train in (True, False):
                n = max(3, min(40 if train else 100, args.steps * 4))
                c, m = cpu_mlp_steps(b, n, 2, threads, train=train)
                legs[f"B{b}_{'train' if train else 'fwd'}"] = {"columns_per_s": c, "ms_per_step": m, "steps": n, "cores": threads}
        line["baseline_md_legs"] = legs
    if wl == "mlp_v1" and args.extras:
        line["workloads"] = {}
        for x in args.extras:
            n = {"cnn": 2, "hsr": 3, "ed": 5}[x]
            cb = cpu_baseline(x, CPU_SAMPLE[x], n, 1, threads_all)
            line["workloads"][x] = {"impl": "reference", "metric": "columns/sec", "value": cb["value"], "unit": "columns/s",
                                    "ms_per_step": cb["ms_per_step"], "config": config_of(x, 1) | {"columns_per_step": CPU_SAMPLE[x]},
                                    "cpu_baseline": cb}
    return line


# ------------------------------------------------------------------------------------------------------------------ GPU arm helpers
class Ctx:
    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.peaks = measured_peaks()
        self.sampler = None

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps: int):
        """EXACTLY `steps` calls of step(it) bracketed by barrier + synchronize; device time by CUDA events, max over ranks.
        Returns (ms_total, wall clock begin, end)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t_begin = time.time()
        ev0.record()
        for it in range(steps):
            step(it)
        ev1.record()
        self.barrier()
        t_end = time.time()
        return self.max_over_ranks(ev0.elapsed_time(ev1)), t_begin, t_end

    def clocks(self, t_begin, t_end):
        return self.sampler.window(t_begin, t_end) if (self.sampler is not None and self.rank == 0) else None

    def step_roofline(self, flop_per_col: int, cols_per_step_per_gpu: int, ms_per_step: float, clocks, timed_s: float, what: str) -> dict:
        peak, src = choose_peak(self.peaks, clocks, timed_s)
        tf = flop_per_col * cols_per_step_per_gpu / (ms_per_step * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": what, "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "peak_source": src,
                "traffic": None, "flop_per_column": flop_per_col}


def config_of(workload: str, world: int, batch: int | None = None) -> dict:
    if workload == "mlp_v1":
        b = batch or 65536
        return {"workload": "MLP_v1 (124->768->640->512->640->640->128->[120|8], LeakyReLU .15) train step: normalised x, "
                            "fwd + MSE + bwd + Keras-Adam, cyclical LR", "columns_per_gpu_per_step": b, "global_batch": b * world,
                "parallelism": f"dp{world}",
                "l2": "inputs rotate over 4 distinct device-resident batches (264 MB > 126 MB L2); ~1.3 GB of activations written/read per step"}
    if workload == "cnn":
        b = batch or CNN_BATCH
        return {"workload": f"CNN ResNet-1D {CNN_DEPTH} x (Conv1D k3 {CNN_WIDTH} -> ReLU -> Dropout {CNN_DROPOUT}) x 2 + 1x1 residual, per-level "
                            "Dense heads; train step: fwd + mae_adjusted + bwd + Keras-Adam 1e-4 (hpo_train.py:131-200)",
                "columns_per_gpu_per_step": b, "global_batch": b * world, "parallelism": f"dp{world}",
                "l2": "inputs rotate over 2 device-resident batches; ~7 GB of activations written/read per step (>> 126 MB L2)"}
    if workload == "hsr":
        b = batch or VARIANT_BATCH
        return {"workload": f"HSR: TWO LayerNorm MLPs 124 -> {HSR_LAYERS} x [{HSR_HIDDEN}, LayerNorm, ReLU] -> 128 (mean, log-precision), Gaussian "
                            f"NLL with clip, torch-Adam lr {HSR_LR} with per-group L2 (gamma {HSR_GAMMA}) (hsr.py:14-140, hpo.py:225-238)",
                "columns_per_gpu_per_step": b, "global_batch": b * world, "parallelism": f"dp{world}",
                "l2": "inputs rotate over 2 device-resident batches (264 MB > 126 MB L2); ~4 GB of activations per step"}
    b = batch or VARIANT_BATCH
    return {"workload": "ED encoder-decoder 124->463->463->231->115->57->28->5->28->...->463->128 (ReLU, ELU out), train step: fwd + MSE + bwd "
                        "+ Keras-Adam 1e-4 (ClimSIM_ED_1_3_train.py:56-96)", "columns_per_gpu_per_step": b, "global_batch": b * world,
            "parallelism": f"dp{world}", "l2": "inputs rotate over 2 device-resident batches (264 MB > 126 MB L2)"}


def e2e_block(value, world, B, row_bytes, steps, ms, wall_ms, loss, api, link=None) -> dict:
    d = {"value": value, "unit": "columns/s", "h2d_bytes_per_step": world * B * row_bytes, "d2h_bytes_per_step": world * 4, "steps": steps,
         "ms_per_step": ms, "wall_ms_per_step": wall_ms, "last_loss": loss, "api": api}
    if link:
        d.update(link)
    return d


def host_link(ctx: Ctx, copy_fn, bytes_per_step_per_gpu: int, steps: int, e2e_ms_per_step: float) -> dict:
    """H2D-only: the same pinned buffers through the same staging slots, nothing else in the loop -- the host-link roofline of the
    end-to-end number (per GPU, all ranks copying at once, max over ranks)."""
    for it in range(3):
        copy_fn(it)
    ms, _, _ = ctx.timed(copy_fn, steps)
    gbs = bytes_per_step_per_gpu * steps / (ms * 1e-3) / 1e9
    return {"host_link_gbs_per_gpu": gbs, "host_link_ms_per_step": ms / steps, "frac_of_host_link": (ms / steps) / e2e_ms_per_step,
            "host_link_note": "H2D-only pass over the same pinned buffers and staging slots (all ranks at once, max over ranks); "
                              "frac_of_host_link = H2D-only time / end-to-end time per step"}


# ------------------------------------------------------------------------------------------------------------------ MLP_v1 (headline)
def run_mlp_v1(ctx: Ctx) -> dict:
    from climsim_b200 import MLPEngine
    from climsim_b200.synthetic import synthetic_batch
    from climsim_b200.trainer import Trainer, cyclical_lr, glorot_uniform_flat
    args, rank, world = ctx.args, ctx.rank, ctx.world
    B = args.batch or 65536
    eng = MLPEngine.mlp_v1(units=UNITS, dtype=args.dtype, max_batch=B)
    eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))       # identical replicas on every rank
    trainer = Trainer(eng, rule="adam_keras", lr=lambda it: cyclical_lr(it, step_size=2000))
    batches = [synthetic_batch(B, seed=rank * 16 + i, device="cuda") for i in range(4)]

    # ---- device-resident timing (the `value`)
    for it in range(max(args.warmup, 8)):          # >= 8 so that each of the 4 rotating batches has its CUDA graph captured
        trainer.step(*batches[it % 4], return_loss=False)
    ctx.barrier()
    launches0 = eng.launch_count
    ms_total, t_begin, t_end = ctx.timed(lambda it: trainer.step(*batches[it % 4], return_loss=False), args.steps)
    launches = eng.launch_count - launches0
    clocks = ctx.clocks(t_begin, t_end)
    value = world * B * args.steps / (ms_total * 1e-3)
    ms_step = ms_total / args.steps

    # ---- per-kernel-kind device time: a second pass over the same steps with a CUDA event recorded on the launching stream after
    # every launch (the step then runs as individual, fully serialised launches instead of the cached CUDA graph with programmatic
    # dependent launch, so this pass is slower than the timed region; the kinds' SHARES of it are applied to the timed step)
    prof_steps = min(args.steps, 100)
    eng.profile(True)
    prof_ms_total, _, _ = ctx.timed(lambda it: trainer.step(*batches[it % 4], return_loss=False), prof_steps)
    prof = eng.profile_read()
    eng.profile(False)

    # ---- end to end from pinned host buffers: every step copies its own x,y (66 MB) and reads its own loss back; the engine's two
    # staging slots let the copy of step i+1 overlap the compute of step i (sync=False), as an input pipeline with prefetch does
    e2e_steps = max(3, min(args.steps, 50))
    host = [tuple(t_.cpu().pin_memory() for t_ in batches[i]) for i in range(4)]
    for it in range(4):
        trainer.step(*host[it % 4], sync=False)
    slots = []
    t0 = time.perf_counter()
    e2e_ms, _, _ = ctx.timed(lambda it: slots.append(trainer.step(*host[it % 4], sync=False)), e2e_steps)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    loss = float(slots[-1].item())
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)

    def copy_only(it):
        eng.stage_host_batch(*host[it % 4])
        eng.release_staged()

    link = host_link(ctx, copy_only, B * (IN_DIM + OUT_DIM) * 4, e2e_steps, e2e_ms / e2e_steps)

    line = None
    if rank == 0:
        timed_s = ms_total * 1e-3
        peak, peak_src = choose_peak(ctx.peaks, clocks, timed_s)
        kinds, tot = {}, sum(ms for ms, _ in prof.values())
        kind_flops = dict(KIND_FLOPS)
        if prof.get("gemm_tn_dgrad", (0, 0))[1] / prof_steps < 5.5:
            # fused tail (tail_kernel.cuh): the output layer's data gradient and weight gradient run inside the "head" launch
            tail = 2 * 128 * 128
            kind_flops["gemm_tn_head"] += 2 * tail; kind_flops["gemm_tn_dgrad"] -= tail; kind_flops["gemm_nt_wgrad"] -= tail
        for k, (ms, n) in prof.items():
            d = {"ms_per_step_serialised": ms / prof_steps, "launches_per_step": n / prof_steps, "share_of_step": ms / tot,
                 "ms_per_step": ms / tot * ms_step}
            if k in kind_flops:
                d["tflops"] = kind_flops[k] * B / (d["ms_per_step"] * 1e-3) / 1e12
                d["tflops_serialised"] = kind_flops[k] * B / (ms / prof_steps * 1e-3) / 1e12
            kinds[k] = d
        dom = max((k for k in kinds if k in KIND_FLOPS), key=lambda k: kinds[k]["ms_per_step"])
        n_dom = kinds[dom]["launches_per_step"]
        achieved = kinds[dom]["tflops"]
        traffic = None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath) and B == 65536:
                traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
                traffic_src = name
                break
        step_tf = FLOP_TRAIN * B / (ms_step * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": peak_src, "peak_burst": ctx.peaks["tf_burst"], "peak_sustained": ctx.peaks["tf_sustained"],
                    "achieved_note": "algorithmic FLOPs of the kind / (its share of the per-launch-event pass x the timed step): the timed "
                                     "region replays the step as ONE CUDA graph, inside which single launches cannot be event-timed",
                    "achieved_serialised_pass": kinds[dom]["tflops_serialised"],
                    "traffic": traffic, "traffic_note": f"bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum from profiles/{traffic_src} "
                                                        "(ncu --set full)" if traffic else None,
                    "flops_per_launch": kind_flops[dom] * B / max(n_dom, 1), "avg_launch_ms": kinds[dom]["ms_per_step"] / max(n_dom, 1),
                    "step_tflops": step_tf, "step_frac_of_peak": step_tf / peak, "step_frac_of_burst": step_tf / ctx.peaks["tf_burst"],
                    "step_frac_of_sustained": step_tf / ctx.peaks["tf_sustained"]}
        cpu = cpu_baseline("mlp_v1", args.cpu_sample or CPU_SAMPLE["mlp_v1"], 8, 2, os.cpu_count() or 1) if world == 1 else None
        line = {"metric": "columns/sec", "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic", "config": config_of("mlp_v1", world, B), "clocks": clocks,
                "e2e": e2e_block(e2e_value, world, B, (IN_DIM + OUT_DIM) * 4, e2e_steps, e2e_ms / e2e_steps, wall_ms / e2e_steps, loss,
                                 "climsim_b200.Trainer.step(x_pinned, y_pinned, sync=False) -> csb_mlp_train_step_host_async (N=1) / "
                                 "csb_mlp_stage_host_batch + train_step + all-reduce + apply_opt (N>1); per step: H2D of x,y into one of two "
                                 "staging slots on the copy stream (overlaps the previous step's compute), D2H of the loss into a pinned slot", link),
                "gpu_launches": launches, "roofline": roofline, "kernels": kinds,
                "kernels_note": f"per-kind shares from a second pass of {prof_steps} steps with per-launch CUDA events ({prof_ms_total / prof_steps:.4f} "
                                f"ms/step serialised); ms_per_step / tflops = share x the timed graph-replayed step ({ms_step:.4f} ms)",
                "cpu_baseline": cpu}
    eng.close()
    del batches, host
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------------------------ ED
def run_modes(ctx: Ctx) -> dict:
    """The MLP_v1 training step of the other two arithmetic modes at the benchmark batch, beside the bf16 headline (rank 0 of a
    single-GPU run): CSB_TF32 = the same tcgen05 kernels on fp32 storage with kind::tf32 products (the reference's A100 arithmetic),
    CSB_F32 = the FFMA parity engine.  Device-resident, CUDA events, a few steps each."""
    from climsim_b200 import MLPEngine
    from climsim_b200.synthetic import synthetic_batch
    from climsim_b200.trainer import Trainer, glorot_uniform_flat
    B = ctx.args.batch or 65536
    batches = [synthetic_batch(B, seed=100 + i, device="cuda") for i in range(2)]
    out = {}
    for dtype, steps in (("tf32", 10), ("tf32x3", 5), ("fp32", 3)):
        eng = MLPEngine.mlp_v1(units=UNITS, dtype=dtype, max_batch=B)
        eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
        tr = Trainer(eng, rule="adam_keras", lr=1e-3)
        for it in range(3):
            tr.step(*batches[it % 2], return_loss=False)
        ms, _, _ = ctx.timed(lambda it: tr.step(*batches[it % 2], return_loss=False), steps)
        loss = float(tr.step(*batches[0]))
        out[dtype] = {"ms_per_step": ms / steps, "columns_per_s": B * steps / (ms * 1e-3), "steps": steps, "last_loss": loss,
                      "step_tflops": FLOP_TRAIN * B / (ms / steps * 1e-3) / 1e12}
        eng.close()
    out["note"] = ("tf32: fp32 storage, every GEMM on the tcgen05 kernels with kind::tf32 (weight gradient = split contraction over "
                   "transposed operands); tf32x3: the same kernels over hi/lo-split operands, three products per fp32 product "
                   "(<= 3e-5 of the fp32 oracle); fp32: 64x64x16 FFMA tiles (the <= 1e-5 parity engine)")
    return out


def run_ed(ctx: Ctx, steps: int, warmup: int) -> dict:
    from climsim_b200 import MLPEngine
    from climsim_b200.synthetic import synthetic_batch
    from climsim_b200.trainer import Trainer, glorot_uniform_flat
    rank, world = ctx.rank, ctx.world
    B = ctx.args.batch or VARIANT_BATCH
    dims = ed_dims()
    layers = [(n, "relu", 0.0) for _, n in dims[:-1]] + [(128, "elu", 0.0)]
    eng = MLPEngine(124, layers, head_relu_from=-1, dtype=ctx.args.dtype, max_batch=B)
    eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))
    tr = Trainer(eng, rule="adam_keras", lr=1e-4)
    batches = [synthetic_batch(B, seed=rank * 16 + i, device="cuda") for i in range(2)]
    for it in range(max(warmup, 4)):
        tr.step(*batches[it % 2], return_loss=False)
    l0 = eng.launch_count
    ms, tb, te = ctx.timed(lambda it: tr.step(*batches[it % 2], return_loss=False), steps)
    launches = eng.launch_count - l0
    clocks = ctx.clocks(tb, te)
    host = [tuple(t_.cpu().pin_memory() for t_ in batches[i]) for i in range(2)]
    e2e_steps = max(3, min(steps, 10))
    for it in range(2):
        tr.step(*host[it % 2], sync=False)
    slots = []
    t0 = time.perf_counter()
    e2e_ms, _, _ = ctx.timed(lambda it: slots.append(tr.step(*host[it % 2], sync=False)), e2e_steps)
    wall = 1e3 * (time.perf_counter() - t0)
    loss = float(slots[-1].item())
    out = None
    if rank == 0:
        flop = dense_train_flops(dims)
        out = {"metric": "columns/sec", "value": world * B * steps / (ms * 1e-3), "unit": "columns/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": ms / steps, "dtype": ctx.args.dtype, "config": config_of("ed", world, B), "clocks": clocks, "gpu_launches": launches,
               "roofline": ctx.step_roofline(flop, B, ms / steps, clocks, ms * 1e-3, "whole step (28 GEMM launches on <= 463-wide layers + fused reduce/Adam)"),
               "e2e": e2e_block(world * B * e2e_steps / (e2e_ms * 1e-3), world, B, (IN_DIM + OUT_DIM) * 4, e2e_steps, e2e_ms / e2e_steps, wall / e2e_steps,
                                loss, "Trainer.step(x_pinned, y_pinned, sync=False)"),
               "cpu_baseline": cpu_baseline("ed", CPU_SAMPLE["ed"], 5, 1, os.cpu_count() or 1) if world == 1 else None}
    eng.close()
    del batches, host
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------ HSR
def run_hsr(ctx: Ctx, steps: int, warmup: int) -> dict:
    from climsim_b200 import MLPEngine, _lib
    from climsim_b200.engine import hsr_train_step
    from climsim_b200.synthetic import synthetic_batch
    from climsim_b200.trainer import glorot_uniform_flat
    rank, world = ctx.rank, ctx.world
    B = ctx.args.batch or VARIANT_BATCH
    spec = [(HSR_HIDDEN, "relu", 0.0)] * HSR_LAYERS + [(128, "none", 0.0)]
    ln = [True] * HSR_LAYERS + [False]
    nets = []
    for seed in (0, 1):
        e = MLPEngine(124, spec, dtype=ctx.args.dtype, max_batch=B, layernorm=ln)
        e.set_params_flat(glorot_uniform_flat(e.layer_dims, seed=seed, layernorm=ln))
        nets.append(e)
    mean, logprec = nets
    rho = 1 - HSR_GAMMA
    alpha, beta = (1 - rho) / rho * HSR_GAMMA, (1 - rho) / rho * (1 - HSR_GAMMA)
    scratch = torch.zeros(_lib.BATCH_METRICS_SCRATCH, dtype=torch.float64, device="cuda")
    loss_dev = torch.zeros(1, dtype=torch.float32, device="cuda")
    grads = [e.grad_buffer() for e in nets] if world > 1 else None
    batches = [synthetic_batch(B, seed=rank * 16 + i, device="cuda") for i in range(2)]

    def step_dev(x, y):
        hsr_train_step(mean, logprec, x, y, mle=True, loss_out=loss_dev, scratch=scratch, rule="adam_torch", lr=HSR_LR,
                       wd_mean=alpha, wd_logprec=beta, apply_opt=world == 1)
        if world > 1:                             # each rank's loss is the mean over ITS columns: average the gradients
            for g in grads:
                torch.distributed.all_reduce(g, op=torch.distributed.ReduceOp.AVG)
            mean.apply_opt("adam_torch", lr=HSR_LR, eps=1e-8, weight_decay=alpha)
            logprec.apply_opt("adam_torch", lr=HSR_LR, eps=1e-8, weight_decay=beta)

    for it in range(max(warmup, 3)):
        step_dev(*batches[it % 2])
    l0 = mean.launch_count + logprec.launch_count
    ms, tb, te = ctx.timed(lambda it: step_dev(*batches[it % 2]), steps)
    launches = mean.launch_count + logprec.launch_count - l0 + 2 * steps       # + the loss and gradient kernels of csb_hsr_train_step
    clocks = ctx.clocks(tb, te)
    # end to end: x, y from pinned host memory through the mean engine's staging slots, the loss read back every step
    host = [tuple(t_.cpu().pin_memory() for t_ in batches[i]) for i in range(2)]
    loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(steps, 10))

    def step_host(it):
        x, y = mean.stage_host_batch(*host[it % 2])
        step_dev(x, y)
        mean.release_staged()
        loss_host[it % 4:it % 4 + 1].copy_(loss_dev, non_blocking=True)

    for it in range(2):
        step_host(it)
    t0 = time.perf_counter()
    e2e_ms, _, _ = ctx.timed(step_host, e2e_steps)
    wall = 1e3 * (time.perf_counter() - t0)
    loss = float(loss_host[(e2e_steps - 1) % 4].item())
    out = None
    if rank == 0:
        flop = 2 * dense_train_flops(hsr_dims())
        out = {"metric": "columns/sec", "value": world * B * steps / (ms * 1e-3), "unit": "columns/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": ms / steps, "dtype": ctx.args.dtype, "config": config_of("hsr", world, B), "clocks": clocks, "gpu_launches": launches,
               "roofline": ctx.step_roofline(flop, B, ms / steps, clocks, ms * 1e-3, "whole step of both networks (GEMMs + LayerNorm passes + NLL + fused reduce/Adam)"),
               "e2e": e2e_block(world * B * e2e_steps / (e2e_ms * 1e-3), world, B, (IN_DIM + OUT_DIM) * 4, e2e_steps, e2e_ms / e2e_steps, wall / e2e_steps,
                                loss, "csb_mlp_stage_host_batch -> csb_hsr_train_step (both networks, NLL, Adam) -> 4-byte D2H of the loss"),
               "cpu_baseline": cpu_baseline("hsr", CPU_SAMPLE["hsr"], 3, 1, os.cpu_count() or 1) if world == 1 else None}
    for e in nets:
        e.close()
    del batches, host
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------ CNN
def cnn_glorot_flat(eng, seed=0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    parts = []
    for shp in eng.shapes():
        if len(shp) == 1:
            parts.append(np.zeros(shp, np.float32))
        else:
            fan_in = int(np.prod(shp[:-1]))
            fan_out = int(shp[0] * shp[-1]) if len(shp) == 3 else int(shp[-1])
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            parts.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
    return np.concatenate([p.reshape(-1) for p in parts])


def run_cnn(ctx: Ctx, steps: int, warmup: int) -> dict:
    from climsim_b200 import CNNEngine
    rank, world = ctx.rank, ctx.world
    B = ctx.args.batch or CNN_BATCH
    eng = CNNEngine(depth=CNN_DEPTH, width=CNN_WIDTH, loss="mae", max_batch=B, dtype=ctx.args.dtype)
    eng.set_params_flat(cnn_glorot_flat(eng))
    if ctx.args.dtype == "bf16":
        eng.set_dropout(CNN_DROPOUT, seed=1)
    g = torch.Generator().manual_seed(100 + rank)
    xs = [(0.5 * torch.randn(B, 60, 6, generator=g)).cuda() for _ in range(2)]
    ys = [(0.3 * torch.randn(B, 60, 10, generator=g)).cuda() for _ in range(2)]
    grad = eng.grad_buffer() if world > 1 else None

    def step_dev(x, y):
        loss = eng.train_step(x, y)
        if world > 1:
            torch.distributed.all_reduce(grad, op=torch.distributed.ReduceOp.AVG)
        eng.apply_opt("adam_keras", lr=1e-4)
        return loss

    for it in range(max(warmup, 3)):
        step_dev(xs[it % 2], ys[it % 2])
    l0 = eng.launch_count
    ms, tb, te = ctx.timed(lambda it: step_dev(xs[it % 2], ys[it % 2]), steps)
    launches = eng.launch_count - l0
    clocks = ctx.clocks(tb, te)
    # end to end: (B,60,6) / (B,60,10) fp32 from pinned host memory into two rotating device slots on a copy stream
    hx = [t.cpu().pin_memory() for t in xs]
    hy = [t.cpu().pin_memory() for t in ys]
    copy_stream = torch.cuda.Stream()
    slots = [(torch.empty_like(xs[0]), torch.empty_like(ys[0])) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    used = [False, False]
    loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
    main = torch.cuda.current_stream()
    e2e_steps = max(3, min(steps, 6))

    def step_host(it):
        s = it & 1
        with torch.cuda.stream(copy_stream):
            if used[s]:
                copy_stream.wait_event(freed[s])
            slots[s][0].copy_(hx[s], non_blocking=True)
            slots[s][1].copy_(hy[s], non_blocking=True)
            copied[s].record(copy_stream)
        main.wait_event(copied[s])
        loss = step_dev(*slots[s])
        freed[s].record(main)
        used[s] = True
        loss_host[it % 4:it % 4 + 1].copy_(loss, non_blocking=True)

    for it in range(2):
        step_host(it)
    t0 = time.perf_counter()
    e2e_ms, _, _ = ctx.timed(step_host, e2e_steps)
    wall = 1e3 * (time.perf_counter() - t0)
    loss = float(loss_host[(e2e_steps - 1) % 4].item())
    out = None
    if rank == 0:
        out = {"metric": "columns/sec", "value": world * B * steps / (ms * 1e-3), "unit": "columns/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": ms / steps, "dtype": ctx.args.dtype, "config": config_of("cnn", world, B), "clocks": clocks, "gpu_launches": launches,
               "roofline": ctx.step_roofline(cnn_train_flops(), B, ms / steps, clocks, ms * 1e-3, "whole step (convolutions as row-shifted tcgen05 GEMMs)"),
               "e2e": e2e_block(world * B * e2e_steps / (e2e_ms * 1e-3), world, B, (360 + 600) * 4, e2e_steps, e2e_ms / e2e_steps, wall / e2e_steps, loss,
                                "pinned (B,60,6) / (B,60,10) -> two device slots on a copy stream -> csb_cnn_train_step + csb_cnn_apply_opt -> D2H of the loss"),
               "cpu_baseline": cpu_baseline("cnn", CPU_SAMPLE["cnn"], 2, 1, os.cpu_count() or 1) if world == 1 else None}
    eng.close()
    del xs, ys, slots
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------ --check
def run_check(ctx: Ctx) -> dict:
    """Parity of the benchmarked step itself.  (1) MLP_v1 bf16 at the benchmark batch: loss and every gradient tensor of step 0
    against the bf16-emulating CPU oracle.  (2) world > 1: `reduce partials -> ncclAllReduce -> apply_opt` on N ranks equals ONE
    engine stepping on the joined batch -- fp32 engines (gradients to 2e-4 of the largest entry: each rank sums ITS rows first, so
    the fp32 summation order over the batch differs from the single engine's and the result is not bit for bit; 6e-5 measured) and
    bf16 engines (relative L2, through the peer-memory kernel csb_mlp_dp_step)."""
    from climsim_b200 import MLPEngine
    from climsim_b200.synthetic import synthetic_batch
    from climsim_b200.trainer import Trainer, glorot_uniform_flat
    from oracle import models as M           # the checker (bench.py --check only)
    rank, world = ctx.rank, ctx.world
    res = {"world": world}
    ok = True
    if rank == 0:
        B = ctx.args.batch or 65536
        ref = M.MLPRef(units=UNITS, seed=0)
        ref.randomize_biases(1)
        eng = MLPEngine.mlp_v1(units=UNITS, dtype="bf16", max_batch=B)
        eng.set_params_flat(MLPEngine.keras_to_flat([p.detach().numpy() for p in ref.params]))
        x, y = synthetic_batch(B, 0)
        t0 = time.perf_counter()
        want_loss, want_grads = ref.manual_train_step(x, y, emulate_bf16=True)
        t_cpu = time.perf_counter() - t0
        got_loss = float(eng.train_step(x.cuda(), y.cuda()).item())
        got = eng.split_flat(eng.get_grads_flat())
        want = eng.split_flat(MLPEngine.keras_to_flat([g.numpy() for g in want_grads]))
        errs = [float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)) for a, b in zip(got, want)]
        gn = float(np.sqrt(sum(float(np.sum(np.square(a, dtype=np.float64))) for a in got)))
        gn_want = float(np.sqrt(sum(float(np.sum(np.square(b, dtype=np.float64))) for b in want)))
        res["step0"] = {"batch": B, "loss": got_loss, "oracle_loss": float(want_loss), "loss_rel_err": abs(got_loss - float(want_loss)) / float(want_loss),
                        "grad_norm": gn, "oracle_grad_norm": gn_want, "worst_grad_rel_l2": max(errs), "oracle_cpu_s": t_cpu,
                        "tolerance": {"loss_rel": 1e-3, "grad_rel_l2": 1e-2}}
        ok &= res["step0"]["loss_rel_err"] <= 1e-3 and max(errs) <= 1e-2
        eng.close()
    if world > 1:
        Bl = 8192
        xs, ys = synthetic_batch(Bl * world, 7)
        shard = slice(rank * Bl, (rank + 1) * Bl)
        res["data_parallel"] = {}
        for dtype, tol, kind in (("fp32", 2e-4, "relmax"), ("bf16", 1e-2, "rel_l2")):
            init = glorot_uniform_flat(LAYER_DIMS, seed=3)
            eng = MLPEngine.mlp_v1(units=UNITS, dtype=dtype, max_batch=Bl)
            eng.set_params_flat(init)
            tr = Trainer(eng, rule="adam_keras", lr=1e-3)
            losses = [tr.step(xs[shard].cuda(), ys[shard].cuda()) for _ in range(2)]
            g_dp = eng.get_grads_flat()
            w_dp = eng.get_params_flat()
            eng.close()
            if rank == 0:
                one = MLPEngine.mlp_v1(units=UNITS, dtype=dtype, max_batch=Bl * world)
                one.set_params_flat(init)
                losses1 = []
                for _ in range(2):
                    losses1.append(float(one.train_step(xs.cuda(), ys.cuda()).item()))
                    g_one = one.get_grads_flat()
                    one.apply_opt("adam_keras", lr=1e-3)
                w_one = one.get_params_flat()
                one.close()
                if kind == "relmax":
                    g_err = float(np.abs(g_dp - g_one).max() / np.abs(g_one).max())
                else:
                    g_err = float(np.linalg.norm(g_dp - g_one) / np.linalg.norm(g_one))
                # Adam moves every weight by ~lr per step whatever the gradient's size: compare the weight CHANGE in relative L2
                dw_err = float(np.linalg.norm((w_dp - init) - (w_one - init)) / np.linalg.norm(w_one - init))
                l_err = max(abs(a - b) / abs(b) for a, b in zip(losses, losses1))
                res["data_parallel"][dtype] = {"ranks": world, "columns_per_rank": Bl, "steps": 2, "grad_err": g_err, "grad_err_kind": kind,
                                               "weight_change_rel_l2": dw_err, "loss_rel_err": l_err, "tolerance": tol}
                ok &= g_err <= tol and l_err <= 1e-4 and dw_err <= (2e-2 if dtype == "fp32" else 2e-1)
        ctx.barrier()
    res["ok"] = bool(ok)
    return res


# ------------------------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mlp_v1", choices=["mlp_v1", "cnn", "hsr", "ed"])
    ap.add_argument("--extras", default="cnn,hsr,ed", help="with --workload mlp_v1: the other BASELINE configurations measured in the same run "
                                                           "and reported under `workloads` (comma list or 'none')")
    ap.add_argument("--batch", type=int, default=0, help="columns per GPU per step (weak scaling); 0 = the BASELINE configuration's")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="columns per CPU-baseline step (0 = per-workload default)")
    ap.add_argument("--no-legs", dest="legs", action="store_false", help="reference arm: skip the BASELINE.md section-3 legs (B = 1024 / 3072)")
    ap.add_argument("--check", action="store_true", help="parity of the benchmarked step against the oracle (+ the data-parallel check under torchrun)")
    args = ap.parse_args()
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    args.extras = [] if (args.extras == "none" or args.workload != "mlp_v1") else [x for x in args.extras.split(",") if x]
    assert all(x in ("cnn", "hsr", "ed") for x in args.extras)
    # stdout must carry exactly ONE line, the JSON: libraries (NCCL prints its version banner with printf) get stderr as their
    # file descriptor 1 for the whole run, and emit() writes the line to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict) -> None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ctx = Ctx(args)
    if args.impl == "reference":
        if ctx.rank != 0:
            return
        emit(reference_arm(args, config_of))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl ours (there is no CPU fallback)"
    affinity = bind_to_gpu_cpus(ctx.local_rank)
    torch.cuda.set_device(ctx.local_rank)
    if ctx.world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", ctx.local_rank))
    if args.check:
        res = run_check(ctx)
        if ctx.rank == 0:
            emit({"check": res, "ok": res["ok"]})
        if ctx.world > 1:
            torch.distributed.destroy_process_group()
        sys.exit(0 if res["ok"] else 1)

    ctx.sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        ctx.sampler.start()
    sub_warm = max(3, min(args.warmup, 5))
    caps = {"cnn": 10, "hsr": 30, "ed": 50}
    if args.workload == "mlp_v1":
        line = run_mlp_v1(ctx)
        subs = {}
        for x in args.extras:
            fn = {"cnn": run_cnn, "hsr": run_hsr, "ed": run_ed}[x]
            try:
                subs[x] = fn(ctx, max(3, min(args.steps, caps[x])), sub_warm)
            except Exception as e:               # a secondary workload must never take the headline down with it
                subs[x] = {"error": f"{type(e).__name__}: {e}"}
                if ctx.world > 1:
                    raise
        if ctx.rank == 0 and subs:
            line["workloads"] = subs
        if ctx.world == 1 and args.extras and args.dtype == "bf16":
            try:
                line["arithmetic_modes"] = run_modes(ctx)
            except Exception as e:
                line["arithmetic_modes"] = {"error": f"{type(e).__name__}: {e}"}
    else:
        fn = {"cnn": run_cnn, "hsr": run_hsr, "ed": run_ed}[args.workload]
        sub = fn(ctx, args.steps, args.warmup)
        line = None
        if ctx.rank == 0:
            line = {"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic"}
            line.update(sub)
    ctx.sampler.stop()
    if ctx.rank == 0:
        line["cpu_affinity"] = affinity
        emit(line)
    if ctx.world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
