#!/usr/bin/env python
"""bench.py -- columns/sec of the MLP_v1 training step (BASELINE.json configs[1]: MLP_v1 bf16, batch 65536 per GPU,
forward + weighted-MSE + backward + Adam) on N B200s of one node, plus the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` = whole-job columns/s with the inputs resident in HBM (device-timed, CUDA
events, max over ranks); `e2e` = the same step driven through the public Trainer.step() from pinned HOST buffers
(H2D of x,y and D2H of the loss inside the timed region); `roofline` = the dominant kernel kind, timed live with CUDA
events on the launching stream during the timed steps; `cpu_baseline` = the CPU oracle (PyTorch fp32 restatement of the
reference's Keras model + Keras Adam) on a bounded sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tf_sustained": d.get("bf16_tflops_sustained"), "tf_burst": d.get("bf16_tflops"), "hbm": d.get("hbm_gbs"), "src": "measured"}
    return {"tf_sustained": 1400.0, "tf_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


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

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for ts, ln in self.lines if t_begin <= ts <= t_end + 0.25]
        window = "timed region"
        if len(inside) < 2:                       # timed region shorter than two sampling periods: use warm-up + timed
            inside, window = [ln for _, ln in self.lines], "warm-up + timed region (timed region < 0.4 s)"
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


def best_cpu_threads(batch: int, max_threads: int) -> int:
    """PyTorch-CPU GEMMs of this size do not scale to every core of a big host (128 threads are ~10x slower than 16 on
    the GPU boxes): time one step at a few thread counts and keep the fastest, so the baseline is not sandbagged."""
    cands = sorted({c for c in (8, 16, 32, 64, max_threads) if c <= max_threads})
    best, best_cps = cands[0], 0.0
    for c in cands:
        cps, _ = cpu_reference_steps(batch, 2, 1, c)
        if cps > best_cps:
            best, best_cps = c, cps
    return best


def cpu_reference_steps(batch: int, steps: int, warmup: int, threads: int):
    """The reference arm: the CPU oracle (PyTorch fp32 restatement of the Keras MLP_v1 graph, Keras 'mse', Keras Adam)
    on `batch` synthetic columns per step.  Returns (columns/s, ms/step)."""
    from oracle import models as M          # the one place bench.py may execute oracle/: the CPU baseline
    from climsim_b200.synthetic import synthetic_batch
    torch.set_num_threads(threads)
    ref = M.MLPRef(units=UNITS, seed=0)
    x, y = synthetic_batch(batch, 0)
    m = [torch.zeros_like(p) for p in ref.params]
    v = [torch.zeros_like(p) for p in ref.params]
    t0 = None
    for it in range(warmup + steps):
        if it == warmup:
            t0 = time.perf_counter()
        for p in ref.params:
            p.grad = None
        M.mse(y, ref(x)).backward()
        M.keras_adam_step(ref.params, [p.grad for p in ref.params], m, v, it + 1, M.cyclical_lr(it, step_size=2000))
    dt = time.perf_counter() - t0
    return batch * steps / dt, 1e3 * dt / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="columns per GPU per step (weak scaling)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=4096, help="columns per CPU-baseline step")
    args = ap.parse_args()
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    # stdout must carry exactly ONE line, the JSON: libraries (NCCL prints its version banner with printf) get stderr as their
    # file descriptor 1 for the whole run, and emit() writes the line to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict) -> None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    config = {"workload": "MLP_v1 (124->768->640->512->640->640->128->[120|8], LeakyReLU .15) train step: normalised x, "
                          "fwd + MSE + bwd + Keras-Adam, cyclical LR", "columns_per_gpu_per_step": args.batch,
              "global_batch": args.batch * max(world, args.gpus if world == 1 else world), "parallelism": f"dp{max(world, 1)}",
              "l2": "inputs rotate over 4 distinct device-resident batches (264 MB > 126 MB L2); ~1.3 GB of activations "
                    "written/read per step"}

    if args.impl == "reference":
        if rank != 0:
            return
        threads = best_cpu_threads(args.cpu_sample, threads)
        cps, ms = cpu_reference_steps(args.cpu_sample, args.steps, args.warmup, threads)
        line = {"impl": "reference", "metric": "columns/sec", "value": cps, "unit": "columns/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": cps, "unit": "columns/s", "cores": threads, "kind": "port",
                                 "sample": f"{args.steps} steps of {args.cpu_sample} columns (PyTorch-CPU fp32 oracle of the Keras MLP_v1 train step)"},
                "e2e": {"value": cps, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl ours (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from climsim_b200 import MLPEngine
    from climsim_b200.synthetic import synthetic_batch
    from climsim_b200.trainer import Trainer, cyclical_lr, glorot_uniform_flat

    B = args.batch
    eng = MLPEngine.mlp_v1(units=UNITS, dtype=args.dtype, max_batch=B)
    eng.set_params_flat(glorot_uniform_flat(eng.layer_dims, seed=0))       # identical replicas on every rank
    trainer = Trainer(eng, rule="adam_keras", lr=lambda it: cyclical_lr(it, step_size=2000))
    batches = [synthetic_batch(B, seed=rank * 16 + i, device="cuda") for i in range(4)]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident timing (the `value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for it in range(max(args.warmup, 8)):          # >= 8 so that each of the 4 rotating batches has its CUDA graph captured
        x, y = batches[it % 4]
        trainer.step(x, y, return_loss=False)
    barrier()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    ev0.record()
    for it in range(args.steps):
        x, y = batches[it % 4]
        trainer.step(x, y, return_loss=False)
    ev1.record()
    barrier()
    t_end = time.time()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    t = torch.tensor([ms_total], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # ------------------------------------------------------------------ per-kernel-kind device time: a second pass over the same
    # steps with a CUDA event recorded on the launching stream after every launch (the step then runs as individual launches
    # instead of the cached CUDA graph, so this pass is a few per cent slower than the timed region above)
    prof_steps = min(args.steps, 100)
    eng.profile(True)
    barrier()
    ev0.record()
    for it in range(prof_steps):
        x, y = batches[it % 4]
        trainer.step(x, y, return_loss=False)
    ev1.record()
    barrier()
    prof_ms_total = ev0.elapsed_time(ev1)
    prof = eng.profile_read()
    eng.profile(False)

    # ------------------------------------------------------------------ end to end from pinned host buffers
    e2e_steps = max(3, min(args.steps, 50))
    # every step copies its own x,y from pinned host memory and reads its own loss back; the engine's two staging slots
    # let the copy of step i+1 overlap the compute of step i (sync=False), as an input pipeline with prefetch does
    host = [tuple(t_.cpu().pin_memory() for t_ in batches[i]) for i in range(4)]
    for it in range(4):
        trainer.step(*host[it % 4], sync=False)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for it in range(e2e_steps):
        loss_slot = trainer.step(*host[it % 4], sync=False)
    ev1.record()
    barrier()
    loss = float(loss_slot.item())
    e2e_ms = max(ev0.elapsed_time(ev1), 0.0)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([e2e_ms], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = measured_peaks()
        kinds = {}
        for k, (ms, n) in prof.items():
            d = {"ms_per_step": ms / prof_steps, "launches_per_step": n / prof_steps}
            if k in KIND_FLOPS:
                d["tflops"] = KIND_FLOPS[k] * B / (ms / prof_steps * 1e-3) / 1e12
            kinds[k] = d
        dom = max((k for k in kinds if k in KIND_FLOPS), key=lambda k: kinds[k]["ms_per_step"])
        n_dom = kinds[dom]["launches_per_step"]
        achieved = kinds[dom]["tflops"]
        peak = peaks["tf_sustained"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and B == 65536:
            traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a long step)", "traffic": traffic,
                    "traffic_note": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum from profiles/r01_traffic.json (ncu --set full)",
                    "flops_per_launch": KIND_FLOPS[dom] * B / max(n_dom, 1), "avg_launch_ms": kinds[dom]["ms_per_step"] / max(n_dom, 1),
                    "step_tflops": FLOP_TRAIN * B / (ms_total / args.steps * 1e-3) / 1e12,
                    "step_frac_of_peak": FLOP_TRAIN * B / (ms_total / args.steps * 1e-3) / 1e12 / peak}
        cpu_baseline = None
        if world == 1:
            threads = best_cpu_threads(args.cpu_sample, threads)
            cps, cms = cpu_reference_steps(args.cpu_sample, 8, 2, threads)
            cpu_baseline = {"value": cps, "unit": "columns/s", "cores": threads, "kind": "port",
                            "sample": f"8 steps of {args.cpu_sample} columns of the same synthetic workload (PyTorch-CPU fp32 oracle, "
                                      f"fwd+bwd+Keras-Adam), {cms:.0f} ms/step"}
        line = {"metric": "columns/sec", "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "columns/s", "h2d_bytes_per_step": world * B * (IN_DIM + OUT_DIM) * 4,
                        "d2h_bytes_per_step": world * 4, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                        "wall_ms_per_step": wall_ms / e2e_steps, "last_loss": loss,
                        "api": "climsim_b200.Trainer.step(x_pinned, y_pinned, sync=False) -> csb_mlp_train_step_host_async (N=1) / "
                               "csb_mlp_stage_host_batch + train_step + all-reduce + apply_opt (N>1); per step: H2D of x,y into one of two "
                               "staging slots on the copy stream (overlaps the previous step's compute), D2H of the loss into a pinned slot"},
                "gpu_launches": launches, "roofline": roofline, "kernels": kinds,
                "kernels_note": f"per-kind times from a second pass of {prof_steps} steps with per-launch CUDA events (eager launches, "
                                f"{prof_ms_total / prof_steps:.4f} ms/step); the timed region replays the step as a CUDA graph",
                "cpu_baseline": cpu_baseline}
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
