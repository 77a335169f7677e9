"""Seeded synthetic ClimSim columns (there is no dataset offline): the shapes, column order and value ranges of the
low-res V1 arrays.  Used by the tests and by bench.py for BOTH the CUDA path and the CPU baseline."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

IN_DIM, OUT_DIM = 124, 128


def synthetic_norm(seed: int = 1234) -> Dict[str, np.ndarray]:
    """inp_sub / inp_div / out_scale vectors with plausible magnitudes (T 180-320 K, q 0-0.03, ps 5e4-1.05e5,
    SOLIN 0-1400, LHFLX -100..600, SHFLX -200..500); one level has max == min to exercise the nan/inf -> 0 rule."""
    rng = np.random.default_rng(seed)
    lev = np.linspace(0.0, 1.0, 60)
    t_mean = 200.0 + 90.0 * lev + rng.normal(0, 2, 60)
    q_mean = 1e-6 + 0.012 * lev ** 3
    sub = np.concatenate([t_mean, q_mean, [9.8e4, 350.0, 80.0, 20.0]])
    div = np.concatenate([40.0 + 30.0 * rng.random(60), 1e-6 + 0.03 * lev ** 2 + 1e-5 * rng.random(60),
                          [5.5e4, 1400.0, 700.0, 700.0]])
    div[60] = 0.0                                            # max == min at the model top for q
    out_scale = 10.0 ** rng.uniform(0.0, 7.0, size=OUT_DIM)
    return {"inp_sub": sub, "inp_div": div, "out_scale": out_scale}


def synthetic_batch(B: int, seed: int = 0, device: str = "cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Normalised inputs x ~ N(0, 0.2^2) (B,124) and scaled targets y ~ N(0, 0.1^2) (B,128) with the eight scalar
    targets made non-negative (fluxes / precipitation)."""
    g = torch.Generator().manual_seed(1234 + seed)
    x = 0.2 * torch.randn(B, IN_DIM, generator=g, dtype=torch.float32)
    y = 0.1 * torch.randn(B, OUT_DIM, generator=g, dtype=torch.float32)
    y[:, 120:] = y[:, 120:].abs()
    return x.to(device), y.to(device)


def synthetic_raw_batch(B: int, norm: Dict[str, np.ndarray], seed: int = 0) -> np.ndarray:
    """Raw-space (un-normalised) inputs, fp64, consistent with ``norm``."""
    rng = np.random.default_rng(4321 + seed)
    z = 0.2 * rng.standard_normal((B, IN_DIM))
    div = np.where(norm["inp_div"] == 0, 1.0, norm["inp_div"])
    return norm["inp_sub"] + z * div
