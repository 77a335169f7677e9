"""CPU oracle: PyTorch restatements of the reference's column-emulator models, losses and optimizer steps.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every class cites the reference file:line it follows
(paths relative to the reference checkout).  All weights use the *Keras* convention -- Dense kernel stored
(in, out), Conv1D kernel stored (k, C_in, C_out), channels-last activations -- except ``HSRRef`` which restates a
PyTorch model and therefore keeps ``torch.nn.Linear``'s (out, in).

PARITY STATUS: MLPRef / EDRef / CNNRef / keras_adam_step / cyclical_lr are UNPINNED (tensorflow is not installable
here; the reference ships no golden outputs) -- only parameter/FLOP counts are pinned.  HSRRef is PINNED against the
reference's own hsr.py (tests/golden/hsr_small.npz; tests/test_oracle_pinning.py).
"""
from __future__ import annotations

import math

import numpy as np
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------------------
# activations (Keras layer semantics)
# --------------------------------------------------------------------------------------------------------------

def activation(name: str, x: torch.Tensor, alpha: float = 0.15) -> torch.Tensor:
    """keras.layers.ReLU / ELU(alpha=1) / LeakyReLU(alpha) -- baseline_v1/hpo_baseline_v1.py:82-87,91-96."""
    if name in ("none", "linear"):
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "elu":
        return F.elu(x, alpha=1.0)
    if name == "leakyrelu":
        return F.leaky_relu(x, negative_slope=alpha)
    raise ValueError(f"unknown activation {name!r}")


def glorot_uniform(fan_in: int, fan_out: int, shape: Sequence[int], gen: torch.Generator,
                   dtype=torch.float32) -> torch.Tensor:
    """Keras default kernel initializer: U(-l, l), l = sqrt(6 / (fan_in + fan_out))."""
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(*shape, generator=gen, dtype=torch.float64) * 2.0 - 1.0) * limit).to(dtype)


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundNode(torch.autograd.Function):
    """Marks a tensor the tensor-core engine keeps in bf16: ``fwd`` rounds the value on the way forward (a stored activation),
    ``bwd`` rounds the gradient flowing back through this point (a stored dZ / dA).  With these nodes at the engine's storage
    points, ordinary autograd over fp32 ops reproduces the arithmetic contract of the CSB_BF16 mode (bf16 operands, fp32
    accumulation) for graphs that are awkward to back-propagate by hand (the ResNet-1D, the encoder-decoder)."""

    @staticmethod
    def forward(ctx, t, fwd, bwd):
        ctx.bwd = bwd
        return _bf16_round(t) if fwd else t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return (_bf16_round(g) if ctx.bwd else g), None, None


def _tf32_round(t: torch.Tensor) -> torch.Tensor:
    """Round to nearest onto the TF32 grid (sign, 8 exponent bits, 10 mantissa bits): the storage rounding of the CSB_TF32 engine
    (climsim_b200/csrc/tc_gemm.cuh::round_tf32: add half an ulp of the 13 dropped bits, then clear them)."""
    t = t.to(torch.float32).contiguous()
    return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _mark(t: torch.Tensor, fwd: bool, bwd: bool, on: bool) -> torch.Tensor:
    return _RoundNode.apply(t, fwd, bwd) if on else t


def _emulated_leaves(params: Sequence[torch.Tensor], on: bool) -> List[torch.Tensor]:
    """Leaves to differentiate with respect to: the parameters themselves, or -- emulating the engine -- bf16-rounded copies of
    the kernels (the engine's bf16 weight copies; dW does not depend on W's own rounding) next to the fp32 biases."""
    if not on:
        return list(params)
    return [(_bf16_round(p.detach()) if p.dim() > 1 else p.detach().clone()).requires_grad_(True) for p in params]


# --------------------------------------------------------------------------------------------------------------
# MLP_v1  (Keras functional model)
# --------------------------------------------------------------------------------------------------------------

class MLPRef:
    """Restatement of ``MyHyperModel.build`` -- baseline_models/MLP/training/HPO/baseline_v1/
    hpo_baseline_v1.py:75-103 (identical graph: step2_retrain/step2_retrain.py:95-126).

        x -> [Dense(units_k) -> act] * n_layers -> Dense(128) -> act -> concat(Dense(120, linear), Dense(8, relu))

    Defaults are the shipped best trial (step1_results.csv lot-147/trial_0027): units [768,640,512,640,640],
    LeakyReLU(alpha=.15).  ``params`` is the Keras ``model.get_weights()`` list:
    [W0 (in,out), b0, ..., W_u (h,128), b_u, W_lin (128,120), b_lin, W_relu (128,8), b_relu].
    """

    def __init__(self, units: Sequence[int] = (768, 640, 512, 640, 640), act: str = "leakyrelu",
                 alpha: float = 0.15, in_dim: int = 124, out_lin: int = 120, out_relu: int = 8,
                 seed: int = 0, dtype: torch.dtype = torch.float32):
        self.units, self.act, self.alpha = list(units), act, alpha
        self.in_dim, self.out_lin, self.out_relu = in_dim, out_lin, out_relu
        self.dtype = dtype
        gen = torch.Generator().manual_seed(seed)
        dims = [in_dim] + self.units + [out_lin + out_relu]
        self.params: List[torch.Tensor] = []
        for k, n in zip(dims[:-1], dims[1:]):
            self.params += [glorot_uniform(k, n, (k, n), gen, dtype), torch.zeros(n, dtype=dtype)]
        h = out_lin + out_relu
        self.params += [glorot_uniform(h, out_lin, (h, out_lin), gen, dtype), torch.zeros(out_lin, dtype=dtype)]
        self.params += [glorot_uniform(h, out_relu, (h, out_relu), gen, dtype), torch.zeros(out_relu, dtype=dtype)]
        for p in self.params:
            p.requires_grad_(True)

    # -- bookkeeping -------------------------------------------------------------------------------------------
    def num_parameters(self) -> int:
        return sum(p.numel() for p in self.params)

    def flops_per_sample(self) -> int:
        """2*MAC + bias adds, the count keras-flops reports (FLOP_calculation.ipynb cells 5-6: 3 503 488)."""
        return sum(2 * w.shape[0] * w.shape[1] + w.shape[1] for w in self.params[0::2])

    def randomize_biases(self, seed: int = 1, scale: float = 0.05) -> None:
        """Keras biases start at zero; tests perturb them so that a bias bug cannot hide."""
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for b in self.params[1::2]:
                b.copy_((torch.rand(b.shape, generator=gen, dtype=torch.float64) * 2 - 1).to(self.dtype) * scale)

    # -- graph -------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, emulate_bf16: bool = False, return_hidden: bool = False, masks=None):
        """``emulate_bf16`` rounds the GEMM operands (weights, layer inputs) to bf16 but accumulates in fp32 --
        the numerics of the tensor-core path -- so that the bf16 CUDA mode can be checked tightly.
        ``masks``: optional list (one (B, units) multiplier per hidden layer) standing for a Dropout layer behind every activation
        (0 or 1/(1-p): the engine's own keep decisions replayed, csb_mlp_debug_dropout_mask)."""
        rnd = _bf16_round if emulate_bf16 else (lambda t: t)
        p = self.params
        n_hidden = len(self.units) + 1                       # hidden layers + the Dense(128) "upper output" layer
        h = rnd(x.to(self.dtype))
        hidden = []
        for i in range(n_hidden):
            h = rnd(activation(self.act, h @ rnd(p[2 * i]) + p[2 * i + 1], self.alpha))
            if masks is not None:
                h = rnd(h * masks[i].to(self.dtype))
            hidden.append(h)
        w_lin, b_lin, w_relu, b_relu = p[2 * n_hidden: 2 * n_hidden + 4]
        out = torch.cat([h @ rnd(w_lin) + b_lin, torch.relu(h @ rnd(w_relu) + b_relu)], dim=1)
        return (out, hidden) if return_hidden else out

    __call__ = forward

    def manual_train_step(self, x: torch.Tensor, y: torch.Tensor, w: Optional[torch.Tensor] = None,
                          grad_scale: Optional[float] = None, emulate_bf16: bool = False):
        """Explicit forward + weighted-MSE + backward (no autograd): returns (loss, grads in ``params`` order).

        With ``emulate_bf16`` every tensor the tensor-core path stores in bf16 is rounded at the same point
        (normalised input, weights, post-activation outputs, every dZ) while all products accumulate in fp32 --
        the arithmetic contract of the CSB_BF16 mode (climsim_b200/csrc/mlp_engine.cu), so that its gradients can
        be checked tightly instead of against a loose fp32 tolerance.  Without it this is plain fp32 backprop and
        must agree with autograd (tests/test_oracle_pinning.py)."""
        # ``emulate_bf16="tf32"``: the same storage points rounded onto the TF32 grid instead (the CSB_TF32 engine: fp32 storage whose
        # tensor-core operands carry 10 mantissa bits, products exact, fp32 accumulation)
        rnd = _tf32_round if emulate_bf16 == "tf32" else (_bf16_round if emulate_bf16 else (lambda t: t))
        with torch.no_grad():
            p = [t.detach() for t in self.params]
            n_hidden = len(self.units) + 1
            B = x.shape[0]
            out_dim = self.out_lin + self.out_relu
            scale = grad_scale if grad_scale is not None else 1.0 / (B * out_dim)
            w = torch.ones(out_dim, dtype=self.dtype) if w is None else w.to(self.dtype)
            acts = [rnd(x.to(self.dtype))]
            for i in range(n_hidden):
                acts.append(rnd(activation(self.act, acts[-1] @ rnd(p[2 * i]) + p[2 * i + 1], self.alpha)))
            w_head = torch.cat([p[2 * n_hidden], p[2 * n_hidden + 2]], dim=1)
            b_head = torch.cat([p[2 * n_hidden + 1], p[2 * n_hidden + 3]])
            z = acts[-1] @ rnd(w_head) + b_head
            pred = torch.cat([z[:, :self.out_lin], torch.relu(z[:, self.out_lin:])], dim=1)
            d = pred - y.to(self.dtype)
            loss = (w * d * d).sum() * scale
            dact_head = torch.ones_like(pred)
            dact_head[:, self.out_lin:] = (pred[:, self.out_lin:] > 0).to(self.dtype)
            dz = rnd(2.0 * w * d * scale * dact_head)
            grads: List[Optional[torch.Tensor]] = [None] * len(p)
            g_w_head, g_b_head = acts[-1].t() @ dz, dz.sum(dim=0)
            grads[2 * n_hidden], grads[2 * n_hidden + 2] = g_w_head[:, :self.out_lin], g_w_head[:, self.out_lin:]
            grads[2 * n_hidden + 1], grads[2 * n_hidden + 3] = g_b_head[:self.out_lin], g_b_head[self.out_lin:]
            w_next = rnd(w_head)
            for i in range(n_hidden - 1, -1, -1):
                a = acts[i + 1]
                if self.act == "relu":
                    da = (a > 0).to(self.dtype)
                elif self.act == "leakyrelu":
                    da = torch.where(a > 0, torch.ones_like(a), torch.full_like(a, self.alpha))
                elif self.act == "elu":
                    da = torch.where(a > 0, torch.ones_like(a), a + 1.0)
                else:
                    da = torch.ones_like(a)
                dz = rnd((dz @ w_next.t()) * da)
                grads[2 * i], grads[2 * i + 1] = acts[i].t() @ dz, dz.sum(dim=0)
                w_next = rnd(p[2 * i])
        return loss, grads


# --------------------------------------------------------------------------------------------------------------
# ED  (Keras encoder-decoder MLP)
# --------------------------------------------------------------------------------------------------------------

def ed_widths(intermediate_dim: int = 463, latent_dim: int = 5, out_dim: int = 128) -> List[int]:
    """Layer widths of baseline_models/ED/training/ClimSIM_ED_1_3_train.py:56-78.  The script passes floats such
    as ``intermediate_dim/2`` to ``Dense``; Keras applies ``int()`` (truncation): 231, 115, 57, 28."""
    d = intermediate_dim
    enc = [d, d, int(d / 2), int(d / 4), int(d / 8), int(d / 16), latent_dim]
    dec = [int(d / 16), int(d / 8), int(d / 4), int(d / 2), d, d, out_dim]
    return enc + dec


class EDRef:
    """ReLU MLP 124->463->463->231->115->57->28->5->28->57->115->231->463->463->128(ELU);
    ClimSIM_ED_1_3_train.py:56-92.  ``params`` = [W0, b0, W1, b1, ...] Keras order/layout."""

    def __init__(self, in_dim: int = 124, seed: int = 0, dtype: torch.dtype = torch.float32):
        gen = torch.Generator().manual_seed(seed)
        dims = [in_dim] + ed_widths()
        self.dims, self.dtype = dims, dtype
        self.params: List[torch.Tensor] = []
        for k, n in zip(dims[:-1], dims[1:]):
            self.params += [glorot_uniform(k, n, (k, n), gen, dtype), torch.zeros(n, dtype=dtype)]
        for p in self.params:
            p.requires_grad_(True)

    def num_parameters(self) -> int:
        return sum(p.numel() for p in self.params)

    def forward(self, x: torch.Tensor, emulate_bf16: bool = False, params: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """``emulate_bf16``: the rounding points of the CSB_BF16 engine (input, weights, every stored activation; on the way back
        every stored dZ) -- see ``_RoundNode``."""
        p = self.params if params is None else params
        h = _bf16_round(x.to(self.dtype)) if emulate_bf16 else x.to(self.dtype)
        n = len(p) // 2
        for i in range(n):
            z = h @ p[2 * i] + p[2 * i + 1]
            if i == n - 1:
                h = F.elu(_mark(z, False, True, emulate_bf16), alpha=1.0)          # dZ of the output layer is stored in bf16
            else:
                h = _mark(torch.relu(z), True, True, emulate_bf16)                # activation stored in bf16; dZ = bf16(dA * relu')
        return h

    __call__ = forward

    def emulated_train_step(self, x: torch.Tensor, y: torch.Tensor, emulate_bf16: bool = True):
        """(loss, grads in ``params`` order) of Keras ``loss='mse'`` with the engine's bf16 rounding points (autograd over
        ``_RoundNode``-marked fp32 ops); ``emulate_bf16=False`` is plain fp32 autograd."""
        leaves = _emulated_leaves(self.params, emulate_bf16)
        loss = mse(y.to(self.dtype), self.forward(x, emulate_bf16, leaves))
        return loss.detach(), list(torch.autograd.grad(loss, leaves))


# --------------------------------------------------------------------------------------------------------------
# CNN  (Keras ResNet-1D)
# --------------------------------------------------------------------------------------------------------------

def conv1d_same_cl(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Keras ``Conv1D(padding='same')`` on channels-last input.  x (B, L, Cin); w (k, Cin, Cout); b (Cout).
    out[b, l, co] = sum_{t, ci} x[b, l + t - (k-1)//2, ci] * w[t, ci, co] + b[co], zero outside [0, L)."""
    k = w.shape[0]
    y = F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), b, padding=(k - 1) // 2)
    return y.transpose(1, 2)


class CNNRef:
    """Restatement of ``CNNHyperModel.build`` -- baseline_models/CNN/training/hpo_train.py:131-200.

    12 x { Conv1D(406,k=3,same) -> ReLU -> Dropout -> Conv1D(406,k=3,same) -> ReLU -> Dropout -> + Conv1D(406,k=1)(block input) }
    -> Conv1D(10, k=1, ELU) -> per-level Dense(2, linear) || Dense(8, relu) -> (B, 60, 10).
    Dropout: inference mode by default; ``forward(x, masks=...)`` applies given multiplicative masks (0 or 1/(1-p)) behind the two
    ReLUs of every block, which is what keras.layers.Dropout does in training mode with whatever its RNG drew (TF's stream itself
    cannot be reproduced) -- the GPU test feeds the masks the engine used.
    ``params`` order = Keras ``get_weights()``: per block [Wc1, bc1, Wc2, bc2, Wres, bres], then [Wout, bout,
    Wlin, blin, Wrelu, brelu].
    """

    def __init__(self, depth: int = 12, width: int = 406, kernel: int = 3, in_ch: int = 6, out_ch: int = 10,
                 out_lin: int = 2, seed: int = 0, dtype: torch.dtype = torch.float32):
        self.depth, self.width, self.kernel = depth, width, kernel
        self.in_ch, self.out_ch, self.out_lin, self.dtype = in_ch, out_ch, out_lin, dtype
        gen = torch.Generator().manual_seed(seed)

        def conv(k, ci, co):
            return [glorot_uniform(k * ci, k * co, (k, ci, co), gen, dtype), torch.zeros(co, dtype=dtype)]

        def dense(ci, co):
            return [glorot_uniform(ci, co, (ci, co), gen, dtype), torch.zeros(co, dtype=dtype)]

        self.params: List[torch.Tensor] = []
        c = in_ch
        for _ in range(depth):
            self.params += conv(kernel, c, width) + conv(kernel, width, width) + conv(1, c, width)
            c = width
        self.params += conv(1, c, out_ch) + dense(out_ch, out_lin) + dense(out_ch, out_ch - out_lin)
        for p in self.params:
            p.requires_grad_(True)

    def num_parameters(self) -> int:
        return sum(p.numel() for p in self.params)

    def randomize_biases(self, seed: int = 1, scale: float = 0.05) -> None:
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for b in self.params[1::2]:
                b.copy_((torch.rand(b.shape, generator=gen, dtype=torch.float64) * 2 - 1).to(self.dtype) * scale)

    def forward(self, x: torch.Tensor, masks=None, emulate_bf16: bool = False, params: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """``masks``: optional list (one entry per block) of pairs of (B, 60, width) multipliers for the two Dropout layers.
        ``emulate_bf16``: the storage points of the CSB_BF16 engine (climsim_b200/csrc/cnn_engine.cuh) are marked with
        ``_RoundNode``: forward -- the packed input, relu(conv1), relu(conv2) (again after the dropout scaling), the block output
        ``conv1x1(x_in) + relu(conv2)`` and elu(conv_out) are stored in bf16; backward -- dL/dz of the heads, dze, the block-output
        gradient G, dz2 and dz1 are stored in bf16; d(block input) = ``conv1^T(dz1) + conv1x1^T(G)`` is ONE fp32 accumulation (the
        engine's merged GEMM over [dz1 taps | G]) with one rounding, and dz2 derives from the ROUNDED G.  (With CSB_CNN_NO_MERGE=1
        the engine stores the conv1 branch T in bf16 first: one more rounding, inside the test tolerances.)  Weights enter as bf16
        copies (``params`` = ``_emulated_leaves``)."""
        p = self.params if params is None else params
        e = emulate_bf16
        h = _bf16_round(x.to(self.dtype)) if e else x.to(self.dtype)
        prev = h
        for i in range(self.depth):
            wc1, bc1, wc2, bc2, wr, br = p[6 * i: 6 * i + 6]
            h = prev                                                                  # conv1 branch: its gradient joins G's sum unrounded
            h = _mark(torch.relu(conv1d_same_cl(h, wc1, bc1)), True, True, e)         # h1 bf16; dz1 = bf16(conv2^T(dz2) * relu'(h1))
            if masks is not None:
                h = _mark(h * masks[i][0], True, True, e)
            h = _mark(torch.relu(conv1d_same_cl(h, wc2, bc2)), True, True, e)         # h2 bf16; dz2 = bf16(G * relu'(h2))
            if masks is not None:
                h = _mark(h * masks[i][1], True, True, e)
            h = _mark(h + conv1d_same_cl(prev, wr, br), True, True, e)                # block output bf16; G = bf16(conv1x1^T(G') + T')
            prev = h
        wo, bo, wl, bl, wrl, brl = p[6 * self.depth: 6 * self.depth + 6]
        u = _mark(conv1d_same_cl(h, wo, bo), False, True, e)                          # dze = bf16((dzh . Wd^T) * elu'(e))
        h = _mark(F.elu(u, alpha=1.0), True, False, e)                                # e stored in bf16
        zl = _mark(h @ wl + bl, False, True, e)                                       # dzh (both heads) stored in bf16
        zr = _mark(h @ wrl + brl, False, True, e)
        return torch.cat([zl, torch.relu(zr)], dim=-1)

    __call__ = forward

    def emulated_train_step(self, x: torch.Tensor, y: torch.Tensor, loss: str = "mae", masks=None, emulate_bf16: bool = True):
        """(loss, grads in ``params`` order) of ``mae_adjusted`` / ``mse_adjusted`` with the engine's bf16 rounding points;
        ``emulate_bf16=False`` is plain fp32 autograd."""
        leaves = _emulated_leaves(self.params, emulate_bf16)
        fn = mae_adjusted if loss == "mae" else mse_adjusted
        val = fn(y.to(self.dtype), self.forward(x, masks, emulate_bf16, leaves))
        return val.detach(), list(torch.autograd.grad(val, leaves))


# --------------------------------------------------------------------------------------------------------------
# HSR  (PyTorch model in the reference)
# --------------------------------------------------------------------------------------------------------------

class HSRMLPRef(torch.nn.Module):
    """Restatement of ``MLP`` -- baseline_models/HSR/training/hsr.py:14-35:
    layers x [Linear -> LayerNorm(eps 1e-5) -> Dropout(p) -> ReLU] -> Linear.  Same ``state_dict`` keys as the
    reference (``linear{i}.0.weight`` ... ``final_linear.weight``) so the shipped ``final_hsr.cp`` loads."""

    def __init__(self, in_dims: int, out_dims: int, hidden_dims: int = 512, layers: int = 1, dropout: float = 0.0):
        super().__init__()
        self.n_layers = layers
        for i in range(layers):
            self.add_module("linear%d" % i, torch.nn.Sequential(
                torch.nn.Linear(in_dims if i == 0 else hidden_dims, hidden_dims),
                torch.nn.LayerNorm(hidden_dims),
                torch.nn.Dropout(p=dropout)))
        self.final_linear = torch.nn.Linear(hidden_dims, out_dims)

    def forward(self, x: torch.Tensor, masks=None) -> torch.Tensor:
        """``masks``: optional explicit Dropout multipliers (one (B, hidden) tensor per block, 0 or 1/(1-p)) in place of the
        module's own random ``Dropout`` -- the engine's keep decisions replayed (hsr.py:20-25: Linear -> LayerNorm -> Dropout -> ReLU)."""
        for i in range(self.n_layers):
            seq = getattr(self, "linear%d" % i)
            x = torch.relu(seq(x) if masks is None else seq[1](seq[0](x)) * masks[i])
        return self.final_linear(x)


class HSRRef(torch.nn.Module):
    """Restatement of ``HeteroskedasticRegression`` -- hsr.py:38-67 (forward) and :126-138 (loss)."""

    def __init__(self, in_dims: int = 124, out_dims: int = 128, hidden_dims: int = 512, layers: int = 1,
                 dropout: float = 0.0):
        super().__init__()
        self.mean = HSRMLPRef(in_dims, out_dims, hidden_dims, layers, dropout)
        self.logprec = HSRMLPRef(in_dims, out_dims, hidden_dims, layers, dropout)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.mean(x), self.logprec(x)

    @staticmethod
    def loss(mu: torch.Tensor, logprec: torch.Tensor, y: torch.Tensor, mle: bool) -> torch.Tensor:
        """hsr.py:128-138: MSE for the first third of the epochs, then Gaussian NLL; clipped to +-1e5."""
        if mle:
            loss = (torch.exp(logprec) * (y - mu) ** 2 - logprec).mean()
        else:
            loss = ((y - mu) ** 2).mean()
        return torch.clip(loss, min=-1e5, max=1e5)

    @staticmethod
    def weight_decays(gamma: float = 0.022, rho: Optional[float] = None) -> Tuple[float, float]:
        """hsr.py:100-107: per-group L2 weight decay (alpha for the mean net, beta for the log-precision net)."""
        rho = rho if rho is not None else 1 - gamma
        return (1 - rho) / rho * gamma, (1 - rho) / rho * (1 - gamma)


# --------------------------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------------------------

def mse(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """Keras ``loss='mse'`` (hpo_baseline_v1.py:127-129): mean over the feature axis, then over the batch
    == global mean of the squared error."""
    return ((y_pred - y_true) ** 2).mean(dim=-1).mean()


def weighted_mse(y_true: torch.Tensor, y_pred: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """mean_ij( w_j * (yhat_ij - y_ij)^2 ); w == 1 reproduces ``mse`` (SURVEY.md section 0 resolution)."""
    return (w * (y_pred - y_true) ** 2).mean()


def mse_adjusted(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """baseline_models/CNN/training/hpo_train.py:114-116."""
    se = (y_pred - y_true) ** 2
    return se[:, :, 0:2].mean() * (120 / 128) + se[:, :, 2:10].mean() * (8 / 128)


def mae_adjusted(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """baseline_models/CNN/training/hpo_train.py:119-121."""
    ae = (y_pred - y_true).abs()
    return ae[:, :, 0:2].mean() * (120 / 128) + ae[:, :, 2:10].mean() * (8 / 128)


def cnn_loss_weights(levels: int = 60, dtype=torch.float32) -> torch.Tensor:
    """(60,10) weight w such that sum_{l,c} w[l,c]*e[b,l,c] averaged over b equals ``*_adjusted`` of e:
    1/128 per profile entry, 1/(128*60) per level-replicated scalar entry."""
    w = torch.empty(levels, 10, dtype=dtype)
    w[:, 0:2] = (120 / 128) / (levels * 2)
    w[:, 2:10] = (8 / 128) / (levels * 8)
    return w


# --------------------------------------------------------------------------------------------------------------
# optimizers / schedules (Keras + tfa semantics)
# --------------------------------------------------------------------------------------------------------------

def cyclical_lr(step: int, initial_lr: float = 2.5e-4, max_lr: float = 2.5e-3, step_size: float = 2.0,
                scale_mode: str = "cycle") -> float:
    """tfa.optimizers.CyclicalLearningRate as configured at hpo_baseline_v1.py:105-114
    (scale_fn = 1/2**(x-1), scale_mode='cycle').  Closed form of tensorflow-addons 0.19
    ``CyclicalLearningRate.__call__``."""
    cycle = math.floor(1 + step / (2 * step_size))
    x = abs(step / step_size - 2 * cycle + 1)
    mode_step = cycle if scale_mode == "cycle" else step
    return initial_lr + (max_lr - initial_lr) * max(0.0, 1 - x) * (1.0 / (2.0 ** (mode_step - 1)))


def keras_adam_step(params: List[torch.Tensor], grads: List[torch.Tensor], m: List[torch.Tensor],
                    v: List[torch.Tensor], t: int, lr: float, beta1: float = 0.9, beta2: float = 0.999,
                    eps: float = 1e-7) -> None:
    """keras.optimizers.Adam (2.11) ``update_step``; defaults as used at hpo_baseline_v1.py:116-117:
        alpha = lr * sqrt(1 - b2^t) / (1 - b1^t);  m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);
        w -= alpha * m / (sqrt(v) + eps)            (epsilon = 1e-7, *outside* the bias correction)."""
    alpha = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    with torch.no_grad():
        for p, g, mi, vi in zip(params, grads, m, v):
            mi.add_((g - mi) * (1 - beta1))
            vi.add_((g * g - vi) * (1 - beta2))
            p.sub_(alpha * mi / (vi.sqrt() + eps))


def torch_adam_step(params: List[torch.Tensor], grads: List[torch.Tensor], m: List[torch.Tensor],
                    v: List[torch.Tensor], t: int, lr: float, beta1: float = 0.9, beta2: float = 0.999,
                    eps: float = 1e-8, weight_decay: float = 0.0) -> None:
    """torch.optim.Adam (the optimizer the reference's HSR trainer builds, hsr.py:109-112) with L2
    ``weight_decay`` folded into the gradient; restated so the per-group decay can be checked elementwise."""
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    with torch.no_grad():
        for p, g, mi, vi in zip(params, grads, m, v):
            g = g + weight_decay * p if weight_decay != 0.0 else g
            mi.mul_(beta1).add_(g, alpha=1 - beta1)
            vi.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            p.sub_((lr / bc1) * mi / (vi.sqrt() / math.sqrt(bc2) + eps))


def radam_step(params: List[torch.Tensor], grads: List[torch.Tensor], m: List[torch.Tensor], v: List[torch.Tensor], t: int,
               lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7, sma_threshold: float = 5.0) -> None:
    """tensorflow_addons.optimizers.RectifiedAdam (0.19) ``_resource_apply_dense`` with its defaults (no warm-up, no
    weight decay, no amsgrad) -- the optimizer of the shipped best MLP_v1 trial (hpo_baseline_v1.py:118-119).  UNPINNED."""
    sma_inf = 2.0 / (1.0 - beta2) - 1.0
    b2t = beta2 ** t
    sma_t = sma_inf - 2.0 * t * b2t / (1.0 - b2t)
    with torch.no_grad():
        for p, g, mi, vi in zip(params, grads, m, v):
            mi.mul_(beta1).add_(g, alpha=1 - beta1)
            vi.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            mhat = mi / (1 - beta1 ** t)
            if sma_t >= sma_threshold:
                r = math.sqrt((sma_t - 4) / (sma_inf - 4) * (sma_t - 2) / (sma_inf - 2) * sma_inf / sma_t)
                p.sub_(lr * r * mhat / ((vi / (1 - b2t)).sqrt() + eps))
            else:
                p.sub_(lr * mhat)


def keras_rmsprop_step(params: List[torch.Tensor], grads: List[torch.Tensor], v: List[torch.Tensor], lr: float,
                       rho: float = 0.9, eps: float = 1e-7) -> None:
    """keras.optimizers.RMSprop (2.11) ``update_step``, momentum 0, not centered (hpo_baseline_v1.py:120-121):
    v = rho v + (1-rho) g^2;  w -= lr * g * rsqrt(v + eps).  UNPINNED."""
    with torch.no_grad():
        for p, g, vi in zip(params, grads, v):
            vi.mul_(rho).addcmul_(g, g, value=1 - rho)
            p.sub_(lr * g * torch.rsqrt(vi + eps))


class OnlineMLPRef(torch.nn.Module):
    """Restatement of the online MLP -- online_testing/baseline_models/MLP_v2rh/training/mlp.py:24-68 -- without the Modulus
    base class (not installable here): same ``state_dict`` keys (``linears.{i}.0.*``, ``final_linear.*``), same forward incl.
    ``output_prune`` and the ReLU on the last eight outputs.  PINNED: tests/golden/online_mlp.npz holds outputs of the reference
    class itself (imported with a stand-in for the modulus base classes, tests/golden/make_golden.py::make_online)."""

    def __init__(self, in_dims, out_dims, hidden_dims, layers, dropout=0.0, output_prune=False, strato_lev_out=15):
        super().__init__()
        hidden = list(hidden_dims) if isinstance(hidden_dims, (list, tuple)) else [hidden_dims] * layers
        assert len(hidden) == layers
        self.output_prune, self.strato_lev_out = output_prune, strato_lev_out
        self.linears = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(in_dims if i == 0 else hidden[i - 1], hidden[i]),
                                                                torch.nn.Dropout(p=dropout)) for i in range(layers)])
        self.final_linear = torch.nn.Linear(hidden[-1], out_dims)

    def forward(self, x):
        for linear in self.linears:
            x = torch.relu(linear(x))
        x = self.final_linear(x)
        if self.output_prune:
            mask = torch.ones(x.shape[1], dtype=x.dtype)
            for start in (60, 120, 180, 240):
                mask[start:start + self.strato_lev_out] = 0
            x = x * mask
        return torch.cat([x[:, :-8], torch.relu(x[:, -8:])], dim=1)


class OnlineWrapperRef(torch.nn.Module):
    """Restatement of ``NewModel`` -- online_testing/model_postprocessing/v2_nn_wrapper.ipynb, cell 5 -- the module exported for the
    E3SM coupling: raw (B, 557) inputs -> physical-unit (B, 368) tendencies.  Same constructor arguments; the pruning ranges the
    notebook hard-codes are the defaults here.  PINNED by tests/golden/online_mlp.npz (outputs of the notebook's own class).
    The same pre-processing is what the training pipeline applies per sample
    (online_testing/baseline_models/MLP_v2rh/training/climsim_datapip_h5.py:132-168)."""

    def __init__(self, original_model, input_sub, input_div, out_scale, lbd_qc, lbd_qi, prune_qn_levels=15, rh_clip=(0.0, 1.2),
                 out_prune=((60, 75), (120, 148), (180, 195), (240, 255), (300, 315))):
        super().__init__()
        self.original_model = original_model
        f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
        self.input_sub, self.input_div, self.out_scale = f32(input_sub), f32(input_div), f32(out_scale)
        self.lbd_qc, self.lbd_qi = f32(lbd_qc), f32(lbd_qi)
        self.prune_qn_levels, self.rh_clip, self.out_prune = prune_qn_levels, rh_clip, out_prune

    def preprocessing(self, x):
        x = x.clone()
        x[:, 120:180] = 1 - torch.exp(-x[:, 120:180] * self.lbd_qc)
        x[:, 180:240] = 1 - torch.exp(-x[:, 180:240] * self.lbd_qi)
        x = (x - self.input_sub) / self.input_div
        x = torch.where(torch.isnan(x) | torch.isinf(x), torch.zeros((), dtype=x.dtype), x)
        x[:, 120:120 + self.prune_qn_levels] = 0
        x[:, 180:180 + self.prune_qn_levels] = 0
        x[:, 60:120] = torch.clamp(x[:, 60:120], self.rh_clip[0], self.rh_clip[1])
        return x

    def postprocessing(self, y):
        y = y.clone()
        for a, b in self.out_prune:
            y[:, a:b] = 0
        return y / self.out_scale

    def forward(self, x):
        return self.postprocessing(self.original_model(self.preprocessing(x)))
