"""Python face of the csb_mlp_* C ABI: ``MLPEngine`` owns one handle (one model replica on one GPU).

Everything here is plumbing -- pointer extraction from torch tensors, the current CUDA stream, error mapping.
No arithmetic happens in Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


class _DevPtr:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _f32_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (climsim_b200 has no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


class MLPEngine:
    """A dense-stack column emulator bound to libclimsim_b200.so.

    ``layers`` is a list of ``(units, activation, alpha)``; the last entry is the output layer.  ``head_relu_from``
    turns the output layer into the reference's two-head output (columns >= head_relu_from get ReLU).
    """

    def __init__(self, in_dim: int, layers: Sequence[Tuple[int, str, float]], head_relu_from: int = -1,
                 dtype: str = "bf16", loss: str = "mse", max_batch: int = 65536, layernorm: Optional[Sequence[bool]] = None):
        self.lib = _lib.load()
        cfg = _lib.MlpCfg()
        cfg.in_dim, cfg.n_layers = in_dim, len(layers)
        self.layernorm = [bool(b) for b in layernorm] if layernorm is not None else [False] * len(layers)
        for i, (n, act, alpha) in enumerate(layers):
            cfg.units[i], cfg.act[i], cfg.alpha[i], cfg.layernorm[i] = n, _lib.ACT[act], alpha, int(self.layernorm[i])
        cfg.head_relu_from, cfg.dtype, cfg.loss, cfg.max_batch = head_relu_from, _lib.DTYPE[dtype], _lib.LOSS[loss], max_batch
        self._h = C.c_void_p()
        _lib.check(self.lib.csb_mlp_create(C.byref(cfg), C.byref(self._h)), "csb_mlp_create")
        self.in_dim, self.out_dim = in_dim, layers[-1][0]
        self.layer_dims: List[Tuple[int, int]] = []
        k = in_dim
        for n, _, _ in layers:
            self.layer_dims.append((k, n))
            k = n
        self.dtype, self.max_batch = dtype, max_batch
        self.n_params = int(self.lib.csb_mlp_param_count(self._h))
        self._loss_buf: Optional[torch.Tensor] = None

    # -- construction helpers -----------------------------------------------------------------------------------
    @classmethod
    def mlp_v1(cls, units: Sequence[int] = (768, 640, 512, 640, 640), act: str = "leakyrelu", alpha: float = 0.15,
               in_dim: int = 124, out_lin: int = 120, out_relu: int = 8, **kw) -> "MLPEngine":
        """MLP_v1 (baseline_models/MLP/training/HPO/baseline_v1/hpo_baseline_v1.py:75-103), best-trial defaults."""
        out = out_lin + out_relu
        layers = [(u, act, alpha) for u in units] + [(out, act, alpha), (out, "none", 0.0)]
        return cls(in_dim, layers, head_relu_from=out_lin, **kw)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.csb_mlp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters ---------------------------------------------------------------------------------------------
    def set_params_flat(self, flat: np.ndarray) -> None:
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.n_params, (flat.size, self.n_params)
        _lib.check(self.lib.csb_mlp_set_params(self._h, flat.ctypes.data), "csb_mlp_set_params")

    def get_params_flat(self) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float32)
        _lib.check(self.lib.csb_mlp_get_params(self._h, out.ctypes.data), "csb_mlp_get_params")
        return out

    def get_grads_flat(self) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float32)
        _lib.check(self.lib.csb_mlp_get_grads(self._h, out.ctypes.data), "csb_mlp_get_grads")
        return out

    def set_params_device(self, flat: torch.Tensor) -> None:
        flat = _f32_cuda(flat, "flat")
        assert flat.numel() == self.n_params
        _lib.check(self.lib.csb_mlp_set_params_device(self._h, flat.data_ptr(), _lib.current_stream_ptr()), "csb_mlp_set_params_device")

    def get_params_device(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = out if out is not None else torch.empty(self.n_params, dtype=torch.float32, device="cuda")
        _lib.check(self.lib.csb_mlp_get_params_device(self._h, out.data_ptr(), _lib.current_stream_ptr()), "csb_mlp_get_params_device")
        return out

    def get_grads_device(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = out if out is not None else torch.empty(self.n_params, dtype=torch.float32, device="cuda")
        _lib.check(self.lib.csb_mlp_get_grads_device(self._h, out.data_ptr(), _lib.current_stream_ptr()), "csb_mlp_get_grads_device")
        return out

    def split_flat(self, flat: np.ndarray) -> List[np.ndarray]:
        """flat blob -> [W0 (in,out), b0, W1, b1, ...] views."""
        out, off = [], 0
        for (k, n), ln in zip(self.layer_dims, self.layernorm):
            out.append(flat[off:off + k * n].reshape(k, n)); off += k * n
            out.append(flat[off:off + n]); off += n
            if ln:                                            # gamma, beta
                out.append(flat[off:off + n]); off += n
                out.append(flat[off:off + n]); off += n
        return out

    @staticmethod
    def keras_to_flat(weights: Sequence[np.ndarray], fused_head: bool = True) -> np.ndarray:
        """Keras ``get_weights()`` list -> the engine's flat blob.  With ``fused_head`` the last two Dense layers
        (linear 120 | relu 8) are concatenated column-wise into one layer, matching keras.layers.Concatenate."""
        ws = [np.asarray(w, dtype=np.float32) for w in weights]
        if fused_head:
            w_lin, b_lin, w_relu, b_relu = ws[-4:]
            ws = ws[:-4] + [np.concatenate([w_lin, w_relu], axis=1), np.concatenate([b_lin, b_relu])]
        return np.concatenate([w.reshape(-1) for w in ws])

    def flat_to_keras(self, flat: np.ndarray, out_lin: Optional[int] = None) -> List[np.ndarray]:
        ws = [w.copy() for w in self.split_flat(flat)]
        if out_lin is not None:
            w, b = ws[-2], ws[-1]
            ws = ws[:-2] + [w[:, :out_lin].copy(), b[:out_lin].copy(), w[:, out_lin:].copy(), b[out_lin:].copy()]
        return ws

    def get_opt_state(self) -> Tuple[np.ndarray, np.ndarray, int]:
        m, v, step = np.empty(self.n_params, np.float32), np.empty(self.n_params, np.float32), C.c_int64()
        _lib.check(self.lib.csb_mlp_get_opt_state(self._h, m.ctypes.data, v.ctypes.data, C.byref(step)), "csb_mlp_get_opt_state")
        return m, v, int(step.value)

    def set_opt_state(self, m: np.ndarray, v: np.ndarray, step: int) -> None:
        m, v = np.ascontiguousarray(m, np.float32), np.ascontiguousarray(v, np.float32)
        _lib.check(self.lib.csb_mlp_set_opt_state(self._h, m.ctypes.data, v.ctypes.data, step), "csb_mlp_set_opt_state")

    def set_norm(self, inp_sub=None, inp_div=None, out_scale=None, loss_w=None) -> None:
        keep = []

        def p(a, n):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.size == n, (a.size, n)
            keep.append(a)
            return a.ctypes.data

        _lib.check(self.lib.csb_mlp_set_norm(self._h, p(inp_sub, self.in_dim), p(inp_div, self.in_dim),
                                             p(out_scale, self.out_dim), p(loss_w, self.out_dim)), "csb_mlp_set_norm")

    def set_input_transform(self, exp_lambda=None, keep=None, clip_lo=None, clip_hi=None) -> None:
        """Generalised input prologue of the online models (``1 - exp(-lambda x)`` columns, pruned columns, clipping), applied by
        every ``normalize_in=True`` call -- see ``csb_mlp_set_input_transform``.  All ``None`` restores the plain normalisation."""
        hold = []

        def p(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.size == self.in_dim, (a.size, self.in_dim)
            hold.append(a)
            return a.ctypes.data

        _lib.check(self.lib.csb_mlp_set_input_transform(self._h, p(exp_lambda), p(keep), p(clip_lo), p(clip_hi)), "csb_mlp_set_input_transform")

    def set_output_mask(self, mask) -> None:
        """0/1 mask over the output columns (the online MLP's ``output_prune``); ``None`` removes it."""
        if mask is None:
            _lib.check(self.lib.csb_mlp_set_output_mask(self._h, None), "csb_mlp_set_output_mask")
            return
        m = np.ascontiguousarray(mask, dtype=np.float32)
        assert m.size == self.out_dim
        _lib.check(self.lib.csb_mlp_set_output_mask(self._h, m.ctypes.data), "csb_mlp_set_output_mask")

    def set_dropout(self, rate: float, seed: int = 0) -> None:
        """``torch.nn.Dropout(rate)`` behind every hidden layer (hsr.py:20-25, online mlp.py:41-45), active in training forwards
        only (``train_step``, ``hsr_train_step``, ``forward(training=True)``); 0 switches it off."""
        _lib.check(self.lib.csb_mlp_set_dropout(self._h, float(rate), int(seed) & 0xFFFFFFFF), "csb_mlp_set_dropout")
        self.dropout = float(rate)

    def dropout_mask(self, layer: int, B: int) -> torch.Tensor:
        """Test hook: the multipliers (0 or 1/(1-p)) the LAST training forward applied behind hidden layer ``layer``, (B, units)."""
        m = torch.empty(B, self.layer_dims[layer][1], dtype=torch.float32, device="cuda")
        _lib.check(self.lib.csb_mlp_debug_dropout_mask(self._h, int(layer), m.data_ptr(), B, _lib.current_stream_ptr()),
                   "csb_mlp_debug_dropout_mask")
        return m

    # -- compute ------------------------------------------------------------------------------------------------
    @staticmethod
    def _flags(normalize_in: bool, denorm_out: bool, keep: bool, training: bool = False) -> int:
        return (_lib.FWD_NORMALIZE_IN if normalize_in else 0) | (_lib.FWD_DENORM_OUT if denorm_out else 0) | \
               (_lib.FWD_KEEP_ACTIVATIONS if keep else 0) | (_lib.FWD_TRAINING if training else 0)

    def forward(self, x: torch.Tensor, normalize_in: bool = False, denorm_out: bool = False,
                keep_activations: bool = False, out: Optional[torch.Tensor] = None, training: bool = False) -> torch.Tensor:
        x = _f32_cuda(x, "x")
        B = x.shape[0]
        y = out if out is not None else torch.empty(B, self.out_dim, dtype=torch.float32, device=x.device)
        if B == 0:
            return y
        _lib.check(self.lib.csb_mlp_forward(self._h, x.data_ptr(), y.data_ptr(), B,
                                            self._flags(normalize_in, denorm_out, keep_activations, training),
                                            _lib.current_stream_ptr()), "csb_mlp_forward")
        return y

    def backward(self, dy: torch.Tensor, need_dx: bool = False) -> Optional[torch.Tensor]:
        dy = _f32_cuda(dy, "dy")
        B = dy.shape[0]
        dx = torch.empty(B, self.in_dim, dtype=torch.float32, device=dy.device) if need_dx else None
        _lib.check(self.lib.csb_mlp_backward(self._h, dy.data_ptr(), dx.data_ptr() if need_dx else None, B,
                                             _lib.current_stream_ptr()), "csb_mlp_backward")
        return dx

    def train_step(self, x: torch.Tensor, y: torch.Tensor, grad_scale: float = 0.0, normalize_in: bool = False,
                   loss_out: Optional[torch.Tensor] = None, fused_opt: bool = False) -> torch.Tensor:
        """forward + loss + backward; gradients land in the engine's gradient buffer.  Returns the device scalar loss.
        ``fused_opt``: ``apply_opt`` follows directly (no all-reduce in between), so the split-partial reduction and the loss
        sum ride in the optimizer launch (CSB_TRAIN_FUSED_OPT); the loss scalar is then valid after ``apply_opt``."""
        x, y = _f32_cuda(x, "x"), _f32_cuda(y, "y")
        if loss_out is None:
            if self._loss_buf is None:
                self._loss_buf = torch.zeros(1, dtype=torch.float32, device=x.device)
            loss_out = self._loss_buf
        _lib.check(self.lib.csb_mlp_train_step(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], grad_scale,
                                               self._flags(normalize_in, False, False) | (_lib.TRAIN_FUSED_OPT if fused_opt else 0),
                                               loss_out.data_ptr(), _lib.current_stream_ptr()), "csb_mlp_train_step")
        return loss_out

    def apply_opt(self, rule: str = "adam_keras", lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999,
                  eps: Optional[float] = None, weight_decay: float = 0.0) -> None:
        if eps is None:
            eps = 1e-8 if rule == "adam_torch" else 1e-7          # Keras / tfa default epsilon is 1e-7
        _lib.check(self.lib.csb_mlp_apply_opt(self._h, _lib.OPT[rule], lr, beta1, beta2, eps, weight_decay,
                                              _lib.current_stream_ptr()), "csb_mlp_apply_opt")

    def grad_buffer(self) -> torch.Tensor:
        """The engine's flat (padded) fp32 gradient buffer as a torch tensor aliasing device memory: the message of the
        data-parallel all-reduce."""
        ptr, n = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.csb_mlp_grad_buffer(self._h, C.byref(ptr), C.byref(n)), "csb_mlp_grad_buffer")
        return torch.as_tensor(_DevPtr(ptr.value, n.value), device="cuda")

    # -- data parallelism over NVLink peer memory (csb_mlp_dp_*) -------------------------------------------------------
    def dp_export(self) -> bytes:
        """Move the gradient buffer into an IPC-exportable slab; returns its handle (gather the handles of all ranks, then ``dp_attach``)."""
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(self.lib.csb_mlp_dp_export(self._h, buf), "csb_mlp_dp_export")
        return bytes(buf.raw)

    def dp_attach(self, rank: int, world: int, handles: Sequence[bytes]) -> None:
        assert len(handles) == world and all(len(hd) == _lib.IPC_HANDLE_BYTES for hd in handles)
        blob = C.create_string_buffer(b"".join(handles), world * _lib.IPC_HANDLE_BYTES)
        _lib.check(self.lib.csb_mlp_dp_attach(self._h, rank, world, blob), "csb_mlp_dp_attach")

    def dp_step(self, rule: str = "adam_keras", lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, eps: Optional[float] = None,
                weight_decay: float = 0.0, loss_out: Optional[torch.Tensor] = None) -> None:
        """Cross-rank gradient sum + optimizer + bf16 weight copies in one kernel over peer memory; follows
        ``train_step(..., fused_opt=True)`` on every rank.  ``loss_out`` receives the loss of the global batch."""
        if eps is None:
            eps = 1e-8 if rule == "adam_torch" else 1e-7
        _lib.check(self.lib.csb_mlp_dp_step(self._h, _lib.OPT[rule], lr, beta1, beta2, eps, weight_decay,
                                            loss_out.data_ptr() if loss_out is not None else None, _lib.current_stream_ptr()), "csb_mlp_dp_step")

    def train_step_host(self, x_host: torch.Tensor, y_host: torch.Tensor, rule: str = "adam_keras", lr: float = 1e-3,
                        beta1: float = 0.9, beta2: float = 0.999, eps: Optional[float] = None,
                        weight_decay: float = 0.0, grad_scale: float = 0.0, normalize_in: bool = False) -> float:
        """End-to-end step from HOST tensors (pinned recommended): H2D + fwd/loss/bwd + optimizer + D2H loss."""
        assert not x_host.is_cuda and not y_host.is_cuda and x_host.dtype == torch.float32 and y_host.dtype == torch.float32
        x_host, y_host = x_host.contiguous(), y_host.contiguous()
        if eps is None:
            eps = 1e-8 if rule == "adam_torch" else 1e-7          # Keras / tfa default epsilon is 1e-7
        loss = C.c_float()
        _lib.check(self.lib.csb_mlp_train_step_host(self._h, x_host.data_ptr(), y_host.data_ptr(), x_host.shape[0],
                                                    grad_scale, self._flags(normalize_in, False, False), _lib.OPT[rule],
                                                    lr, beta1, beta2, eps, weight_decay, C.byref(loss),
                                                    _lib.current_stream_ptr()), "csb_mlp_train_step_host")
        return float(loss.value)

    def train_step_host_async(self, x_host: torch.Tensor, y_host: torch.Tensor, loss_slot: torch.Tensor, rule: str = "adam_keras",
                              lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, eps: Optional[float] = None,
                              weight_decay: float = 0.0, grad_scale: float = 0.0, normalize_in: bool = False) -> None:
        """``train_step_host`` without the final synchronisation: the H2D copies go to one of two staging slots on the engine's
        copy stream (overlapping the previous step's compute) and the loss lands in ``loss_slot`` (a pinned CPU float32
        tensor) once the current stream has passed this step.  Host buffers may be reused after two further calls."""
        assert not x_host.is_cuda and not y_host.is_cuda and x_host.dtype == torch.float32 and y_host.dtype == torch.float32
        assert x_host.is_contiguous() and y_host.is_contiguous() and loss_slot.is_pinned()
        if eps is None:
            eps = 1e-8 if rule == "adam_torch" else 1e-7
        _lib.check(self.lib.csb_mlp_train_step_host_async(self._h, x_host.data_ptr(), y_host.data_ptr(), x_host.shape[0],
                                                          grad_scale, self._flags(normalize_in, False, False), _lib.OPT[rule],
                                                          lr, beta1, beta2, eps, weight_decay, loss_slot.data_ptr(),
                                                          _lib.current_stream_ptr()), "csb_mlp_train_step_host_async")

    def stage_host_batch(self, x_host: torch.Tensor, y_host: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Enqueue the H2D copies of a host batch into one of the engine's two staging slots (internal copy stream; the
        current stream waits for them) and return CUDA tensors aliasing the slot.  Pair with ``release_staged()``."""
        assert not x_host.is_cuda and not y_host.is_cuda and x_host.dtype == torch.float32 and y_host.dtype == torch.float32
        assert x_host.is_contiguous() and y_host.is_contiguous()
        B = x_host.shape[0]
        xd, yd = C.c_void_p(), C.c_void_p()
        _lib.check(self.lib.csb_mlp_stage_host_batch(self._h, x_host.data_ptr(), y_host.data_ptr(), B, _lib.current_stream_ptr(),
                                                     C.byref(xd), C.byref(yd)), "csb_mlp_stage_host_batch")
        x = torch.as_tensor(_DevPtr(xd.value, B * self.in_dim), device="cuda").view(B, self.in_dim)
        y = torch.as_tensor(_DevPtr(yd.value, B * self.out_dim), device="cuda").view(B, self.out_dim)
        return x, y

    def release_staged(self) -> None:
        _lib.check(self.lib.csb_mlp_release_staged(self._h, _lib.current_stream_ptr()), "csb_mlp_release_staged")

    def forward_host(self, x_host: torch.Tensor, normalize_in: bool = False, denorm_out: bool = False) -> torch.Tensor:
        x_host = x_host.contiguous()
        y = torch.empty(x_host.shape[0], self.out_dim, dtype=torch.float32)
        _lib.check(self.lib.csb_mlp_forward_host(self._h, x_host.data_ptr(), y.data_ptr(), x_host.shape[0],
                                                 self._flags(normalize_in, denorm_out, False), _lib.current_stream_ptr()),
                   "csb_mlp_forward_host")
        return y

    def batch_metrics(self, pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Device fp64 vector [sum (p-y)^2, sum |p-y|, rows with argmax(p) == argmax(y), elements, rows] of one batch: the
        sufficient statistics of Keras' ``metrics=['mse','mae','accuracy']`` (``csb_batch_metrics``); no synchronisation."""
        pred, y = _f32_cuda(pred, "pred"), _f32_cuda(y, "y")
        assert pred.shape == y.shape and pred.dim() == 2
        if getattr(self, "_metrics_scratch", None) is None:
            self._metrics_scratch = torch.zeros(_lib.BATCH_METRICS_SCRATCH, dtype=torch.float64, device=pred.device)
        out = torch.empty(5, dtype=torch.float64, device=pred.device)
        _lib.check(self.lib.csb_batch_metrics(pred.data_ptr(), y.data_ptr(), pred.shape[0], pred.shape[1], out.data_ptr(),
                                              self._metrics_scratch.data_ptr(), _lib.current_stream_ptr()), "csb_batch_metrics")
        return out

    def profile(self, enable: bool = True) -> None:
        _lib.check(self.lib.csb_mlp_profile(self._h, 1 if enable else 0), "csb_mlp_profile")

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """kernel kind -> (accumulated device milliseconds, launches) since profile(True); synchronises."""
        n = int(self.lib.csb_profile_kind_count())
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(self.lib.csb_mlp_profile_read(self._h, ms, cnt, n), "csb_mlp_profile_read")
        return {self.lib.csb_profile_kind_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}

    @property
    def launch_count(self) -> int:
        return int(self.lib.csb_mlp_launch_count(self._h))


def hsr_train_step(mean: MLPEngine, logprec: MLPEngine, x: torch.Tensor, y: torch.Tensor, mle: bool, loss_out: torch.Tensor,
                   scratch: torch.Tensor, rule: str = "adam_torch", lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999,
                   eps: float = 1e-8, wd_mean: float = 0.0, wd_logprec: float = 0.0, normalize_in: bool = False,
                   apply_opt: bool = True) -> None:
    """``csb_hsr_train_step``: one step of both heteroskedastic-regression networks (forward, MSE / Gaussian-NLL loss with the
    reference's clip, backward, per-group L2 optimizer) entirely inside the engine; the loss lands in ``loss_out`` (a one-element
    CUDA fp32 tensor, e.g. a slice of a per-epoch loss array) without any host synchronisation.  ``apply_opt=False`` stops after the
    backward passes (data parallelism: all-reduce ``grad_buffer()`` of each engine, then ``apply_opt`` with its own decay)."""
    x, y = _f32_cuda(x, "x"), _f32_cuda(y, "y")
    assert loss_out.is_cuda and loss_out.dtype == torch.float32 and scratch.is_cuda and scratch.dtype == torch.float64
    assert scratch.numel() >= _lib.BATCH_METRICS_SCRATCH
    _lib.check(mean.lib.csb_hsr_train_step(mean._h, logprec._h, x.data_ptr(), y.data_ptr(), x.shape[0], int(bool(mle)),
                                           (_lib.FWD_NORMALIZE_IN if normalize_in else 0) | (0 if apply_opt else _lib.HSR_NO_OPT),
                                           _lib.OPT[rule], lr, beta1, beta2, eps,
                                           wd_mean, wd_logprec, loss_out.data_ptr(), scratch.data_ptr(), _lib.current_stream_ptr()),
               "csb_hsr_train_step")


class CNNEngine:
    """The ResNet-1D column emulator (baseline_models/CNN/training/hpo_train.py:131-200) bound to the csb_cnn_* C ABI.
    Tensors are channels-last: x (B, 60, 6), y / predictions (B, 60, 10)."""

    def __init__(self, depth: int = 12, width: int = 406, kernel: int = 3, in_ch: int = 6, out_ch: int = 10, out_lin: int = 2,
                 levels: int = 60, act: str = "relu", pre_out_act: str = "elu", loss: str = "mae", max_batch: int = 4096,
                 dtype: str = "bf16"):
        self.lib = _lib.load()
        cfg = _lib.CnnCfg(depth, width, kernel, in_ch, out_ch, out_lin, levels, _lib.ACT[act], _lib.ACT[pre_out_act],
                          _lib.DTYPE[dtype], _lib.LOSS[loss], max_batch)
        self.dtype = dtype
        self._h = C.c_void_p()
        _lib.check(self.lib.csb_cnn_create(C.byref(cfg), C.byref(self._h)), "csb_cnn_create")
        self.depth, self.width, self.kernel, self.in_ch, self.out_ch, self.out_lin, self.levels = depth, width, kernel, in_ch, out_ch, out_lin, levels
        self.n_params = int(self.lib.csb_cnn_param_count(self._h))
        self._loss_buf: Optional[torch.Tensor] = None

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.csb_cnn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def keras_to_flat(weights: Sequence[np.ndarray]) -> np.ndarray:
        """Keras ``get_weights()`` -> flat blob (the two Dense heads concatenated column-wise)."""
        ws = [np.asarray(w, dtype=np.float32) for w in weights]
        w_lin, b_lin, w_relu, b_relu = ws[-4:]
        ws = ws[:-4] + [np.concatenate([w_lin, w_relu], axis=1), np.concatenate([b_lin, b_relu])]
        return np.concatenate([w.reshape(-1) for w in ws])

    def shapes(self) -> List[Tuple[int, ...]]:
        out, c = [], self.in_ch
        for _ in range(self.depth):
            out += [(self.kernel, c, self.width), (self.width,), (self.kernel, self.width, self.width), (self.width,), (1, c, self.width), (self.width,)]
            c = self.width
        return out + [(1, c, self.out_ch), (self.out_ch,), (self.out_ch, self.out_ch), (self.out_ch,)]

    def split_flat(self, flat: np.ndarray) -> List[np.ndarray]:
        out, off = [], 0
        for shp in self.shapes():
            n = int(np.prod(shp))
            out.append(flat[off:off + n].reshape(shp)); off += n
        return out

    def set_params_flat(self, flat: np.ndarray) -> None:
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.n_params, (flat.size, self.n_params)
        _lib.check(self.lib.csb_cnn_set_params(self._h, flat.ctypes.data), "csb_cnn_set_params")

    def get_params_flat(self) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float32)
        _lib.check(self.lib.csb_cnn_get_params(self._h, out.ctypes.data), "csb_cnn_get_params")
        return out

    def get_grads_flat(self) -> np.ndarray:
        out = np.empty(self.n_params, dtype=np.float32)
        _lib.check(self.lib.csb_cnn_get_grads(self._h, out.ctypes.data), "csb_cnn_get_grads")
        return out

    def set_dropout(self, rate: float, seed: int = 0) -> None:
        """Dropout behind the two ReLUs of every block in ``train_step`` (the reference: 0.175); ``forward`` never drops."""
        _lib.check(self.lib.csb_cnn_set_dropout(self._h, float(rate), int(seed) & 0xFFFFFFFF), "csb_cnn_set_dropout")

    def debug_hidden(self, which: int, block: int, B: int) -> torch.Tensor:
        """Test hook: hidden activation after conv1 / conv2 (``which`` = 1 / 2) of ``block`` from the last step, fp32 (B, levels, width)."""
        out = torch.empty(B, self.levels, self.width, dtype=torch.float32, device="cuda")
        _lib.check(self.lib.csb_cnn_debug_read_hidden(self._h, which, block, out.data_ptr(), B, _lib.current_stream_ptr()), "csb_cnn_debug_read_hidden")
        return out

    def set_loss_weights(self, w) -> None:
        w = np.ascontiguousarray(w, dtype=np.float32)
        assert w.size == self.out_ch
        _lib.check(self.lib.csb_cnn_set_loss_weights(self._h, w.ctypes.data), "csb_cnn_set_loss_weights")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _f32_cuda(x, "x")
        y = torch.empty(x.shape[0], self.levels, self.out_ch, dtype=torch.float32, device=x.device)
        _lib.check(self.lib.csb_cnn_forward(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], _lib.current_stream_ptr()), "csb_cnn_forward")
        return y

    def train_step(self, x: torch.Tensor, y: torch.Tensor, grad_scale: float = 0.0) -> torch.Tensor:
        x, y = _f32_cuda(x, "x"), _f32_cuda(y, "y")
        if self._loss_buf is None:
            self._loss_buf = torch.zeros(1, dtype=torch.float32, device=x.device)
        _lib.check(self.lib.csb_cnn_train_step(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], grad_scale, self._loss_buf.data_ptr(),
                                               _lib.current_stream_ptr()), "csb_cnn_train_step")
        return self._loss_buf

    def backward(self, y_pred: torch.Tensor, dy: torch.Tensor) -> None:
        """Parameter gradients for an upstream gradient ``dy`` w.r.t. the output of the preceding ``forward`` on the same batch."""
        y_pred, dy = _f32_cuda(y_pred, "y_pred"), _f32_cuda(dy, "dy")
        _lib.check(self.lib.csb_cnn_backward(self._h, y_pred.data_ptr(), dy.data_ptr(), dy.shape[0], _lib.current_stream_ptr()), "csb_cnn_backward")

    def set_params_device(self, flat: torch.Tensor) -> None:
        flat = _f32_cuda(flat, "flat")
        assert flat.numel() == self.n_params
        _lib.check(self.lib.csb_cnn_set_params_device(self._h, flat.data_ptr(), _lib.current_stream_ptr()), "csb_cnn_set_params_device")

    def get_grads_device(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = out if out is not None else torch.empty(self.n_params, dtype=torch.float32, device="cuda")
        _lib.check(self.lib.csb_cnn_get_grads_device(self._h, out.data_ptr(), _lib.current_stream_ptr()), "csb_cnn_get_grads_device")
        return out

    def apply_opt(self, rule: str = "adam_keras", lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999, eps: Optional[float] = None,
                  weight_decay: float = 0.0) -> None:
        if eps is None:
            eps = 1e-8 if rule == "adam_torch" else 1e-7
        _lib.check(self.lib.csb_cnn_apply_opt(self._h, _lib.OPT[rule], lr, beta1, beta2, eps, weight_decay, _lib.current_stream_ptr()),
                   "csb_cnn_apply_opt")

    def grad_buffer(self) -> torch.Tensor:
        ptr, n = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.csb_cnn_grad_buffer(self._h, C.byref(ptr), C.byref(n)), "csb_cnn_grad_buffer")
        return torch.as_tensor(_DevPtr(ptr.value, n.value), device="cuda")

    @property
    def launch_count(self) -> int:
        return int(self.lib.csb_cnn_launch_count(self._h))
