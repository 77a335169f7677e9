"""Training driver on top of ``MLPEngine``: what ``model.fit`` does per batch in the reference
(baseline_models/MLP/.../step2_retrain.py:280-285; HSR/training/hsr.py:122-140), one process per GPU.

Data parallelism: columns are i.i.d., so each rank takes B/N columns, the flat gradient buffer
is summed with one NCCL all-reduce over NVLink (the reference's only collective is the DDP gradient all-reduce of its
online trainer, online_testing/.../train_mlp_h5loader.py:195-207), and every rank applies the same optimizer step.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from .engine import MLPEngine


def glorot_uniform_flat(layer_dims: Sequence[tuple], seed: int = 0, layernorm: Optional[Sequence[bool]] = None) -> np.ndarray:
    """Keras default initialisation (glorot_uniform kernels, zero biases; LayerNorm gamma 1, beta 0) as the engine's flat blob."""
    rng = np.random.default_rng(seed)
    parts = []
    for i, (k, n) in enumerate(layer_dims):
        lim = math.sqrt(6.0 / (k + n))
        parts += [rng.uniform(-lim, lim, size=(k, n)).astype(np.float32).reshape(-1), np.zeros(n, np.float32)]
        if layernorm is not None and layernorm[i]:
            parts += [np.ones(n, np.float32), np.zeros(n, np.float32)]
    return np.concatenate(parts)


def cyclical_lr(step: int, initial_lr: float = 2.5e-4, max_lr: float = 2.5e-3, step_size: float = 2.0) -> float:
    """tfa CyclicalLearningRate(scale_fn=1/2**(x-1), scale_mode='cycle') as configured at hpo_baseline_v1.py:105-114."""
    cycle = math.floor(1 + step / (2 * step_size))
    x = abs(step / step_size - 2 * cycle + 1)
    return initial_lr + (max_lr - initial_lr) * max(0.0, 1 - x) / (2.0 ** (cycle - 1))


class Trainer:
    """``step(x, y)`` = H2D (if host tensors) -> forward + loss + backward -> gradient all-reduce -> optimizer.

    Host batches are fed through the engine's two staging slots on its copy stream, so with ``sync=False`` the H2D copy of
    batch i+1 overlaps the compute of batch i (the role of ``tf.data`` prefetch / ``DataLoader(pin_memory=True)`` in the
    reference's drivers); every step still performs its own H2D copy and the D2H read of its loss."""

    def __init__(self, engine: MLPEngine, rule: str = "adam_keras", lr: float | Callable[[int], float] = 1e-3,
                 beta1: float = 0.9, beta2: float = 0.999, eps: Optional[float] = None, weight_decay: float = 0.0,
                 process_group=None):
        self.engine, self.rule, self.lr = engine, rule, lr
        self.beta1, self.beta2, self.eps, self.weight_decay = beta1, beta2, eps, weight_decay
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.iteration = 0
        self.device = getattr(engine, "device", "cuda")
        self._grad = engine.grad_buffer() if self.world > 1 else None
        self._x = self._y = None
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._loss_host = None            # pinned ring of loss slots for the asynchronous host-fed path

    def _lr(self) -> float:
        return float(self.lr(self.iteration)) if callable(self.lr) else float(self.lr)

    def _loss_slot(self) -> torch.Tensor:
        if self._loss_host is None:
            self._loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        return self._loss_host[self.iteration % 4:self.iteration % 4 + 1]

    def synchronize(self) -> None:
        """Wait for every step enqueued so far (the loss slots returned by ``step(..., sync=False)`` are valid afterwards)."""
        torch.cuda.current_stream().synchronize()

    def step(self, x: torch.Tensor, y: torch.Tensor, normalize_in: bool = False, return_loss: bool = True, sync: bool = True):
        """x (B_local, in_dim), y (B_local, out_dim): CUDA tensors, or (pinned, contiguous) host tensors that are copied first.

        Device tensors: returns the loss of the GLOBAL batch as a Python float when ``return_loss`` (one D2H read), else the
        device scalar holding this rank's share.
        Host tensors: the loss is always read back (4-byte D2H into a pinned slot).  ``sync=True`` waits and returns a float;
        ``sync=False`` returns the pinned one-element tensor, valid after ``synchronize()`` (or once two further steps have
        been issued) -- x and y must stay untouched for the same span."""
        eng = self.engine
        B = x.shape[0]
        lr = self._lr()
        on_device = x.device.type == torch.device(self.device).type
        slot = None
        if not on_device:
            slot = self._loss_slot()
            if self.world == 1:
                eng.train_step_host_async(x, y, slot, rule=self.rule, lr=lr, beta1=self.beta1, beta2=self.beta2, eps=self.eps,
                                          weight_decay=self.weight_decay, normalize_in=normalize_in)
                self.iteration += 1
                if not sync:
                    return slot
                self.synchronize()
                return float(slot.item())
            x, y = eng.stage_host_batch(x, y)
        scale = 1.0 / (B * self.world * eng.out_dim)            # global-mean MSE, as Keras computes on the global batch
        eng.train_step(x, y, grad_scale=scale, normalize_in=normalize_in, loss_out=self._loss, fused_opt=self.world == 1)
        if slot is not None:
            eng.release_staged()
        if self.world > 1:
            torch.distributed.all_reduce(self._grad, group=self.pg)
            if return_loss or slot is not None:
                torch.distributed.all_reduce(self._loss, group=self.pg)
        eng.apply_opt(self.rule, lr=lr, beta1=self.beta1, beta2=self.beta2, eps=self.eps, weight_decay=self.weight_decay)
        self.iteration += 1
        if slot is not None:
            slot.copy_(self._loss, non_blocking=True)
            if not sync:
                return slot
            self.synchronize()
            return float(slot.item())
        return float(self._loss.item()) if return_loss else self._loss

    # ------------------------------------------------------------------------------------------------------------------ model.fit
    def save_checkpoint(self, path: str) -> None:
        """Parameters + optimizer state + step counters: the content of Keras' ``ModelCheckpoint(save_weights_only=False)`` file."""
        m, v, step = self.engine.get_opt_state()
        with open(path, "wb") as f:                 # a file object: np.savez would append ".npz" to a bare path
            np.savez(f, params=self.engine.get_params_flat(), m=m, v=v, step=np.int64(step), iteration=np.int64(self.iteration))

    def load_checkpoint(self, path: str) -> None:
        ck = np.load(path)
        self.engine.set_params_flat(ck["params"])
        self.engine.set_opt_state(ck["m"], ck["v"], int(ck["step"]))
        self.iteration = int(ck["iteration"])

    def evaluate(self, data) -> float:
        """Mean squared error over ``data`` (an iterable of ``(x, y)``; exact mean over all elements): Keras' ``val_loss`` for
        ``loss='mse'`` (hpo_baseline_v1.py:127-129).  One D2H read at the end."""
        se, n = None, 0
        for x, y in data:
            d = self.engine.forward(x) - y
            s = (d * d).sum()
            se = s if se is None else se + s
            n += d.numel()
        return float(se.item()) / max(n, 1) if se is not None else float("nan")

    def fit(self, train, epochs: int, validation_data=None, checkpoint_best: Optional[str] = None, checkpoint_last: Optional[str] = None,
            csv_log: Optional[str] = None, early_stopping_patience: Optional[int] = None, initial_epoch: int = 0, verbose: int = 2) -> dict:
        """``model.fit(tds, epochs=.., validation_data=tds_val, callbacks=[checkpoint_best, checkpoint_last, csv_logger, earlystop])`` as
        the reference's retraining script drives it (baseline_v1/step2_retrain/step2_retrain.py:252-286):

        * ``train`` / ``validation_data``: objects with ``.epoch(e)`` yielding ``(x, y)`` (``NpyColumnStream``: reshuffled every
          epoch like ``shuffle(reshuffle_each_iteration=True)``) or plain re-iterable collections of ``(x, y)``;
        * ``loss`` of an epoch = mean of its batch losses (what Keras prints), accumulated on the device: one D2H read per epoch;
        * ``checkpoint_best``: saved when ``val_loss`` improves (``ModelCheckpoint(monitor='val_loss', save_best_only=True)``);
          ``checkpoint_last``: saved every epoch; ``csv_log``: ``epoch,loss,val_loss`` rows appended (``CSVLogger(append=True)``);
          ``early_stopping_patience``: stop after that many epochs without a new best ``val_loss`` (``EarlyStopping('val_loss', patience)``).

        Returns ``{"loss": [...], "val_loss": [...], "stopped_epoch": e or None}`` (Keras' ``History.history`` plus the stop epoch)."""
        history = {"loss": [], "val_loss": [], "stopped_epoch": None}
        best, wait = float("inf"), 0
        if csv_log is not None:
            import os
            if not os.path.exists(csv_log) or os.path.getsize(csv_log) == 0:
                with open(csv_log, "a") as f:
                    f.write("epoch,loss,val_loss\n")
        for epoch in range(initial_epoch, epochs):
            batches = train.epoch(epoch) if hasattr(train, "epoch") else train
            total, nb = None, 0
            for x, y in batches:
                l = self.step(x, y, return_loss=False)
                l = l if isinstance(l, torch.Tensor) else torch.as_tensor(l)
                total = l.detach().clone().reshape(()) if total is None else total + l.detach().reshape(())
                nb += 1
            if self.world > 1 and total is not None:                 # every rank holds its share of the global-mean loss
                torch.distributed.all_reduce(total, group=self.pg)
            loss = float(total.item()) / nb if nb else float("nan")
            history["loss"].append(loss)
            val = None
            if validation_data is not None:
                vb = validation_data.epoch(epoch) if hasattr(validation_data, "epoch") else validation_data
                val = self.evaluate(vb)
                history["val_loss"].append(val)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - loss: {loss:.6g}" + (f" - val_loss: {val:.6g}" if val is not None else ""), flush=True)
            if csv_log is not None:
                with open(csv_log, "a") as f:
                    f.write(f"{epoch},{loss!r},{'' if val is None else repr(val)}\n")
            if checkpoint_last is not None:
                self.save_checkpoint(checkpoint_last)
            if val is not None:
                if val < best:
                    best, wait = val, 0
                    if checkpoint_best is not None:
                        self.save_checkpoint(checkpoint_best)
                else:
                    wait += 1
                    if early_stopping_patience is not None and wait >= early_stopping_patience:
                        history["stopped_epoch"] = epoch
                        break
        return history

