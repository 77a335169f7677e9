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


def ed_step_lr(epoch: int, lr_init: float = 1e-4, drop: float = 5.0, every: int = 7) -> float:
    """The encoder-decoder's ``LearningRateScheduler`` (ClimSIM_ED_1_3_train.py:98-122): the learning rate is divided by 5 after
    every 7th epoch -- lr_init for epochs 0-6, /5 for 7-13, /25, /125, /625, /3125 for 35-41 (the reference's function returns None
    from epoch 42 on, which Keras rejects; the run has 40 epochs)."""
    return lr_init / (drop ** (epoch // every))


class Trainer:
    """``step(x, y)`` = H2D (if host tensors) -> forward + loss + backward -> gradient all-reduce -> optimizer.

    Host batches are fed through the engine's two staging slots on its copy stream, so with ``sync=False`` the H2D copy of
    batch i+1 overlaps the compute of batch i (the role of ``tf.data`` prefetch / ``DataLoader(pin_memory=True)`` in the
    reference's drivers); every step still performs its own H2D copy and the D2H read of its loss.

    Data parallelism: every rank must start from the same parameters -- rank 0's are broadcast here, as DDP does at construction
    (train_mlp_h5loader.py:195-207)."""

    _NO_WEIGHT_DECAY = ("adam_keras", "adam", "rmsprop")      # rules whose reference optimizers have no decay term

    def __init__(self, engine: MLPEngine, rule: str = "adam_keras", lr: float | Callable[[int], float] = 1e-3,
                 beta1: float = 0.9, beta2: float = 0.999, eps: Optional[float] = None, weight_decay: float = 0.0,
                 process_group=None):
        if weight_decay != 0.0 and rule in self._NO_WEIGHT_DECAY:
            raise ValueError(f"optimizer rule {rule!r} has no weight-decay term (Keras Adam / RMSprop); use 'adam_torch', 'sgd' or 'radam'")
        self.engine, self.rule, self.lr = engine, rule, lr
        self.beta1, self.beta2, self.eps, self.weight_decay = beta1, beta2, eps, weight_decay
        self.pg = process_group
        self.world, self.rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        self.iteration = 0
        self.device = getattr(engine, "device", "cuda")
        self._grad = engine.grad_buffer() if self.world > 1 else None
        self._x = self._y = None
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._loss_host = None            # pinned ring of loss slots for the asynchronous host-fed path
        self._peer = False                # gradient exchange fused with the optimizer over NVLink peer memory (csb_mlp_dp_step)
        if self.world > 1:
            self._broadcast_params()
            self._try_peer_exchange()

    def _try_peer_exchange(self) -> None:
        """One node, bf16 engine: map the peers' gradient slabs (CUDA IPC) so that the all-reduce rides inside the optimizer kernel.
        Anything else (fp32 parity engine, no peer access, the CPU test double, CSB_DP_NCCL=1) keeps ncclAllReduce + apply_opt.
        The decision is collective: every rank takes the peer path or none does."""
        import os
        eng, dist = self.engine, torch.distributed
        ok = (hasattr(eng, "dp_export") and getattr(eng, "dtype", None) == "bf16" and torch.device(self.device).type == "cuda"
              and os.environ.get("CSB_DP_NCCL") is None and self.world <= 8)
        handle = None
        if ok:
            try:
                handle = eng.dp_export()
            except Exception:
                ok = False
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handle if ok else None, group=self.pg)
        if any(g is None for g in gathered):
            return
        try:
            eng.dp_attach(self.rank, self.world, gathered)
            attached = 1
        except Exception:
            attached = 0
        flag = torch.tensor([attached], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)
        self._peer = bool(flag.item())
        self._grad = None if self._peer else eng.grad_buffer()       # (dp_export moved the gradient buffer into the slab)

    def _broadcast_params(self) -> None:
        eng, dist = self.engine, torch.distributed
        src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
        if hasattr(eng, "get_params_device") and torch.device(self.device).type == "cuda":
            flat = eng.get_params_device()
            dist.broadcast(flat, src=src, group=self.pg)
            eng.set_params_device(flat)
        else:
            flat = torch.from_numpy(np.ascontiguousarray(eng.get_params_flat()))
            dist.broadcast(flat, src=src, group=self.pg)
            eng.set_params_flat(flat.numpy())

    def _lr(self) -> float:
        return float(self.lr(self.iteration)) if callable(self.lr) else float(self.lr)

    def _loss_slot(self) -> torch.Tensor:
        if self._loss_host is None:
            self._loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        return self._loss_host[self.iteration % 4:self.iteration % 4 + 1]

    def _on_device(self, t: torch.Tensor) -> bool:
        return t.device.type == torch.device(self.device).type

    def synchronize(self) -> None:
        """Wait for every step enqueued so far (the loss slots returned by ``step(..., sync=False)`` are valid afterwards)."""
        torch.cuda.current_stream().synchronize()

    def _step_local(self, x: torch.Tensor, y: torch.Tensor, normalize_in: bool = False) -> torch.Tensor:
        """One training step; returns the DEVICE scalar holding THIS RANK'S share of the global-mean loss (no loss all-reduce, no
        host synchronisation).  Host batches go through the staging slots."""
        eng = self.engine
        B = x.shape[0]
        lr = self._lr()
        staged = not self._on_device(x)
        if staged:
            x, y = eng.stage_host_batch(x, y)
        scale = 1.0 / (B * self.world * eng.out_dim)            # global-mean MSE, as Keras computes on the global batch
        eng.train_step(x, y, grad_scale=scale, normalize_in=normalize_in, loss_out=self._loss, fused_opt=self.world == 1 or self._peer)
        if staged:
            eng.release_staged()
        if self._peer:                                          # cross-rank sum + optimizer in one kernel; self._loss = GLOBAL loss
            eng.dp_step(self.rule, lr=lr, beta1=self.beta1, beta2=self.beta2, eps=self.eps, weight_decay=self.weight_decay, loss_out=self._loss)
        else:
            if self.world > 1:
                torch.distributed.all_reduce(self._grad, group=self.pg)
            eng.apply_opt(self.rule, lr=lr, beta1=self.beta1, beta2=self.beta2, eps=self.eps, weight_decay=self.weight_decay)
        self.iteration += 1
        return self._loss

    def step(self, x: torch.Tensor, y: torch.Tensor, normalize_in: bool = False, return_loss: bool = True, sync: bool = True):
        """x (B_local, in_dim), y (B_local, out_dim): CUDA tensors, or (pinned, contiguous) host tensors that are copied first.

        Device tensors: returns the loss of the GLOBAL batch as a Python float when ``return_loss`` (one D2H read), else the
        device scalar holding this rank's share.
        Host tensors: the loss is always read back (4-byte D2H into a pinned slot).  ``sync=True`` waits and returns a float;
        ``sync=False`` returns the pinned one-element tensor, valid after ``synchronize()`` (or once two further steps have
        been issued) -- x and y must stay untouched for the same span."""
        eng = self.engine
        on_device = self._on_device(x)
        slot = None
        if not on_device:
            slot = self._loss_slot()
            if self.world == 1:
                eng.train_step_host_async(x, y, slot, rule=self.rule, lr=self._lr(), beta1=self.beta1, beta2=self.beta2, eps=self.eps,
                                          weight_decay=self.weight_decay, normalize_in=normalize_in)
                self.iteration += 1
                if not sync:
                    return slot
                self.synchronize()
                return float(slot.item())
        loss = self._step_local(x, y, normalize_in)
        if self.world > 1 and not self._peer and (return_loss or slot is not None):
            torch.distributed.all_reduce(loss, group=self.pg)
        if slot is not None:
            slot.copy_(loss, non_blocking=True)
            if not sync:
                return slot
            self.synchronize()
            return float(slot.item())
        return float(loss.item()) if return_loss else loss

    # ------------------------------------------------------------------------------------------------------------------ model.fit
    def load_keras_h5(self, path: str, load_optimizer: bool = True, fused_head: bool = True) -> dict:
        """Continue from one of the reference's Keras ``.h5`` checkpoints (step2_retrain.py:252-262 ``ModelCheckpoint``; the `sw_continue`
        workflow): parameters, and -- when the file has them and ``load_optimizer`` -- the optimizer's (m, v) slots and iteration
        count, so that the next ``step`` is the next step of that run.  ``fused_head``: MLP_v1's two output layers are one layer here."""
        from .keras_h5 import read_keras_h5
        ck = read_keras_h5(path)
        to_flat = type(self.engine).keras_to_flat
        self.engine.set_params_flat(to_flat(ck["weights"], fused_head=fused_head))
        opt = ck.get("optimizer")
        if load_optimizer and opt and opt["m"]:
            self.engine.set_opt_state(to_flat(opt["m"], fused_head=fused_head), to_flat(opt["v"], fused_head=fused_head), opt["iterations"])
            self.iteration = opt["iterations"]
        return ck

    def save_checkpoint(self, path: str) -> None:
        """Parameters + optimizer state + step counters: the CONTENT of Keras' ``ModelCheckpoint(save_weights_only=False)`` file, as an
        ``.npz`` of flat fp32 blobs in ``get_weights()`` order.  The format is this library's own -- the reference's artefact is a
        Keras ``.h5`` (h5py is not available here); ``MLP.keras_weights()`` / ``MLPEngine.flat_to_keras`` give the ``set_weights``
        list for a model on the Keras side, ``MLP.load_keras_weights`` the way back."""
        m, v, step = self.engine.get_opt_state()
        with open(path, "wb") as f:                 # a file object: np.savez would append ".npz" to a bare path
            np.savez(f, params=self.engine.get_params_flat(), m=m, v=v, step=np.int64(step), iteration=np.int64(self.iteration))

    def load_checkpoint(self, path: str) -> None:
        ck = np.load(path)
        self.engine.set_params_flat(ck["params"])
        self.engine.set_opt_state(ck["m"], ck["v"], int(ck["step"]))
        self.iteration = int(ck["iteration"])

    METRICS = ("loss", "mse", "mae", "accuracy")

    def _batch_metrics(self, pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """[sum of squared errors, sum of absolute errors, rows whose argmax agrees, elements, rows] of one batch as a device
        vector (fp64): the sufficient statistics of Keras' ``metrics=['mse', 'mae', 'accuracy']`` (hpo_baseline_v1.py:127-129;
        with a 128-wide regression output Keras resolves 'accuracy' to categorical accuracy)."""
        eng = self.engine
        if hasattr(eng, "batch_metrics"):
            return eng.batch_metrics(pred, y)
        d = (pred - y).double()
        hits = (pred.argmax(dim=1) == y.argmax(dim=1)).double().sum()
        return torch.stack([(d * d).sum(), d.abs().sum(), hits, torch.tensor(float(d.numel()), dtype=torch.float64, device=d.device),
                            torch.tensor(float(d.shape[0]), dtype=torch.float64, device=d.device)])

    def evaluate(self, data, return_dict: bool = False):
        """``model.evaluate``: exact means over all elements of ``data`` (an iterable of ``(x, y)``, CUDA or pinned host tensors) --
        Keras' ``val_loss`` for ``loss='mse'`` plus ``mse``, ``mae``, ``accuracy`` (hpo_baseline_v1.py:127-129).  Under data
        parallelism the sufficient statistics are all-reduced, so every rank returns the value of the GLOBAL validation set and
        takes the same early-stopping / checkpoint decisions.  One D2H read at the end."""
        eng = self.engine
        acc = None
        for x, y in data:
            if not self._on_device(x):
                x, y = eng.stage_host_batch(x, y)
                s = self._batch_metrics(eng.forward(x), y)
                eng.release_staged()
            else:
                s = self._batch_metrics(eng.forward(x), y)
            acc = s.clone() if acc is None else acc + s
        if acc is None:
            acc = torch.zeros(5, dtype=torch.float64, device=self.device)
        if self.world > 1:
            torch.distributed.all_reduce(acc, group=self.pg)
        se, ae, hits, n, rows = (float(v) for v in acc.tolist())
        out = {"loss": se / n if n else float("nan"), "mse": se / n if n else float("nan"), "mae": ae / n if n else float("nan"),
               "accuracy": hits / rows if rows else float("nan")}
        return out if return_dict else out["loss"]

    def fit(self, train, epochs: int, validation_data=None, checkpoint_best: Optional[str] = None, checkpoint_last: Optional[str] = None,
            csv_log: Optional[str] = None, early_stopping_patience: Optional[int] = None, initial_epoch: int = 0, verbose: int = 2,
            train_metrics: bool = True, lr_schedule: Optional[Callable[[int], float]] = None) -> dict:
        """``model.fit(tds, epochs=.., validation_data=tds_val, callbacks=[checkpoint_best, checkpoint_last, csv_logger, earlystop])`` as
        the reference's retraining script drives it (baseline_v1/step2_retrain/step2_retrain.py:252-286):

        * ``train`` / ``validation_data``: objects with ``.epoch(e)`` yielding ``(x, y)`` (``NpyColumnStream``: reshuffled every
          epoch like ``shuffle(reshuffle_each_iteration=True)``) or plain re-iterable collections of ``(x, y)``;
        * ``loss`` of an epoch = mean of its batch losses (what Keras prints), accumulated on the device: one D2H read per epoch;
          ``mse`` equals it for ``loss='mse'``; ``mae`` / ``accuracy`` of the training batches (``train_metrics``) come from the
          predictions of a forward pass on each batch BEFORE its update, which is what Keras' running metrics see;
        * ``checkpoint_best``: saved when ``val_loss`` improves (``ModelCheckpoint(monitor='val_loss', save_best_only=True)``);
          ``checkpoint_last``: saved every epoch; ``csv_log``: rows ``epoch,accuracy,loss,mae,mse,val_accuracy,val_loss,val_mae,val_mse``
          appended (``CSVLogger(append=True)`` writes ``epoch`` and then the log keys in sorted order);
          ``early_stopping_patience``: stop after that many epochs without a new best ``val_loss`` (``EarlyStopping('val_loss', patience)``);
        * ``lr_schedule(epoch) -> lr``: Keras' ``LearningRateScheduler`` callback -- the learning rate is set at the start of every
          epoch (``ed_step_lr`` is the encoder-decoder's schedule, ClimSIM_ED_1_3_train.py:98-122);
        * data parallelism: the epoch loss and all validation statistics are all-reduced, so every rank takes the same decisions;
          files are written by rank 0 only.

        Returns Keras' ``History.history`` (``loss, mse, mae, accuracy`` and their ``val_`` twins) plus ``stopped_epoch``."""
        keys = list(self.METRICS) + ["val_" + k for k in self.METRICS]
        history = {k: [] for k in keys}
        history["stopped_epoch"] = None
        csv_keys = sorted(keys)
        best, wait = float("inf"), 0
        writer = self.rank == 0
        if csv_log is not None and writer:
            import os
            if not os.path.exists(csv_log) or os.path.getsize(csv_log) == 0:
                with open(csv_log, "a") as f:
                    f.write("epoch," + ",".join(csv_keys) + "\n")
        for epoch in range(initial_epoch, epochs):
            if lr_schedule is not None:
                self.lr = float(lr_schedule(epoch))
            batches = train.epoch(epoch) if hasattr(train, "epoch") else train
            total, nb, stats = None, 0, None
            for x, y in batches:
                if train_metrics:
                    xs, ys, staged = x, y, False
                    if not self._on_device(x):
                        # the metrics forward and the step share ONE staged copy of the batch
                        xs, ys = self.engine.stage_host_batch(x, y)
                        staged = True
                    s = self._batch_metrics(self.engine.forward(xs), ys)
                    stats = s.clone() if stats is None else stats + s
                    l = self._step_local(xs, ys)
                    if staged:
                        self.engine.release_staged()
                else:
                    l = self._step_local(x, y)
                total = l.detach().clone().reshape(()) if total is None else total + l.detach().reshape(())
                nb += 1
            if total is None:
                total = torch.zeros((), dtype=torch.float32, device=self.device)
            if self.world > 1:
                if not self._peer:                                   # every rank holds its share of the global-mean loss
                    torch.distributed.all_reduce(total, group=self.pg)
                if stats is not None:
                    torch.distributed.all_reduce(stats, group=self.pg)
            loss = float(total.item()) / nb if nb else float("nan")
            row = {"loss": loss, "mse": loss, "mae": float("nan"), "accuracy": float("nan")}
            if stats is not None:
                se, ae, hits, n, rows = (float(v) for v in stats.tolist())
                row["mae"], row["accuracy"] = ae / n, hits / rows
            if validation_data is not None:
                vb = validation_data.epoch(epoch) if hasattr(validation_data, "epoch") else validation_data
                for k, v in self.evaluate(vb, return_dict=True).items():
                    row["val_" + k] = v
            for k in keys:
                if k in row:
                    history[k].append(row[k])
            val = row.get("val_loss")
            if verbose and writer:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {row[k]:.6g}" for k in keys if k in row), flush=True)
            if csv_log is not None and writer:
                with open(csv_log, "a") as f:
                    f.write(f"{epoch}," + ",".join(repr(row[k]) if k in row else "" for k in csv_keys) + "\n")
            if checkpoint_last is not None and writer:
                self.save_checkpoint(checkpoint_last)
            if val is not None:
                if val < best:
                    best, wait = val, 0
                    if checkpoint_best is not None and writer:
                        self.save_checkpoint(checkpoint_best)
                else:
                    wait += 1
                    if early_stopping_patience is not None and wait >= early_stopping_patience:
                        history["stopped_epoch"] = epoch
                        break
        if self.world > 1:
            torch.distributed.barrier(group=self.pg)                 # rank 0's files are complete when any rank returns
        return history
