"""World-size-2 gloo test (CPU) of the data-parallel host logic in ``climsim_b200.trainer.Trainer``.

The engine here is a test double that implements the ``MLPEngine`` surface with the CPU oracle (the real engine needs
a GPU).  What is under test is the Trainer's contract: each rank steps on its shard with grad_scale = 1/(B_global*128),
the flat gradient buffers are summed with ONE all-reduce, every rank applies the same update -- and the result equals a
single-process step on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import models as M

UNITS = (64, 32)


class OracleEngine:
    """MLPEngine look-alike on the CPU oracle (test double)."""
    device = "cpu"
    in_dim, out_dim = 124, 128

    def __init__(self, seed=0):
        self.ref = M.MLPRef(units=UNITS, seed=seed)
        self.ref.randomize_biases(seed + 1)
        self.n = sum(p.numel() for p in self.ref.params)
        self._grad = torch.zeros(self.n)
        self.m = [torch.zeros_like(p) for p in self.ref.params]
        self.v = [torch.zeros_like(p) for p in self.ref.params]
        self.t = 0

    def grad_buffer(self):
        return self._grad

    def train_step(self, x, y, grad_scale=0.0, normalize_in=False, loss_out=None, fused_opt=False):
        assert not (fused_opt and dist.is_initialized() and dist.get_world_size() > 1)   # DP callers need the gradient buffer
        for p in self.ref.params:
            p.grad = None
        scale = grad_scale if grad_scale > 0 else 1.0 / (x.shape[0] * 128)
        loss = ((self.ref(x) - y) ** 2).sum() * scale
        loss.backward()
        self._grad.copy_(torch.cat([p.grad.reshape(-1) for p in self.ref.params]))
        loss_out.fill_(loss.item())
        return loss_out

    def apply_opt(self, rule, lr, beta1, beta2, eps, weight_decay):
        self.t += 1
        grads, off = [], 0
        for p in self.ref.params:
            grads.append(self._grad[off:off + p.numel()].view_as(p).clone()); off += p.numel()
        M.keras_adam_step(self.ref.params, grads, self.m, self.v, self.t, lr, beta1, beta2, eps if eps is not None else 1e-7)

    def flat(self):
        return torch.cat([p.detach().reshape(-1) for p in self.ref.params]).numpy()

    # -- the rest of the MLPEngine surface Trainer.fit / checkpoints use
    def forward(self, x):
        with torch.no_grad():
            return self.ref(x)

    def get_params_flat(self):
        return self.flat().copy()

    def set_params_flat(self, flat):
        off = 0
        with torch.no_grad():
            for p in self.ref.params:
                p.copy_(torch.from_numpy(np.asarray(flat[off:off + p.numel()])).view_as(p)); off += p.numel()

    def get_opt_state(self):
        cat = lambda ts: torch.cat([t.reshape(-1) for t in ts]).numpy().copy()
        return cat(self.m), cat(self.v), self.t

    def set_opt_state(self, m, v, step):
        off = 0
        for mi, vi in zip(self.m, self.v):
            n = mi.numel()
            mi.copy_(torch.from_numpy(np.asarray(m[off:off + n])).view_as(mi)); vi.copy_(torch.from_numpy(np.asarray(v[off:off + n])).view_as(vi)); off += n
        self.t = int(step)


def _batch(B):
    g = torch.Generator().manual_seed(3)
    return 0.2 * torch.randn(B, 124, generator=g), 0.1 * torch.randn(B, 128, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from climsim_b200.trainer import Trainer
    torch.set_num_threads(1)
    eng = OracleEngine()
    tr = Trainer(eng, rule="adam_keras", lr=1e-3)
    x, y = _batch(64)
    shard = slice(rank * 32, (rank + 1) * 32)
    losses = [tr.step(x[shard], y[shard]) for _ in range(3)]
    if rank == 0:
        np.savez(out, flat=eng.flat(), losses=np.array(losses))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "dp.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    # single process, whole batch
    from climsim_b200.trainer import Trainer
    eng = OracleEngine()
    tr = Trainer(eng, rule="adam_keras", lr=1e-3)
    x, y = _batch(64)
    losses = [tr.step(x, y) for _ in range(3)]
    np.testing.assert_allclose(got["losses"], losses, rtol=1e-5)
    np.testing.assert_allclose(got["flat"], eng.flat(), rtol=0, atol=2e-6)


def test_cyclical_lr_matches_oracle_closed_form():
    from climsim_b200.trainer import cyclical_lr
    for step in (0, 1, 7, 10, 11, 25, 40, 41, 99):
        assert cyclical_lr(step, 2.5e-4, 2.5e-3, 10) == pytest.approx(M.cyclical_lr(step, 2.5e-4, 2.5e-3, 10), rel=1e-12)


def test_glorot_flat_blob_shape_and_limits():
    from climsim_b200.trainer import glorot_uniform_flat
    dims = [(124, 768), (768, 640), (640, 128), (128, 128)]
    flat = glorot_uniform_flat(dims, seed=0)
    assert flat.size == sum(k * n + n for k, n in dims)
    w0 = flat[:124 * 768]
    assert np.abs(w0).max() <= np.sqrt(6.0 / (124 + 768)) and flat[124 * 768:124 * 768 + 768].max() == 0


def _stream_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from climsim_b200.stream import StreamPlan
    from climsim_b200.trainer import Trainer
    torch.set_num_threads(1)
    x, y = _batch(96)
    plan = StreamPlan(96, batch_size=16, window=32, seed=5, rank=rank, world=world)
    eng = OracleEngine()
    tr = Trainer(eng, rule="adam_keras", lr=1e-3)
    losses = [tr.step(x[rows], y[rows]) for rows in plan.epoch_rows(epoch=0)]
    if rank == 0:
        np.savez(out, flat=eng.flat(), losses=np.array(losses))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_shares_train_like_one_process_on_the_joined_batches(tmp_path):
    """StreamPlan's per-rank shares feeding the data-parallel Trainer: rank r steps on the batches of its own share; the result equals
    one process stepping on the concatenation of the two ranks' k-th batches (the global batch of step k)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "dp_stream.npz")
    mp.spawn(_stream_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    from climsim_b200.stream import StreamPlan
    from climsim_b200.trainer import Trainer
    x, y = _batch(96)
    plans = [StreamPlan(96, batch_size=16, window=32, seed=5, rank=r, world=2) for r in range(2)]
    assert plans[0].hi == plans[1].lo == 48 and plans[0].batches_per_epoch() == plans[1].batches_per_epoch() == 3
    eng = OracleEngine()
    tr = Trainer(eng, rule="adam_keras", lr=1e-3)
    losses = []
    for r0, r1 in zip(plans[0].epoch_rows(0), plans[1].epoch_rows(0)):
        rows = np.concatenate([r0, r1])
        losses.append(tr.step(x[rows], y[rows]))
    np.testing.assert_allclose(got["losses"], losses, rtol=1e-5)
    np.testing.assert_allclose(got["flat"], eng.flat(), rtol=0, atol=2e-6)


def test_fit_driver_callbacks(tmp_path):
    """Trainer.fit = model.fit with the reference's callbacks (step2_retrain.py:252-286): per-epoch loss / val_loss history, CSV log
    appended per epoch, last / best checkpoints that restore parameters + optimizer state + counters, early stopping on val_loss."""
    from climsim_b200.trainer import Trainer
    x, y = _batch(96)
    xv, yv = 0.2 * torch.randn(40, 124, generator=torch.Generator().manual_seed(9)), 0.1 * torch.randn(40, 128, generator=torch.Generator().manual_seed(10))
    train = [(x[i:i + 32], y[i:i + 32]) for i in range(0, 96, 32)]
    val = [(xv[:24], yv[:24]), (xv[24:], yv[24:])]                       # unequal batches: val_loss is the exact mean over elements
    eng = OracleEngine()
    tr = Trainer(eng, rule="adam_keras", lr=1e-3)
    best, last, log = str(tmp_path / "best.ckpt"), str(tmp_path / "last.ckpt"), str(tmp_path / "metrics.csv")
    h = tr.fit(train, epochs=4, validation_data=val, checkpoint_best=best, checkpoint_last=last, csv_log=log, verbose=0)
    assert len(h["loss"]) == len(h["val_loss"]) == 4 and h["stopped_epoch"] is None
    assert h["loss"][-1] < h["loss"][0] and tr.iteration == 12
    want_val = float(((eng.forward(xv) - yv) ** 2).mean())
    assert h["val_loss"][-1] == pytest.approx(want_val, rel=1e-6)
    # Keras' metrics=['mse','mae','accuracy'] (hpo_baseline_v1.py:127-129): exact means over the validation elements
    pv = eng.forward(xv)
    assert h["val_mse"][-1] == h["val_loss"][-1] and h["val_mae"][-1] == pytest.approx(float((pv - yv).abs().mean()), rel=1e-6)
    assert h["val_accuracy"][-1] == pytest.approx(float((pv.argmax(1) == yv.argmax(1)).float().mean()), abs=1e-9)
    assert h["mse"] == h["loss"] and all(0.0 < m < 1.0 for m in h["mae"]) and all(0.0 <= a <= 1.0 for a in h["accuracy"])
    rows = open(log).read().strip().splitlines()
    # CSVLogger: 'epoch' first, then the log keys in sorted order (step2_retrain.py:262)
    assert rows[0] == "epoch,accuracy,loss,mae,mse,val_accuracy,val_loss,val_mae,val_mse" and len(rows) == 5 and rows[1].startswith("0,")
    assert os.path.exists(best) and os.path.exists(last) and not os.path.exists(best + ".npz")
    # the last checkpoint restores everything: a restored trainer continues exactly like the original
    eng2 = OracleEngine(seed=5)
    tr2 = Trainer(eng2, rule="adam_keras", lr=1e-3)
    tr2.load_checkpoint(last)
    assert tr2.iteration == 12 and eng2.t == eng.t
    np.testing.assert_array_equal(eng2.flat(), eng.flat())
    a, b = tr.step(*train[0]), tr2.step(*train[0])
    assert a == pytest.approx(b, rel=1e-6)
    np.testing.assert_allclose(eng2.flat(), eng.flat(), rtol=0, atol=1e-7)
    # appended log, early stopping: with lr = 0 the validation loss never improves after the first epoch
    tr3 = Trainer(OracleEngine(), rule="adam_keras", lr=0.0)
    h3 = tr3.fit(train, epochs=20, validation_data=val, csv_log=log, early_stopping_patience=3, verbose=0)
    assert h3["stopped_epoch"] == 3 and len(h3["val_loss"]) == 4
    assert len(open(log).read().strip().splitlines()) == 5 + 4


def _fit_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from climsim_b200.stream import StreamPlan
    from climsim_b200.trainer import Trainer
    torch.set_num_threads(1)
    x, y = _batch(99)                                                   # 99 rows over 2 ranks: the shares are padded to 50 rows each
    xv, yv = _batch(41)
    plan = StreamPlan(99, batch_size=16, window=32, seed=5, rank=rank, world=world)
    vplan = StreamPlan(41, batch_size=16, window=16, shuffle=False, rank=rank, world=world)

    class Split:
        def __init__(self, p, a, b):
            self.p, self.a, self.b = p, a, b

        def epoch(self, e):
            return ((self.a[r], self.b[r]) for r in self.p.epoch_rows(e))

    eng = OracleEngine(seed=rank)                                       # different initial parameters per rank: Trainer must broadcast rank 0's
    tr = Trainer(eng, rule="adam_keras", lr=0.0 if os.environ.get("DP_FIT_LR0") else 1e-3)
    log, ck = out + ".csv", out + ".ckpt"
    h = tr.fit(Split(plan, x, y), epochs=6, validation_data=Split(vplan, xv, yv), csv_log=log, checkpoint_last=ck,
               early_stopping_patience=2, verbose=0)
    np.savez(out + f".{rank}.npz", flat=eng.flat(), val=np.array(h["val_loss"]), loss=np.array(h["loss"]),
             stopped=-1 if h["stopped_epoch"] is None else h["stopped_epoch"], it=tr.iteration)
    dist.destroy_process_group()


@pytest.mark.parametrize("lr0", [False, True])
def test_two_rank_fit_takes_the_same_decisions_on_every_rank(tmp_path, lr0, monkeypatch):
    """under data parallelism val_loss must be the GLOBAL value on every rank (else early stopping diverges and the next
    all-reduce deadlocks), the epoch loss is reduced exactly once, only rank 0 writes files, rank 0's parameters are broadcast, and
    ranks take equally many steps although 2 does not divide the 99 training rows."""
    if lr0:
        monkeypatch.setenv("DP_FIT_LR0", "1")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "fit")
    mp.spawn(_fit_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    np.testing.assert_array_equal(a["flat"], b["flat"])
    np.testing.assert_array_equal(a["val"], b["val"])
    np.testing.assert_array_equal(a["loss"], b["loss"])
    assert int(a["stopped"]) == int(b["stopped"]) and int(a["it"]) == int(b["it"])
    if lr0:
        assert int(a["stopped"]) == 2 and len(a["val"]) == 3          # never improves after epoch 0: patience 2 stops at epoch 2
    else:
        assert int(a["stopped"]) == -1 and a["loss"][-1] < a["loss"][0]
    rows = open(out + ".csv").read().strip().splitlines()
    assert len(rows) == 1 + len(a["val"])                               # one header + one row per epoch: a single writer
    # the global validation loss: 41 rows padded to 21 + 21 (one row seen twice), exact mean over what the ranks evaluated
    xv, yv = _batch(41)
    eng = OracleEngine(seed=0)
    if lr0:
        rows_seen = np.concatenate([np.arange(0, 21), np.arange(20, 41)])
        want = float(((eng.forward(xv[rows_seen]) - yv[rows_seen]) ** 2).mean())
        assert a["val"][0] == pytest.approx(want, rel=1e-6)


def test_weight_decay_is_rejected_for_rules_without_it():
    from climsim_b200.trainer import Trainer
    with pytest.raises(ValueError):
        Trainer(OracleEngine(), rule="adam_keras", weight_decay=1e-2)


def test_ed_learning_rate_schedule_and_fit_callback():
    """ClimSIM_ED_1_3_train.py:98-122: lr / 5 after every 7th epoch; Trainer.fit applies it like LearningRateScheduler."""
    from climsim_b200.trainer import Trainer, ed_step_lr
    want = {0: 1e-4, 6: 1e-4, 7: 1e-4 / 5, 13: 1e-4 / 5, 14: 1e-4 / 25, 21: 1e-4 / 125, 28: 1e-4 / 625, 35: 1e-4 / 3125, 41: 1e-4 / 3125}
    for e, lr in want.items():
        assert ed_step_lr(e) == pytest.approx(lr, rel=1e-12)
    seen = []

    class Spy(OracleEngine):
        def apply_opt(self, rule, lr, **kw):
            seen.append(lr)
            super().apply_opt(rule, lr, **kw)

    x, y = _batch(32)
    tr = Trainer(Spy(), rule="adam_keras", lr=123.0)
    tr.fit([(x, y)], epochs=15, lr_schedule=ed_step_lr, verbose=0, train_metrics=False)
    assert seen == [pytest.approx(ed_step_lr(e)) for e in range(15)]
