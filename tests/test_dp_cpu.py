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
