"""Host logic of the column stream (climsim_b200/stream.py::StreamPlan): windows, per-window permutations, batches, rank shares.
No GPU: the device side (csb_gather_rows through NpyColumnStream) is covered by tests/test_stream_gpu.py."""
import numpy as np
import pytest

from climsim_b200.stream import StreamPlan


@pytest.mark.parametrize("n,batch,window,world", [(5000, 512, 2048, 1), (5000, 512, 2048, 3), (384 * 7 + 5, 3072, 384 * 30, 2),
                                                  (100, 128, 64, 1), (1, 4, 4, 1)])
def test_every_row_once_per_epoch(n, batch, window, world):
    seen = np.zeros(n, np.int64)
    for rank in range(world):
        plan = StreamPlan(n, batch, window, shuffle=True, seed=3, rank=rank, world=world)
        assert plan.window % batch == 0 and plan.window >= batch
        nb = 0
        for rows in plan.epoch_rows(epoch=1):
            assert 0 < len(rows) <= batch and rows.min() >= plan.lo and rows.max() < plan.hi
            seen[rows] += 1
            nb += 1
        assert nb == plan.batches_per_epoch()
    # no row dropped; ranks pad to ceil(n / world) rows each (DistributedSampler-style), so at most world - 1 rows repeat
    assert (seen >= 1).all() and int((seen - 1).sum()) == -(-n // world) * world - n and seen.max() <= 2


@pytest.mark.parametrize("n,batch,world,drop_last", [(2049, 1024, 2, False), (4095, 1024, 2, True), (4095, 1024, 2, False),
                                                     (10007, 512, 8, False), (10007, 512, 8, True), (384 * 30 * 8 * 3 + 17, 3072, 4, False)])
def test_every_rank_gets_the_same_number_of_batches(n, batch, world, drop_last):
    """One gradient all-reduce per batch: a rank with an extra batch would hang.  Shares are ceil(n / world) rows each."""
    plans = [StreamPlan(n, batch, 4 * batch, rank=r, world=world, drop_last=drop_last) for r in range(world)]
    assert len({p.batches_per_epoch() for p in plans}) == 1 and len({p.rows for p in plans}) == 1
    sizes = [[len(rows) for rows in p.epoch_rows(0)] for p in plans]
    assert all(sorted(s) == sorted(sizes[0]) for s in sizes)            # same batch sizes too: the gradient scale uses the local B
    assert plans[0].lo == 0 and plans[-1].hi == n


def test_drop_last_and_sizes():
    plan = StreamPlan(5000, 512, 2048, drop_last=True)
    sizes = [len(r) for r in plan.epoch_rows(0)]
    assert set(sizes) == {512} and len(sizes) == plan.batches_per_epoch() == 4 + 4 + 1     # windows of 2048, 2048, 904 rows
    plan = StreamPlan(5000, 512, 2048, drop_last=False)
    assert sorted(len(r) for r in plan.epoch_rows(0))[:1] == [904 - 512]


def test_order_is_reproducible_and_reshuffled():
    a = np.concatenate(list(StreamPlan(4000, 256, 1024, seed=7).epoch_rows(2)))
    b = np.concatenate(list(StreamPlan(4000, 256, 1024, seed=7).epoch_rows(2)))
    c = np.concatenate(list(StreamPlan(4000, 256, 1024, seed=7).epoch_rows(3)))
    d = np.concatenate(list(StreamPlan(4000, 256, 1024, seed=8).epoch_rows(2)))
    assert (a == b).all() and (a != c).any() and (a != d).any()
    assert (np.concatenate(list(StreamPlan(4000, 256, 1024, shuffle=False).epoch_rows(5))) == np.arange(4000)).all()


def test_rows_mix_within_a_window_only():
    """The device holds one window at a time: a batch never mixes rows of two windows (the reference's shuffle buffer of 30 days has
    the same locality)."""
    plan = StreamPlan(10000, 500, 2000, seed=1)
    for rows in plan.epoch_rows(0):
        assert rows.min() // 2000 == rows.max() // 2000
