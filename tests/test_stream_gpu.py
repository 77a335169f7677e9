"""NpyColumnStream on a B200: .npy files -> pinned windows -> HBM -> csb_gather_rows batches, against numpy fancy indexing with the
plan's own row order (the oracle of a gather is the gather)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _files(tmp_path, n, dtype=np.float32):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(n, 124)).astype(dtype)
    y = rng.normal(size=(n, 128)).astype(dtype)
    np.save(tmp_path / "train_input.npy", x)
    np.save(tmp_path / "train_target.npy", y)
    return x.astype(np.float32), y.astype(np.float32)


@pytest.mark.parametrize("n,batch,window,world,dtype", [(5000, 512, 2048, 1, np.float32), (5000, 512, 2048, 2, np.float32),
                                                       (3000, 256, 100000, 1, np.float64), (777, 1024, 1024, 1, np.float32)])
def test_stream_matches_plan_and_covers_the_epoch(tmp_path, n, batch, window, world, dtype):
    from climsim_b200 import NpyColumnStream
    x, y = _files(tmp_path, n, dtype)
    seen = np.zeros(n, np.int64)
    for rank in range(world):
        st = NpyColumnStream(str(tmp_path / "train_input.npy"), str(tmp_path / "train_target.npy"), batch, window=window, seed=9,
                             rank=rank, world=world)
        for epoch in (0, 1):
            want = list(st.plan.epoch_rows(epoch))
            got = 0
            for (bx, by), rows in zip(st.epoch(epoch), want):
                assert bx.shape == (len(rows), 124) and by.shape == (len(rows), 128) and bx.is_cuda
                np.testing.assert_array_equal(bx.cpu().numpy(), x[rows])        # bit-exact: a gather moves bits
                np.testing.assert_array_equal(by.cpu().numpy(), y[rows])        # inputs and targets stay paired
                if epoch == 0:
                    seen[rows] += 1
                got += 1
            assert got == len(want) == len(st)
    assert (seen == 1).all()


@pytest.mark.parametrize("n,batch,world", [(5000, 512, 1), (5003, 512, 2), (777, 1024, 1)])
def test_resident_stream_full_shuffle_matches_plan(tmp_path, n, batch, world):
    """ResidentColumnStream: the rank's share uploaded once (in two chunks here), every epoch a fresh permutation of ALL its rows cut
    by csb_gather_rows; order = StreamPlan with one window spanning the share; ranks get equal batch counts (padded shares)."""
    from climsim_b200 import ResidentColumnStream
    x, y = _files(tmp_path, n)
    seen = np.zeros(n, np.int64)
    counts = []
    for rank in range(world):
        st = ResidentColumnStream(str(tmp_path / "train_input.npy"), str(tmp_path / "train_target.npy"), batch, seed=4, rank=rank, world=world,
                                  chunk_rows=max(1, -(-n // world) // 2 + 1))
        assert len(st.plan.windows()) == 1
        orders = []
        for epoch in (0, 1):
            want = list(st.plan.epoch_rows(epoch))
            got = 0
            for (bx, by), rows in zip(st.epoch(epoch), want):
                np.testing.assert_array_equal(bx.cpu().numpy(), x[rows])
                np.testing.assert_array_equal(by.cpu().numpy(), y[rows])
                if epoch == 0:
                    seen[rows] += 1
                got += 1
            assert got == len(want) == len(st)
            orders.append(np.concatenate(want))
        assert (orders[0] != orders[1]).any()                            # reshuffled every epoch, over the whole share
        counts.append(len(st))
    assert len(set(counts)) == 1 and (seen >= 1).all() and int((seen - 1).sum()) == -(-n // world) * world - n


def test_gather_rows_reports_bad_indices():
    from climsim_b200 import _lib
    lib = _lib.load()
    src = torch.arange(40, dtype=torch.float32, device="cuda").view(10, 4)
    dst = torch.full((3, 4), -1.0, device="cuda")
    idx = torch.tensor([2, 10, 0], dtype=torch.int64, device="cuda")           # 10 is out of range
    _lib.check(lib.csb_gather_rows(src.data_ptr(), idx.data_ptr(), dst.data_ptr(), 3, 4, 10, None), "csb_gather_rows")
    with pytest.raises(_lib.CsbError):
        _lib.check(lib.csb_gather_rows_check(None), "csb_gather_rows_check")
    assert (dst[0] == src[2]).all() and (dst[2] == src[0]).all() and (dst[1] == -1).all()
    _lib.check(lib.csb_gather_rows_check(None), "csb_gather_rows_check")       # the flag was cleared
    # unaligned / odd row length takes the scalar path
    src3 = torch.arange(30, dtype=torch.float32, device="cuda").view(10, 3)
    dst3 = torch.empty(2, 3, device="cuda")
    idx3 = torch.tensor([9, 1], dtype=torch.int64, device="cuda")
    _lib.check(lib.csb_gather_rows(src3.data_ptr(), idx3.data_ptr(), dst3.data_ptr(), 2, 3, 10, None), "csb_gather_rows")
    torch.cuda.synchronize()
    assert torch.equal(dst3, src3[idx3])


def test_end_to_end_stream_train_predict(tmp_path):
    """examples/train_mlp_v1.py: .npy files -> NpyColumnStream -> Trainer.step (MLP_v1, Keras Adam, cyclical LR) -> slab prediction;
    three epochs on a learnable synthetic target cut the validation MSE by more than half."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("train_mlp_v1", os.path.join(os.path.dirname(__file__), "..", "examples", "train_mlp_v1.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.run(columns=60_000, batch=2048, epochs=3, workdir=str(tmp_path), verbose=False)
    assert res["epoch_losses"][-1] < res["epoch_losses"][0]
    assert res["val_mse_after"] < 0.5 * res["val_mse_before"], res


def test_fit_with_streams(tmp_path):
    """Trainer.fit (model.fit with the reference's callbacks) fed by NpyColumnStream for both the training and the validation split
    (scripts/fit_check.py): losses fall, 9 steps per epoch were taken, the CSV log has a header and one row per epoch."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fit_check", os.path.join(os.path.dirname(__file__), "..", "scripts", "fit_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    h = mod.run(str(tmp_path), verbose=0)
    assert h["loss"][-1] < h["loss"][0] and h["val_loss"][-1] < h["val_loss"][0] and h["stopped_epoch"] is None
    assert h["iteration"] == 3 * (20000 // 2048) and h["log_rows"] == 4
