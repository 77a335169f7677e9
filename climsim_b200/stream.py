"""Column batches streamed from the reference's ``.npy`` files with the shuffle done on the device.

The reference feeds its trainers from a Python generator / ``tf.data`` chain -- ``unbatch().shuffle(384*30).batch(B).prefetch(..)``
(baseline_models/MLP/training/HPO/baseline_v1/hpo_baseline_v1.py:140-143, step2_retrain/step2_retrain.py:266-271) over the arrays
``data_utils.save_as_npy`` writes (``<split>_input.npy`` (N,124) / ``<split>_target.npy`` (N,128) fp32,
climsim_utils/data_utils.py:884-921) -- or from a per-row ``h5py`` ``__getitem__`` behind a ``DistributedSampler``
(online_testing/baseline_models/MLP_v2rh/training/climsim_datapip_h5.py:104-177, train_mlp_h5loader.py:126-134).  Once the step runs
at tens of millions of columns per second that host-side chain is the wall, so here:

* the two arrays are memory-mapped and read in contiguous *windows* of ``window`` columns (a multiple of the batch size; default
  30 "days" x 384 columns x 8, the reference's shuffle buffer is 30 days) by a background thread into pinned staging buffers,
* a window goes to HBM with one H2D copy per array on a copy stream (double-buffered: the copy of window i+1 overlaps the batches
  of window i),
* batches are cut from the window by ``csb_gather_rows`` with a per-window permutation (``numpy`` PCG64 keyed by
  (seed, epoch, window), so that the order is reproducible and testable without a GPU); the window order is shuffled per epoch as the
  reference shuffles its file list (step2_retrain.py:198).

Data parallelism: rank r of W takes the r-th contiguous share of ``ceil(N / W)`` rows.  Every rank gets the SAME number of rows --
hence the same windows and the same number of batches per epoch, which the per-batch gradient all-reduce needs (a rank with one batch
more would wait for its peers forever).  When W does not divide N the last shares start a few rows early, i.e. up to W - 1 rows are
visited twice per epoch: ``DistributedSampler``'s padding with repeated samples (train_mlp_h5loader.py:126-134), in contiguous form.

``StreamPlan`` is the pure host logic (no CUDA); ``NpyColumnStream`` executes it.
"""
from __future__ import annotations

import threading
from typing import Iterator, List, Optional, Tuple

import numpy as np

DEFAULT_WINDOW = 384 * 30 * 8


class StreamPlan:
    """Which rows form which batch: windows of the row range of this rank, a permutation per window, batches cut from it."""

    def __init__(self, n_rows: int, batch_size: int, window: int = DEFAULT_WINDOW, shuffle: bool = True, seed: int = 0,
                 rank: int = 0, world: int = 1, drop_last: bool = False):
        assert n_rows > 0 and batch_size > 0 and 0 <= rank < world
        self.n_rows, self.batch_size, self.shuffle, self.seed = int(n_rows), int(batch_size), bool(shuffle), int(seed)
        self.rank, self.world, self.drop_last = rank, world, bool(drop_last)
        self.window = max(self.batch_size, (int(window) // self.batch_size) * self.batch_size)
        # this rank's contiguous share [lo, hi): ceil(n_rows / world) rows on EVERY rank (equal batch counts, see the module docstring);
        # shares that would run past the end are shifted back, overlapping their predecessor by at most world - 1 rows in total
        per = -(-self.n_rows // world)
        assert per * (world - 1) < self.n_rows or world == 1, "more ranks than rows"
        self.lo = min(rank * per, self.n_rows - per)
        self.hi = self.lo + per

    @property
    def rows(self) -> int:
        return self.hi - self.lo

    def windows(self) -> List[Tuple[int, int]]:
        """(first row, length) of every window of this rank's share, in file order."""
        return [(s, min(self.window, self.hi - s)) for s in range(self.lo, self.hi, self.window)]

    def window_order(self, epoch: int) -> List[int]:
        n = len(self.windows())
        if not self.shuffle:
            return list(range(n))
        return [int(i) for i in np.random.Generator(np.random.PCG64([self.seed, epoch, 0xC11A])).permutation(n)]

    def permutation(self, epoch: int, window_index: int, length: int) -> np.ndarray:
        """Window-relative row order (int64) of window ``window_index`` in ``epoch``."""
        if not self.shuffle:
            return np.arange(length, dtype=np.int64)
        return np.random.Generator(np.random.PCG64([self.seed, epoch, window_index + 1])).permutation(length).astype(np.int64)

    def batches_in(self, length: int) -> List[Tuple[int, int]]:
        """(offset into the permutation, batch length) of the batches cut from a window of ``length`` rows."""
        out = [(o, min(self.batch_size, length - o)) for o in range(0, length, self.batch_size)]
        if self.drop_last and out and out[-1][1] < self.batch_size:
            out.pop()
        return out

    def batches_per_epoch(self) -> int:
        return sum(len(self.batches_in(ln)) for _, ln in self.windows())

    def epoch_rows(self, epoch: int) -> Iterator[np.ndarray]:
        """Absolute row indices of every batch of the epoch, in the order they are produced (what a CPU test checks)."""
        wins = self.windows()
        for wi in self.window_order(epoch):
            start, length = wins[wi]
            perm = self.permutation(epoch, wi, length)
            for off, n in self.batches_in(length):
                yield start + perm[off:off + n]


def open_column_array(path: str, dataset: str = "data"):
    """``<split>_input.npy`` / ``_target.npy`` (data_utils.py:884-921) or their ``.h5`` twins (data_utils.py:908-925: one dataset
    ``'data'``) as a read-only memory map of shape (N, F)."""
    if str(path).lower().endswith((".h5", ".hdf5")):
        from .keras_h5 import open_h5_dataset
        return open_h5_dataset(path, dataset)
    return np.load(path, mmap_mode="r")


class NpyColumnStream:
    """Iterate ``(x, y)`` CUDA tensors of ``batch_size`` columns over ``<split>_input.npy`` / ``<split>_target.npy``.

    ``for x, y in stream.epoch(e): trainer.step(x, y)`` -- the tensors are views of two rotating device buffers: they stay valid
    until the second next batch is produced (consume or copy them before that, as with any prefetching loader).
    """

    def __init__(self, input_path: str, target_path: str, batch_size: int, window: int = DEFAULT_WINDOW, shuffle: bool = True,
                 seed: int = 0, rank: int = 0, world: int = 1, drop_last: bool = False, device: str = "cuda"):
        import torch
        from . import _lib
        self.torch, self._lib, self.lib = torch, _lib, _lib.load()           # no CPU fallback: raises without the CUDA library
        self.x_all = open_column_array(input_path)
        self.y_all = open_column_array(target_path)
        assert self.x_all.ndim == 2 and self.y_all.ndim == 2 and self.x_all.shape[0] == self.y_all.shape[0], \
            "input and target arrays must be (N, F_in) and (N, F_out) with the same N"
        self.plan = StreamPlan(self.x_all.shape[0], batch_size, window, shuffle, seed, rank, world, drop_last)
        self.device = torch.device(device)
        self.f_in, self.f_out = int(self.x_all.shape[1]), int(self.y_all.shape[1])
        W = min(self.plan.window, self.plan.rows)
        pin = lambda f: torch.empty(W, f, dtype=torch.float32).pin_memory()
        dev = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=self.device)
        self._host = [(pin(self.f_in), pin(self.f_out)) for _ in range(2)]                 # staging, filled by the reader thread
        self._win = [(dev(W, self.f_in), dev(W, self.f_out), dev(W, dt=torch.int64)) for _ in range(2)]
        self._out = [(dev(batch_size, self.f_in), dev(batch_size, self.f_out)) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._copied = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]
        self._reader: Optional[threading.Thread] = None
        self._reader_error: Optional[BaseException] = None

    def __len__(self) -> int:
        return self.plan.batches_per_epoch()

    # -- host side: read window `wi` of this epoch into staging slot `slot` -----------------------------------------------------
    def _read(self, slot: int, start: int, length: int) -> None:
        try:
            hx, hy = self._host[slot]
            np.copyto(hx.numpy()[:length], self.x_all[start:start + length], casting="same_kind")
            np.copyto(hy.numpy()[:length], self.y_all[start:start + length], casting="same_kind")
        except BaseException as e:                                          # re-raised by the consumer: never train on a stale slot
            self._reader_error = e

    def _join_reader(self, t: threading.Thread) -> None:
        t.join()
        if self._reader_error is not None:
            e, self._reader_error = self._reader_error, None
            raise RuntimeError("NpyColumnStream: reading a window from the .npy files failed") from e

    def epoch(self, epoch: int = 0):
        torch, plan = self.torch, self.plan
        wins, order = plan.windows(), plan.window_order(epoch)
        if not order:
            return
        main = torch.cuda.current_stream(self.device)
        # the window buffers may still be read by gathers of a previous epoch() (or of one abandoned half-way): order after them
        self._copy_stream.wait_stream(main)
        if self._reader is not None:                                        # reader of an abandoned epoch: let it finish with its slot
            self._reader.join()
            self._reader_error = None
        reader: Optional[threading.Thread] = None

        def start_read(k: int) -> threading.Thread:
            s, ln = wins[order[k]]
            t = threading.Thread(target=self._read, args=(k & 1, s, ln), daemon=True)
            t.start()
            self._reader = t
            return t

        reader = start_read(0)
        used = [False, False]
        n_out = 0
        for k, wi in enumerate(order):
            slot = k & 1
            start, length = wins[wi]
            self._join_reader(reader)                                       # staging slot `slot` holds window k
            wx, wy, widx = self._win[slot]
            hx, hy = self._host[slot]
            perm = torch.from_numpy(plan.permutation(epoch, wi, length))
            with torch.cuda.stream(self._copy_stream):
                if used[slot]:
                    self._copy_stream.wait_event(self._consumed[slot])      # the batches of window k-2 have been cut from this buffer
                wx[:length].copy_(hx[:length], non_blocking=True)
                wy[:length].copy_(hy[:length], non_blocking=True)
                widx[:length].copy_(perm, non_blocking=False)               # small; pageable source, so a blocking copy
                self._copied[slot].record(self._copy_stream)
            self._copy_stream.synchronize()                                 # staging slot free for the reader (window k+2 reuses it)
            if k + 1 < len(order):
                reader = start_read(k + 1)                                  # disk / page-cache read of the next window overlaps the batches
            main.wait_event(self._copied[slot])
            for off, n in plan.batches_in(length):
                ox, oy = self._out[n_out & 1]
                n_out += 1
                sp = self._lib.current_stream_ptr()
                self._lib.check(self.lib.csb_gather_rows(wx.data_ptr(), widx[off:].data_ptr(), ox.data_ptr(), n, self.f_in, length, sp), "csb_gather_rows")
                self._lib.check(self.lib.csb_gather_rows(wy.data_ptr(), widx[off:].data_ptr(), oy.data_ptr(), n, self.f_out, length, sp), "csb_gather_rows")
                yield ox[:n], oy[:n]
            self._consumed[slot].record(main)
            used[slot] = True
        self._lib.check(self.lib.csb_gather_rows_check(self._lib.current_stream_ptr()), "csb_gather_rows_check")

    def __iter__(self):
        return self.epoch(0)


class ResidentColumnStream:
    """The same iterator with this rank's whole share of the split RESIDENT in HBM.

    The low-resolution training split is about 10 M columns x 1008 B = 10 GB -- a B200 holds it 17 times over -- and the reference
    walks the same arrays for 18 epochs (step2_retrain.py:280-285), so the share is uploaded ONCE (in chunks through one pinned
    staging buffer) and every epoch is a fresh permutation of ALL its rows (a full shuffle, where the streaming variant and the
    reference's 30-day buffer shuffle locally), cut into batches by ``csb_gather_rows``.  Per epoch the host sends 8 bytes per row (the
    permutation) instead of 1008, which takes the host link out of the training loop.  ``StreamPlan`` with one window spanning the
    share is the host logic (``epoch_rows`` gives the order without a GPU)."""

    def __init__(self, input_path: str, target_path: str, batch_size: int, shuffle: bool = True, seed: int = 0, rank: int = 0,
                 world: int = 1, drop_last: bool = False, device: str = "cuda", chunk_rows: int = 1 << 20):
        import torch
        from . import _lib
        self.torch, self._lib, self.lib = torch, _lib, _lib.load()
        x_all = open_column_array(input_path)
        y_all = open_column_array(target_path)
        assert x_all.ndim == 2 and y_all.ndim == 2 and x_all.shape[0] == y_all.shape[0], \
            "input and target arrays must be (N, F_in) and (N, F_out) with the same N"
        probe = StreamPlan(x_all.shape[0], batch_size, batch_size, shuffle, seed, rank, world, drop_last)
        one_window = -(-probe.rows // batch_size) * batch_size                 # a single window covering the whole share
        self.plan = StreamPlan(x_all.shape[0], batch_size, one_window, shuffle, seed, rank, world, drop_last)
        assert len(self.plan.windows()) == 1
        self.device = torch.device(device)
        self.f_in, self.f_out = int(x_all.shape[1]), int(y_all.shape[1])
        n = self.plan.rows
        self.x = torch.empty(n, self.f_in, dtype=torch.float32, device=self.device)
        self.y = torch.empty(n, self.f_out, dtype=torch.float32, device=self.device)
        chunk = min(chunk_rows, n)
        hx = torch.empty(chunk, self.f_in, dtype=torch.float32).pin_memory()
        hy = torch.empty(chunk, self.f_out, dtype=torch.float32).pin_memory()
        for s in range(0, n, chunk):
            ln = min(chunk, n - s)
            np.copyto(hx.numpy()[:ln], x_all[self.plan.lo + s:self.plan.lo + s + ln], casting="same_kind")
            np.copyto(hy.numpy()[:ln], y_all[self.plan.lo + s:self.plan.lo + s + ln], casting="same_kind")
            self.x[s:s + ln].copy_(hx[:ln], non_blocking=True)
            self.y[s:s + ln].copy_(hy[:ln], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()              # the staging buffer is reused by the next chunk
        self._idx = torch.empty(n, dtype=torch.int64, device=self.device)
        self._out = [(torch.empty(batch_size, self.f_in, dtype=torch.float32, device=self.device),
                      torch.empty(batch_size, self.f_out, dtype=torch.float32, device=self.device)) for _ in range(2)]

    def __len__(self) -> int:
        return self.plan.batches_per_epoch()

    def epoch(self, epoch: int = 0):
        torch, plan = self.torch, self.plan
        n = plan.rows
        self._idx.copy_(torch.from_numpy(plan.permutation(epoch, 0, n)))      # stream-ordered after the previous epoch's gathers
        n_out = 0
        for off, ln in plan.batches_in(n):
            ox, oy = self._out[n_out & 1]
            n_out += 1
            sp = self._lib.current_stream_ptr()
            self._lib.check(self.lib.csb_gather_rows(self.x.data_ptr(), self._idx[off:].data_ptr(), ox.data_ptr(), ln, self.f_in, n, sp), "csb_gather_rows")
            self._lib.check(self.lib.csb_gather_rows(self.y.data_ptr(), self._idx[off:].data_ptr(), oy.data_ptr(), ln, self.f_out, n, sp), "csb_gather_rows")
            yield ox[:ln], oy[:ln]
        self._lib.check(self.lib.csb_gather_rows_check(self._lib.current_stream_ptr()), "csb_gather_rows_check")

    def __iter__(self):
        return self.epoch(0)
