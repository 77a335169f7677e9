"""climsim_b200.keras_h5: the reference's Keras .h5 checkpoints read without h5py / TensorFlow.  The reader is exercised on the
reference's own shipped files (this container has them under /root/reference; elsewhere these tests skip), and the committed
fixture derived from one of them is tied back to the file."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import models as M

REF = os.environ.get("CLIMSIM_REFERENCE", "/root/reference")
MLP_H5 = os.path.join(REF, "baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_trial_0027.best.h5")
ED_H5 = os.path.join(REF, "baseline_models/ED/model/ED_ClimSIM_1_3_model.h5")
needs_ref = pytest.mark.skipif(not (os.path.exists(MLP_H5) and os.path.exists(ED_H5)), reason="reference checkout with the shipped Keras models not present")


@needs_ref
def test_shipped_mlp_v1_checkpoint(golden_dir):
    from climsim_b200.keras_h5 import read_keras_h5
    ck = read_keras_h5(MLP_H5)
    ref = M.MLPRef()
    # get_weights() order and layouts: kernel (in, out), bias; exactly the oracle's parameter list -- and step1_results.csv's count
    assert [tuple(w.shape) for w in ck["weights"]] == [tuple(p.shape) for p in ref.params]
    assert sum(w.size for w in ck["weights"]) == 1753472
    assert all(w.dtype == np.float32 and np.isfinite(w).all() for w in ck["weights"])
    # the configs the committed fixture was generated from are the ones in the file
    cfg = json.load(open(os.path.join(golden_dir, "keras_configs.json")))["mlp_v1"]
    assert [l["config"]["name"] for l in ck["model_config"]["config"]["layers"]] == [l["name"] for l in cfg["layers"]]
    assert ck["training_config"]["optimizer_config"] == cfg["optimizer"]
    # optimizer: RectifiedAdam slots for every weight, in the same order, second moments non-negative
    opt = ck["optimizer"]
    assert opt["name"] == "RectifiedAdam" and opt["iterations"] == 6570 and len(opt["m"]) == len(opt["v"]) == 16
    assert all(m.shape == w.shape == v.shape and (v >= 0).all() for m, v, w in zip(opt["m"], opt["v"], ck["weights"]))
    # the oracle runs the reference's trained model: a trained emulator of normalised targets predicts O(1) numbers, and ReLU on its 8 scalars
    with torch.no_grad():
        for p, w in zip(ref.params, ck["weights"]):
            p.copy_(torch.from_numpy(w))
        g = torch.Generator().manual_seed(0)
        out = ref(torch.rand(64, 124, generator=g))
    assert torch.isfinite(out).all() and out.abs().max() < 1e3 and (out[:, 120:] >= 0).all()
    # the GPU fixture is this file rounded to float16
    fx = np.load(os.path.join(golden_dir, "mlp_v1_shipped_f16.npz"))
    for i, w in enumerate(ck["weights"]):
        np.testing.assert_array_equal(fx[f"w{i:02d}"], w.astype(np.float16))
        np.testing.assert_allclose(fx["fingerprint"][i], [w.astype(np.float64).sum(), np.abs(w.astype(np.float64)).sum()], rtol=1e-12)
    assert int(fx["iterations"]) == opt["iterations"]


@needs_ref
def test_shipped_ed_checkpoint():
    from climsim_b200.keras_h5 import read_keras_h5
    ck = read_keras_h5(ED_H5)
    ref = M.EDRef()
    assert [tuple(w.shape) for w in ck["weights"]] == [tuple(p.shape) for p in ref.params]      # nested encoder / decoder flattened in model order
    assert sum(w.size for w in ck["weights"]) == ref.num_parameters() == 831879
    assert ck["optimizer"]["name"] == "Adam" and len(ck["optimizer"]["m"]) == 28 and ck["optimizer"]["iterations"] > 0
    with torch.no_grad():
        for p, w in zip(ref.params, ck["weights"]):
            p.copy_(torch.from_numpy(w))
        out = ref(torch.rand(32, 124, generator=torch.Generator().manual_seed(1)))
    assert torch.isfinite(out).all() and (out > -1.0).all()                                     # ELU output layer


def test_reader_rejects_what_it_does_not_understand(tmp_path):
    from climsim_b200.keras_h5 import read_keras_h5
    p = tmp_path / "not_hdf5.h5"
    p.write_bytes(b"PK\x03\x04 a zip, not HDF5")
    with pytest.raises(ValueError, match="not an HDF5 file"):
        read_keras_h5(str(p))
    q = tmp_path / "v2.h5"
    q.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + bytes(64))
    with pytest.raises(ValueError, match="version-0 superblocks"):
        read_keras_h5(str(q))


def test_fixture_is_self_consistent(golden_dir):
    """Runs everywhere: the committed float16 copy of the shipped MLP_v1 has the oracle's shapes and 1 753 472 parameters."""
    fx = np.load(os.path.join(golden_dir, "mlp_v1_shipped_f16.npz"))
    ws = [fx[f"w{i:02d}"] for i in range(16)]
    assert [tuple(w.shape) for w in ws] == [tuple(p.shape) for p in M.MLPRef().params] and sum(w.size for w in ws) == 1753472
    assert all(w.dtype == np.float16 for w in ws) and all(np.abs(b).max() > 0 for b in ws[1::2])      # trained biases are not Keras' zeros


@needs_ref
def test_trainer_load_keras_h5_hands_the_engine_parameters_and_optimizer_state():
    """Trainer.load_keras_h5 on a recording stand-in for the engine: the flat blobs it passes are the file's arrays in the engine's
    layout (the two output Dense layers fused column-wise, for the slots as for the weights), and the iteration count carries over."""
    from climsim_b200 import MLPEngine
    from climsim_b200.keras_h5 import read_keras_h5
    from climsim_b200.trainer import Trainer

    class Recorder:
        keras_to_flat = staticmethod(MLPEngine.keras_to_flat)
        dtype = "fp32"

        def __init__(self):
            self.calls = {}

        def set_params_flat(self, flat):
            self.calls["params"] = flat

        def set_opt_state(self, m, v, step):
            self.calls["opt"] = (m, v, step)

    tr = Trainer.__new__(Trainer)                      # no device, no process group: only the method under test
    tr.engine, tr.iteration = Recorder(), 0
    ck = tr.load_keras_h5(MLP_H5)
    want = read_keras_h5(MLP_H5)
    flat = tr.engine.calls["params"]
    assert flat.size == 1753472
    head_w = np.concatenate([want["weights"][-4], want["weights"][-2]], axis=1)              # [W_lin | W_relu], (128, 128)
    n_head = head_w.size + 128
    np.testing.assert_array_equal(flat[-n_head:-128].reshape(128, 128), head_w)
    np.testing.assert_array_equal(flat[-128:], np.concatenate([want["weights"][-3], want["weights"][-1]]))
    np.testing.assert_array_equal(flat[:124 * 768].reshape(124, 768), want["weights"][0])
    m, v, step = tr.engine.calls["opt"]
    assert step == tr.iteration == 6570 and m.size == v.size == flat.size and (v >= 0).all()
    np.testing.assert_array_equal(m[:124 * 768].reshape(124, 768), want["optimizer"]["m"][0])
    assert ck["optimizer"]["name"] == "RectifiedAdam"


@needs_ref
def test_h5_datasets_open_as_memory_maps_like_their_npy_twins(tmp_path):
    """The reference also writes its column arrays as HDF5 (climsim_utils/data_utils.py:908-925: one contiguous dataset 'data' per
    file) and reads them row by row with h5py (climsim_datapip_h5.py:104-126).  ``open_h5_dataset`` maps such a dataset read-only at
    its file offset -- exercised here on datasets h5py itself wrote (the shipped Keras file: 2-D fp32 arrays inside a group, the
    same object-header / dataspace / contiguous-layout structure as a root-level 'data'), row slices and fancy indexing included --
    and ``open_column_array`` is the one place the column streams open their files: ``.npy`` and ``.h5`` give the same array."""
    from climsim_b200.keras_h5 import open_h5_dataset, read_keras_h5
    from climsim_b200.stream import open_column_array
    want = read_keras_h5(MLP_H5)["weights"]
    k1 = open_h5_dataset(MLP_H5, "model_weights/dense_1/dense_1/kernel:0")
    assert isinstance(k1, np.memmap) and k1.shape == (768, 640) and k1.dtype == np.float32 and not k1.flags.writeable
    np.testing.assert_array_equal(k1, want[2])
    np.testing.assert_array_equal(k1[100:164], want[2][100:164])                       # a window of rows, as the stream reads
    np.testing.assert_array_equal(k1[[5, 700, 33]], want[2][[5, 700, 33]])
    with pytest.raises(KeyError):
        open_h5_dataset(MLP_H5, "data")
    # the dispatch: the same columns through a .npy file and through the .h5 dataset
    npy = tmp_path / "cols_input.npy"
    np.save(npy, want[2])
    a = open_column_array(str(npy))
    b = open_column_array(MLP_H5, dataset="model_weights/dense_1/dense_1/kernel:0")
    assert a.shape == b.shape and a.dtype == b.dtype
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


@needs_ref
def test_module_load_keras_h5_layouts_for_mlp_v1_and_ed():
    """``MLP.load_keras_h5`` / ``ED.load_keras_h5`` (one method of their common base) on stand-ins that only record the flat blob:
    MLP_v1 fuses the two output Dense layers column-wise, the encoder-decoder keeps Keras' layer list as it is."""
    import types
    from climsim_b200.baseline_models import _EngineModule
    from climsim_b200.keras_h5 import read_keras_h5

    def stand_in(n, **extra):
        box = types.SimpleNamespace(flat=torch.zeros(n), got=None, **extra)
        box.load_flat = lambda flat: setattr(box, "got", np.asarray(flat))
        return box

    mlp = stand_in(1753472, out_lin=120)
    ck = _EngineModule.load_keras_h5(mlp, MLP_H5)
    w = read_keras_h5(MLP_H5)["weights"]
    assert mlp.got.size == 1753472 and ck["optimizer"]["iterations"] == 6570
    np.testing.assert_array_equal(mlp.got[-128:], np.concatenate([w[-3], w[-1]]))              # fused bias [b_lin | b_relu]
    ed = stand_in(831879)
    _EngineModule.load_keras_h5(ed, ED_H5)
    we = read_keras_h5(ED_H5)["weights"]
    np.testing.assert_array_equal(ed.got, np.concatenate([a.reshape(-1) for a in we]))
    with pytest.raises(AssertionError, match="parameters"):
        _EngineModule.load_keras_h5(stand_in(10), ED_H5)
