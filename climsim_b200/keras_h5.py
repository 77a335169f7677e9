"""Read the reference's Keras ``.h5`` checkpoints without h5py or TensorFlow.

The reference saves its trained emulators with ``ModelCheckpoint`` / ``model.save`` as HDF5 (ref:
baseline_models/MLP/training/HPO/baseline_v1/step2_retrain/step2_retrain.py:252-262; the shipped files are
baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_trial_0027.best.h5 and
baseline_models/ED/model/ED_ClimSIM_1_3_model.h5).  Neither h5py nor TensorFlow exists in this image, so this module parses just
the part of HDF5 such a file uses -- version-0 superblock, version-1 object headers (with continuation blocks), old-style groups
(symbol-table B-trees + local heaps), contiguous little-endian fp32 / int64 datasets -- and the two JSON attributes
(``model_config``, ``training_config``) Keras stores beside them.  ``read_keras_h5`` returns the weights in ``get_weights()`` order
(per Dense layer in model order: kernel (in, out), bias), which is what ``MLPEngine.keras_to_flat`` / ``MLP.load_keras_weights``
take, plus the optimizer's slot variables in the same order -- so the reference's shipped model (and a run's ``checkpoint_best.h5``)
loads into the engine, optimizer state included."""
import json
import mmap
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class _H5:
    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:                                 # mapped, not read: the column files are tens of GB
            try:
                self.b = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            except ValueError:                                      # empty file
                self.b = b""
        if self.b[:8] != _SIG:
            raise ValueError(f"{path}: not an HDF5 file")
        if self.b[8] != 0 or self.b[13] != 8 or self.b[14] != 8:
            raise ValueError(f"{path}: only version-0 superblocks with 8-byte offsets are understood (superblock version {self.b[8]})")
        # superblock v0: signature 8, versions / sizes 8, group K values 4, flags 4, base / free-space / EOF / driver addresses 32,
        # then the root group's symbol-table entry (link name offset 8, object header address 8, ...)
        self.root = struct.unpack_from("<Q", self.b, 56 + 8)[0]

    def _messages(self, addr: int) -> List[Tuple[int, int, int]]:
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHIi", b, addr)
        if ver != 1:
            raise ValueError(f"object header version {ver} at {addr} (only version 1 is understood)")
        out, blocks = [], [(addr + 16, hsize)]
        while blocks:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and len(out) < nmsg + 64:
                mtype, msize, _ = struct.unpack_from("<HHB", b, off)
                body = off + 8
                if mtype == 0x10:                                  # continuation block
                    blocks.append(struct.unpack_from("<QQ", b, body))
                else:
                    out.append((mtype, body, msize))
                off = body + msize
        return out

    def _heap_str(self, heap: int, offset: int) -> str:
        if self.b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        start = struct.unpack_from("<Q", self.b, heap + 24)[0] + offset
        return self.b[start:self.b.find(b"\x00", start)].decode()

    def _children(self, btree: int, heap: int) -> List[Tuple[str, int]]:
        b, out = self.b, []

        def walk(node: int) -> None:
            if b[node:node + 4] != b"TREE":
                raise ValueError("bad group B-tree node")
            level, used = struct.unpack_from("<BH", b, node + 5)
            for i in range(used):
                child = struct.unpack_from("<Q", b, node + 24 + 8 + i * 16)[0]
                if level > 0:
                    walk(child)
                    continue
                if b[child:child + 4] != b"SNOD":
                    raise ValueError("bad symbol-table node")
                for k in range(struct.unpack_from("<H", b, child + 6)[0]):
                    name_off, hdr = struct.unpack_from("<QQ", b, child + 8 + k * 40)
                    out.append((self._heap_str(heap, name_off), hdr))

        walk(btree)
        return out

    def datasets(self, hdr: Optional[int] = None, prefix: str = "") -> Iterator[Tuple[str, np.ndarray]]:
        """(path, array copy) of every contiguous numeric dataset below the object at ``hdr`` (default: the root group)."""
        for name, shape, np_dtype, addr in self.dataset_infos(hdr, prefix):
            n = int(np.prod(shape)) if shape else 1
            yield name, np.frombuffer(self.b, dtype=np_dtype, count=n, offset=addr).reshape(shape).copy()

    def dataset_infos(self, hdr: Optional[int] = None, prefix: str = "") -> Iterator[Tuple[str, tuple, str, int]]:
        """(path, shape, numpy dtype string, byte offset of the contiguous data) of every such dataset."""
        hdr = self.root if hdr is None else hdr
        msgs = self._messages(hdr)
        sym = [m for m in msgs if m[0] == 0x11]
        if sym:
            btree, heap = struct.unpack_from("<QQ", self.b, sym[0][1])
            for name, child in self._children(btree, heap):
                yield from self.dataset_infos(child, prefix + "/" + name)
            return
        shape = dtype = addr = None
        for mtype, body, _ in msgs:
            if mtype == 0x01:                                      # dataspace
                rank = self.b[body + 1]
                shape = struct.unpack_from("<%dQ" % rank, self.b, body + (8 if self.b[body] == 1 else 4))
            elif mtype == 0x03:                                    # datatype: class (0 fixed-point, 1 floating-point), size
                dtype = (self.b[body] & 0x0F, struct.unpack_from("<I", self.b, body + 4)[0])
            elif mtype == 0x08 and self.b[body] == 3 and self.b[body + 1] == 1:      # data layout v3, contiguous
                addr = struct.unpack_from("<Q", self.b, body + 2)[0]
        if shape is None or addr is None or addr == 0xFFFFFFFFFFFFFFFF:
            return
        np_dtype = {(1, 4): "<f4", (0, 8): "<i8", (1, 8): "<f8", (0, 4): "<i4"}.get(dtype)
        if np_dtype is None:
            return
        yield prefix, tuple(int(d) for d in shape), np_dtype, int(addr)

    def json_attributes(self, min_len: int = 100) -> List[dict]:
        """The JSON documents stored as string attributes (Keras: ``model_config``, ``training_config``)."""
        out, i, dec = [], 0, json.JSONDecoder()
        while True:
            i = self.b.find(b'{"', i)
            if i < 0:
                return out
            try:
                obj, end = dec.raw_decode(self.b[i:i + 400000].decode("utf-8", errors="ignore"))
                if isinstance(obj, dict) and end > min_len:
                    out.append(obj)
                    i += end
                    continue
            except ValueError:
                pass
            i += 2


def _dense_layers(model_config: dict) -> List[dict]:
    out = []
    for layer in model_config["config"]["layers"]:
        if layer["class_name"] in ("Functional", "Sequential", "Model"):
            out += _dense_layers(layer)
        elif layer["class_name"] == "Dense":
            out.append(layer["config"])
    return out


def read_keras_h5(path: str) -> Dict[str, object]:
    """``{"weights": [kernel0, bias0, kernel1, bias1, ...]`` (``model.get_weights()`` order, kernels (in, out)), ``"layers"``: the Dense
    layers' configs in model order, ``"model_config"``, ``"training_config"``, ``"optimizer"``: ``{"name", "iterations", "m": [...],
    "v": [...]}`` (slot variables in the order of ``weights``; absent when the file holds weights only)}."""
    h5 = _H5(path)
    docs = h5.json_attributes()
    model = next((d for d in docs if "config" in d and "layers" in d.get("config", {})), None)
    if model is None:
        raise ValueError(f"{path}: no Keras model_config attribute found")
    training = next((d for d in docs if "optimizer_config" in d), None)
    data = dict(h5.datasets())
    layers = _dense_layers(model)

    def find(kind: str, layer: str, what: str, suffix: str = "") -> np.ndarray:
        hits = [k for k in data if k.startswith("/" + kind + "/") and k.endswith(f"/{layer}/{what}{suffix}")]
        if len(hits) != 1:
            raise KeyError(f"{path}: {kind} entry for layer {layer!r} {what}{suffix}: found {hits}")
        return data[hits[0]]

    weights: List[np.ndarray] = []
    for cfg in layers:
        weights += [find("model_weights", cfg["name"], "kernel", ":0"), find("model_weights", cfg["name"], "bias", ":0")]
    res: Dict[str, object] = {"weights": weights, "layers": layers, "model_config": model, "training_config": training}
    opt_keys = [k for k in data if k.startswith("/optimizer_weights/")]
    if opt_keys:
        name = opt_keys[0].split("/")[2]
        it = [data[k] for k in opt_keys if k.endswith("/iter:0")]
        m, v = [], []
        try:
            for cfg in layers:
                for what in ("kernel", "bias"):
                    m.append(find("optimizer_weights", cfg["name"], what, "/m:0"))
                    v.append(find("optimizer_weights", cfg["name"], what, "/v:0"))
        except KeyError:
            m, v = [], []                                          # an optimizer without (m, v) slots
        res["optimizer"] = {"name": name, "iterations": int(it[0]) if it else 0, "m": m, "v": v}
    return res


def open_h5_dataset(path: str, name: str = "data") -> np.memmap:
    """A read-only ``np.memmap`` over one contiguous dataset of an HDF5 file -- the reference writes its column arrays this way
    (climsim_utils/data_utils.py:912-913, 924-925: ``h5py.File(..., 'w').create_dataset('data', data=npy_input)``) and reads them row by
    row through h5py (online_testing/baseline_models/MLP_v2rh/training/climsim_datapip_h5.py:104-126).  The mapping behaves like
    ``np.load(..., mmap_mode='r')`` of the ``.npy`` twin, so the column streams take either file."""
    want = "/" + name.strip("/")
    for ds, shape, np_dtype, addr in _H5(path).dataset_infos():
        if ds == want:
            return np.memmap(path, dtype=np_dtype, mode="r", offset=addr, shape=shape)
    raise KeyError(f"{path}: no contiguous numeric dataset named {name!r}")
