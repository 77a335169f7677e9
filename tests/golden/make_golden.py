#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE in this container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only); writes *.npz next to itself

The fixtures produced (all small enough to commit):

* ``data_utils.npz`` -- outputs of the reference class ``climsim_utils.data_utils.data_utils`` (imported from
  /root/reference with the absent third-party modules xarray / matplotlib / tensorflow / netCDF4 / h5py replaced by
  empty stubs; none of them is touched by the methods exercised) on seeded synthetic arrays: ``save_norm``,
  ``set_pressure_grid``, ``output_weighting`` (both modes), ``calc_MAE/RMSE/R2/bias/CRPS``, the three CNN reshape
  helpers.  The class needs xarray-like ``grid_info`` / normalisation objects; ``FakeDA`` below provides the few
  operators it uses (``.values``, ``.mean(dim=)``, ``/``, ``*``, ``len``).
* ``hsr_small.npz`` -- the reference's ``baseline_models/HSR/training/hsr.py``: ``HeteroskedasticRegression``
  (hidden 32, 2 layers) forward, both losses, gradients, and the end state of its own ``trainer`` loop (3 epochs x 2
  batches: MSE phase then NLL phase, Adam with the per-group weight decay).

* ``online_mlp.npz`` -- the reference's online MLP (``mlp.py``) and its E3SM inference wrapper (``v2_nn_wrapper.ipynb`` cell 5),
  see ``make_online``.

The GPU box has no /root/reference: tests only read the committed .npz files.
"""
import importlib.machinery
import os
import sys
import types

import numpy as np

REF = os.environ.get("CLIMSIM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


class FakeDA:
    """Minimal stand-in for the xarray.DataArray operations data_utils.py performs on grid/normalisation data."""

    def __init__(self, a):
        self.values = np.asarray(a, dtype=np.float64)

    def mean(self, dim=None):
        return FakeDA(self.values.mean())

    def __truediv__(self, o):
        return FakeDA(self.values / (o.values if isinstance(o, FakeDA) else o))

    def __mul__(self, o):
        return FakeDA(self.values * (o.values if isinstance(o, FakeDA) else o))

    __rmul__ = __mul__

    def __len__(self):
        return len(self.values)

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)


def stub_modules():
    for name in ["xarray", "matplotlib", "matplotlib.pyplot", "tensorflow", "netCDF4", "h5py", "seaborn"]:
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__spec__ = importlib.machinery.ModuleSpec(name, None)   # torch._dynamo calls find_spec on these
            sys.modules[name] = mod
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def synthetic_norm(rng):
    """Per-variable normalisation 'datasets' (dict var -> FakeDA), incl. one level with max == min."""
    lens = {"state_t": 60, "state_q0001": 60, "state_ps": 1, "pbuf_SOLIN": 1, "pbuf_LHFLX": 1, "pbuf_SHFLX": 1}
    base = {"state_t": (250.0, 30.0), "state_q0001": (5e-3, 4e-3), "state_ps": (9.8e4, 4e3),
            "pbuf_SOLIN": (400.0, 300.0), "pbuf_LHFLX": (80.0, 60.0), "pbuf_SHFLX": (20.0, 30.0)}
    mean, vmax, vmin = {}, {}, {}
    for v, n in lens.items():
        mu, sd = base[v]
        m = mu + sd * 0.1 * rng.standard_normal(n)
        hi = m + sd * (2 + rng.random(n))
        lo = m - sd * (2 + rng.random(n))
        if v == "state_q0001":
            hi[0] = lo[0] = m[0]                      # max == min -> division by zero -> nan/inf -> 0 rule
        sq = (lambda a: a if n > 1 else a[0])
        mean[v], vmax[v], vmin[v] = FakeDA(sq(m)), FakeDA(sq(hi)), FakeDA(sq(lo))
    olens = {"ptend_t": 60, "ptend_q0001": 60, "cam_out_NETSW": 1, "cam_out_FLWDS": 1, "cam_out_PRECSC": 1,
             "cam_out_PRECC": 1, "cam_out_SOLS": 1, "cam_out_SOLL": 1, "cam_out_SOLSD": 1, "cam_out_SOLLD": 1}
    scale = {}
    for v, n in olens.items():
        s = 10.0 ** rng.uniform(0, 7, size=n)
        scale[v] = FakeDA(s if n > 1 else s[0])
    return mean, vmax, vmin, scale


def make_data_utils():
    stub_modules()
    sys.path.insert(0, REF)
    from climsim_utils.data_utils import data_utils  # the reference's own class
    sys.modules.pop("tensorflow", None)              # only needed for the module-level import

    rng = np.random.default_rng(20260925)
    ncol, T = 6, 4
    hyai = np.sort(rng.random(61)) * 0.3
    hybi = np.sort(rng.random(61))
    grid = {"lev": FakeDA(np.arange(60)), "ncol": FakeDA(np.arange(ncol)), "area": FakeDA(0.5 + rng.random(ncol)),
            "lat": FakeDA(np.linspace(-60, 60, ncol)), "lon": FakeDA(np.linspace(0, 300, ncol)),
            "hyam": FakeDA(rng.random(60)), "hybm": FakeDA(rng.random(60)), "hyai": FakeDA(hyai),
            "hybi": FakeDA(hybi), "P0": FakeDA(1e5)}
    mean, vmax, vmin, scale = synthetic_norm(rng)
    du = data_utils(grid_info=grid, input_mean=mean, input_max=vmax, input_min=vmin, output_scale=scale,
                    ml_backend="pytorch")
    du.set_to_v1_vars()
    inp_sub, inp_div, out_scale = du.save_norm(write=False)

    n = ncol * T
    x_norm = np.float32(0.3 * rng.standard_normal((n, 124)))
    target = np.float32(0.1 * rng.standard_normal((n, 128)))
    pred = np.float32(target + 0.03 * rng.standard_normal((n, 128)))
    du.input_val, du.target_val = x_norm, target
    du.set_pressure_grid("val")
    tw = du.output_weighting(target, "val")
    pw = du.output_weighting(pred, "val")
    jw = du.output_weighting(target, "val", just_weights=True)
    out = {"ncol": ncol, "T": T, "hyai": hyai, "hybi": hybi, "area": grid["area"].values,
           "area_wgt": du.area_wgt, "inp_sub": inp_sub, "inp_div": inp_div, "out_scale": out_scale,
           "ps_mean": mean["state_ps"].values, "ps_max": vmax["state_ps"].values, "ps_min": vmin["state_ps"].values,
           "x_norm": x_norm, "target": target, "pred": pred, "dp_val": du.dp_val, "just_weights": jw}
    for v in du.target_vars:
        out["tw_" + v] = tw[v]
        out["pw_" + v] = pw[v]
        for mname, f in [("MAE", du.calc_MAE), ("RMSE", du.calc_RMSE), ("R2", du.calc_R2), ("bias", du.calc_bias)]:
            out[f"{mname}_{v}"] = np.asarray(f(pw[v], tw[v]))
    # the normalisation arithmetic itself (data_utils.py:806-809 + :894-897,:906) applied as the reference writes it
    x_raw = inp_sub + inp_div * x_norm.astype(np.float64) + 1e-3 * rng.standard_normal((n, 124))
    with np.errstate(divide="ignore", invalid="ignore"):
        xn = (x_raw - inp_sub) / inp_div           # what xarray broadcasting evaluates per variable/level
    xn[np.isinf(xn)] = 0
    xn[np.isnan(xn)] = 0
    out["x_raw"], out["x_renorm"] = x_raw, np.float32(xn)
    # CRPS on a (T, ncol, 60, S) ensemble
    samples = tw["ptend_t"][..., None] + rng.standard_normal(tw["ptend_t"].shape + (5,))
    out["crps_samples"], out["crps"] = samples, du.calc_CRPS(samples, tw["ptend_t"])
    # CNN reshapes (static methods)
    out["cnn_in"] = data_utils.reshape_input_for_cnn(x_norm)
    out["cnn_tgt"] = data_utils.reshape_target_for_cnn(target)
    cnn_pred = np.float32(rng.standard_normal((n, 60, 10)))
    out["cnn_pred"], out["cnn_pred_flat"] = cnn_pred, data_utils.reshape_target_from_cnn(cnn_pred)
    np.savez_compressed(os.path.join(HERE, "data_utils.npz"), **out)
    print("wrote data_utils.npz with", len(out), "arrays")


def make_hsr():
    import torch
    stub_modules()
    sys.path.insert(0, os.path.join(REF, "baseline_models", "HSR", "training"))
    import hsr as ref_hsr  # the reference's own module

    torch.manual_seed(7)
    net = ref_hsr.HeteroskedasticRegression(in_dims=124, out_dims=128, hidden_dims=32, layers=2, dropout=0)
    # LayerNorm affine params start at (1, 0): perturb so that they matter
    with torch.no_grad():
        for k, p in net.named_parameters():
            if ".1." in k:
                p.add_(0.1 * torch.randn_like(p))
    out = {"init::" + k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    batches = [{"x": 0.3 * torch.randn(16, 124, generator=g), "y": 0.1 * torch.randn(16, 128, generator=g)}
               for _ in range(2)]
    for i, b in enumerate(batches):
        out[f"x{i}"], out[f"y{i}"] = b["x"].numpy(), b["y"].numpy()
    # forward + the two losses + gradients on batch 0 (hsr.py:122-138)
    x, y = batches[0]["x"], batches[0]["y"]
    for mode in ("mse", "mle"):
        net.zero_grad()
        mu, logprec = net(x)
        prec = torch.exp(logprec)
        loss = ((y - mu) ** 2).mean() if mode == "mse" else (prec * (y - mu) ** 2 - logprec).mean()
        torch.clip(loss, min=-1e5, max=1e5).backward()
        out[f"loss_{mode}"] = np.float64(loss.item())
        for k, p in net.named_parameters():
            if p.grad is not None:
                out[f"grad_{mode}::" + k] = p.grad.detach().clone().numpy()
    out["mu"], out["logprec"] = mu.detach().numpy(), logprec.detach().numpy()
    # the reference's own training loop: 3 epochs x 2 batches (epoch 0 MSE, epochs 1-2 NLL), Adam + weight decay
    net.zero_grad()
    # the trainer keeps its per-step losses local: record them at the torch.clip(loss, ...) call it makes once per step (hsr.py:138)
    step_losses, real_clip = [], torch.clip

    def recording_clip(t, *a, **k):
        step_losses.append(float(t.detach()))
        return real_clip(t, *a, **k)

    torch.clip = recording_clip
    try:
        net.trainer(batches, epochs=3, save="/tmp/_hsr_golden.cp", plot=False, lr=1e-3, gamma=0.022)
    finally:
        torch.clip = real_clip
    out["trainer_losses"] = np.asarray(step_losses, dtype=np.float64)
    for k, v in net.state_dict().items():
        out["final::" + k] = v.detach().clone().numpy()
    np.savez_compressed(os.path.join(HERE, "hsr_small.npz"), **out)
    print("wrote hsr_small.npz with", len(out), "arrays")

    # real weights: outputs of the shipped checkpoint on a seeded batch (inputs regenerated from the seed in tests)
    cp = os.path.join(REF, "baseline_models", "HSR", "model", "final_hsr.cp")
    big = ref_hsr.HeteroskedasticRegression(in_dims=124, out_dims=128, hidden_dims=1024, layers=4, dropout=0)
    big.load_state_dict(torch.load(cp, map_location="cpu", weights_only=True))
    big.eval()
    xg = 0.3 * torch.randn(8, 124, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        mu, lp = big(xg)
    np.savez_compressed(os.path.join(HERE, "hsr_final_cp_outputs.npz"), x=xg.numpy(), mu=mu.numpy(), logprec=lp.numpy(),
                        n_params=np.int64(sum(p.numel() for p in big.parameters())))
    print("wrote hsr_final_cp_outputs.npz")


def make_online():
    """``online_mlp.npz``: the reference's online MLP (online_testing/baseline_models/MLP_v2rh/training/mlp.py, imported with a
    stand-in for the absent nvidia-modulus base classes, which add nothing to the arithmetic) and its E3SM inference wrapper
    ``NewModel`` (online_testing/model_postprocessing/v2_nn_wrapper.ipynb, cell 5, executed verbatim from the notebook) on a
    seeded raw batch: network outputs, wrapper pre-processing, wrapper outputs."""
    import dataclasses
    import json
    import torch

    modulus = types.ModuleType("modulus")
    modulus.__spec__ = importlib.machinery.ModuleSpec("modulus", None)

    class _Module(torch.nn.Module):
        def __init__(self, meta=None):
            super().__init__()

    @dataclasses.dataclass
    class _Meta:
        name: str = "model"

    modulus.Module, modulus.ModelMetaData = _Module, _Meta
    sys.modules["modulus"] = modulus
    sys.path.insert(0, os.path.join(REF, "online_testing", "baseline_models", "MLP_v2rh", "training"))
    import mlp as ref_mlp

    nb = json.load(open(os.path.join(REF, "online_testing", "model_postprocessing", "v2_nn_wrapper.ipynb")))
    src = "".join(nb["cells"][5]["source"])
    assert "class NewModel" in src
    ns = {"torch": torch, "nn": torch.nn, "np": np}
    exec(src, ns)
    NewModel = ns["NewModel"]

    torch.manual_seed(11)
    hidden = [48, 64, 40]
    net = ref_mlp.MLP(557, 368, hidden, 3, dropout=0.0, output_prune=True, strato_lev_out=15).eval()
    rng = np.random.default_rng(21)
    B = 24
    x_raw = rng.normal(0.0, 1.0, size=(B, 557)).astype(np.float32)
    x_raw[:, 60:120] = rng.uniform(-0.2, 1.6, size=(B, 60))                   # relative humidity incl. values outside [0, 1.2]
    x_raw[:, 120:240] = np.abs(rng.normal(0.0, 3e-5, size=(B, 120)))           # cloud liquid / ice mixing ratios
    input_sub = rng.normal(0.0, 0.3, size=557).astype(np.float32)
    input_div = rng.uniform(0.5, 2.0, size=557).astype(np.float32)
    input_div[[7, 300]] = 0.0                                                  # max == min columns -> inf / nan -> 0
    input_sub[7] = x_raw[0, 7]                                                 # 0 / 0 -> nan in row 0
    out_scale = np.exp(rng.uniform(0.0, np.log(1e4), size=368)).astype(np.float32)
    lbd_qc = np.exp(rng.uniform(np.log(1e4), np.log(1e6), size=60)).astype(np.float32)
    lbd_qi = np.exp(rng.uniform(np.log(1e4), np.log(1e6), size=60)).astype(np.float32)
    wrapper = NewModel(net, input_sub, input_div, out_scale, lbd_qc, lbd_qi).eval()
    with torch.no_grad():
        xt = torch.from_numpy(x_raw)
        pre = wrapper.preprocessing(xt.clone())
        y_net = net(pre.clone())
        y_wrapped = wrapper(xt.clone())
        y_plain = net(torch.from_numpy(x_raw[:, :].copy()) * 0.1)              # the bare network on already-normalised inputs
    out = {"x_raw": x_raw, "input_sub": input_sub, "input_div": input_div, "out_scale": out_scale, "lbd_qc": lbd_qc, "lbd_qi": lbd_qi,
           "pre": pre.numpy(), "y_net": y_net.numpy(), "y_wrapped": y_wrapped.numpy(), "y_plain": y_plain.numpy(),
           "hidden": np.asarray(hidden, np.int64)}
    for k, v in net.state_dict().items():
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "online_mlp.npz"), **out)
    print("wrote online_mlp.npz with", len(out), "arrays; state_dict keys:", list(net.state_dict().keys()))


def _json_objects_in(path, min_len=100):
    """The JSON documents a Keras ``.h5`` carries as HDF5 string attributes (``model_config``, ``training_config``), found by scanning
    the file for decodable objects -- h5py is not installed here, and the attributes are stored contiguously."""
    import json
    blob, out, i, dec = open(path, "rb").read(), [], 0, json.JSONDecoder()
    while True:
        i = blob.find(b'{"', i)
        if i < 0:
            return out
        try:
            obj, end = dec.raw_decode(blob[i:i + 400000].decode("utf-8", errors="ignore"))
            if isinstance(obj, dict) and end > min_len:
                out.append(obj)
                i += end
                continue
        except ValueError:
            pass
        i += 2


def make_keras_configs():
    """``keras_configs.json``: what the reference's OWN saved Keras models say about the graphs and training set-ups the oracle restates
    (TensorFlow cannot run here, but the shipped artefacts carry their ``model_config`` / ``training_config``):
    baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_trial_0027.best.h5 (MLP_v1, the best HPO trial) and
    baseline_models/ED/model/ED_ClimSIM_1_3_model.h5 (encoder-decoder)."""
    import json

    def layers_of(cfg):
        out = []
        for layer in cfg["config"]["layers"]:
            if layer["class_name"] in ("Functional", "Sequential", "Model"):
                out += layers_of(layer)
            else:
                c = layer["config"]
                out.append({"class": layer["class_name"], "name": c.get("name"), "units": c.get("units"), "activation": c.get("activation"),
                            "alpha": c.get("alpha"), "use_bias": c.get("use_bias"),
                            "kernel_initializer": (c.get("kernel_initializer") or {}).get("class_name"),
                            "bias_initializer": (c.get("bias_initializer") or {}).get("class_name"),
                            "input_shape": c.get("batch_input_shape"),
                            "inbound": [n[0] for n in layer["inbound_nodes"][0]] if layer.get("inbound_nodes") else []})
        return out

    res = {}
    for key, rel in (("mlp_v1", "baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_trial_0027.best.h5"),
                     ("ed", "baseline_models/ED/model/ED_ClimSIM_1_3_model.h5")):
        objs = _json_objects_in(os.path.join(REF, rel))
        model = [o for o in objs if "config" in o and "layers" in o.get("config", {})][0]
        train = [o for o in objs if "optimizer_config" in o][0]
        res[key] = {"source": rel, "model_name": model["config"]["name"], "layers": layers_of(model),
                    "loss": train["loss"], "metrics": [m["config"]["fn"] for m in train["metrics"][0]],
                    "optimizer": train["optimizer_config"]}
    # the CNN ships as a TensorFlow SavedModel graph without its variables (baseline_models/CNN/model/saved_model.pb): the layer
    # inventory is still in the node names (conv1d .. conv1d_36/..., dropout_23/..., add_11/...)
    import re
    pb = open(os.path.join(REF, "baseline_models/CNN/model/saved_model.pb"), "rb").read()

    def layer_indices(base):
        return sorted({int(m.group(1)) if m.group(1) else 0 for m in re.finditer(rb"(?<![a-z_])" + base.encode() + rb"(?:_(\d{1,3}))?/", pb)})

    res["cnn"] = {"source": "baseline_models/CNN/model/saved_model.pb",
                  "layers": {k: len(layer_indices(k)) for k in ("conv1d", "dense", "dropout", "activation", "add", "concatenate")},
                  "padding": sorted({m.decode() for m in re.findall(rb"SAME|VALID", pb)}),
                  "padding_same_count": len(re.findall(rb"SAME", pb)), "has_elu": b"Elu" in pb, "has_relu": b"Relu" in pb}
    with open(os.path.join(HERE, "keras_configs.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote keras_configs.json:", {k: len(v["layers"]) for k, v in res.items()})


def make_keras_weights():
    """``mlp_v1_shipped_f16.npz``: the reference's shipped MLP_v1 (baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_
    trial_0027.best.h5, read with climsim_b200.keras_h5) with every array rounded to float16 -- 3.3 MB instead of 7 MB -- so that the
    GPU box, which has no reference checkout, can run the engines on a REAL trained model (non-zero biases, dead units, the trained
    weight spectrum) instead of Glorot noise.  Plus per-array fingerprints of the exact fp32 values (sum, abs-sum) that tie the fixture
    to the file."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from climsim_b200.keras_h5 import read_keras_h5
    ck = read_keras_h5(os.path.join(REF, "baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_trial_0027.best.h5"))
    out = {f"w{i:02d}": w.astype(np.float16) for i, w in enumerate(ck["weights"])}
    out["fingerprint"] = np.array([[float(w.astype(np.float64).sum()), float(np.abs(w.astype(np.float64)).sum())] for w in ck["weights"]])
    out["iterations"] = np.array(ck["optimizer"]["iterations"])
    np.savez_compressed(os.path.join(HERE, "mlp_v1_shipped_f16.npz"), **out)
    print("wrote mlp_v1_shipped_f16.npz:", sum(w.size for w in ck["weights"]), "parameters")


if __name__ == "__main__":
    make_data_utils()
    make_hsr()
    make_online()
    make_keras_configs()
    make_keras_weights()
