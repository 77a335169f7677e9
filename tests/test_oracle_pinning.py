"""Pins the CPU oracle to the reference: golden vectors produced by running the reference's own code
(tests/golden/make_golden.py) and the known answers the reference holds."""
import os

import numpy as np
import pytest
import torch

from oracle import data_utils_ref as R
from oracle import models as M


@pytest.fixture(scope="module")
def du(golden_dir):
    return np.load(os.path.join(golden_dir, "data_utils.npz"))


@pytest.fixture(scope="module")
def hsr(golden_dir):
    return np.load(os.path.join(golden_dir, "hsr_small.npz"))


# ---------------------------------------------------------------- known answers held by the reference
def test_mlp_v1_param_count_matches_step1_results_csv():
    # baseline_v1/step1_analysis/step1_results.csv, lot-147 trial_0027: num_parameters = 1753472
    assert M.MLPRef().num_parameters() == 1_753_472


def test_mlp_v1_flops_match_flop_notebook():
    # step3_prediction/FLOP_calculation.ipynb cells 5-6: 3 503 488 FLOP / sample
    assert M.MLPRef().flops_per_sample() == 3_503_488


def test_cnn_param_count():
    # block1 505 470 + 11 x 1 155 070 + 4 070 + 110
    assert M.CNNRef().num_parameters() == 505_470 + 11 * 1_155_070 + 4_070 + 110


def test_ed_widths():
    assert M.ed_widths() == [463, 463, 231, 115, 57, 28, 5, 28, 57, 115, 231, 463, 463, 128]


def test_hsr_param_count_matches_shipped_checkpoint(golden_dir):
    g = np.load(os.path.join(golden_dir, "hsr_final_cp_outputs.npz"))
    net = M.HSRRef(124, 128, hidden_dims=1024, layers=4)
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"]) == 6_832_384


def test_cyclical_lr_closed_form():
    # triangular2-style schedule: starts at INIT_LR, peaks at MAX_LR after step_size, halves each cycle
    f = lambda s: M.cyclical_lr(s, 2.5e-4, 2.5e-3, step_size=10)
    assert f(0) == pytest.approx(2.5e-4)
    assert f(10) == pytest.approx(2.5e-3)
    assert f(20) == pytest.approx(2.5e-4)
    assert f(30) == pytest.approx(2.5e-4 + (2.5e-3 - 2.5e-4) / 2)
    assert f(50) == pytest.approx(2.5e-4 + (2.5e-3 - 2.5e-4) / 4)


# ---------------------------------------------------------------- data_utils restatement vs the reference class
def test_save_norm_vectors(du):
    assert du["inp_sub"].shape == (124,) and du["inp_div"].shape == (124,) and du["out_scale"].shape == (128,)
    assert du["inp_div"][60] == 0.0          # the max == min level built into the fixture


def test_normalize_nan_inf_rule(du):
    xn = R.normalize_input(du["x_raw"], du["inp_sub"], du["inp_div"])
    assert xn.dtype == np.float32
    np.testing.assert_array_equal(xn, du["x_renorm"])
    assert np.all(xn[:, 60] == 0)


def test_pressure_thickness(du):
    dp = R.pressure_thickness(du["x_norm"], float(du["ps_mean"]), float(du["ps_max"]), float(du["ps_min"]),
                              du["hyai"], du["hybi"], 1e5, int(du["ncol"]))
    np.testing.assert_array_equal(dp, du["dp_val"])


def test_output_weighting_and_just_weights(du):
    dp = du["dp_val"]
    for key, arr in (("tw_", du["target"]), ("pw_", du["pred"])):
        got = R.output_weighting(arr, dp, du["out_scale"], du["area_wgt"])
        for v in R.V1_OUTPUTS:
            np.testing.assert_array_equal(got[v], du[key + v])
    w = R.output_weights(dp, du["out_scale"], du["area_wgt"])
    np.testing.assert_array_equal(w, du["just_weights"])


def test_metrics(du):
    fns = {"MAE": R.calc_mae, "RMSE": R.calc_rmse, "R2": R.calc_r2, "bias": R.calc_bias}
    for v in R.V1_OUTPUTS:
        for name, f in fns.items():
            np.testing.assert_allclose(f(du["pw_" + v], du["tw_" + v]), du[f"{name}_{v}"], rtol=1e-14, atol=0)
    np.testing.assert_allclose(R.calc_crps(du["crps_samples"], du["tw_ptend_t"]), du["crps"], rtol=1e-13)


def test_cnn_reshapes(du):
    np.testing.assert_array_equal(R.reshape_input_for_cnn(du["x_norm"]), du["cnn_in"])
    np.testing.assert_array_equal(R.reshape_target_for_cnn(du["target"]), du["cnn_tgt"])
    np.testing.assert_array_equal(R.reshape_target_from_cnn(du["cnn_pred"]), du["cnn_pred_flat"])


# ---------------------------------------------------------------- HSR restatement vs the reference module
def _load_hsr(hsr, prefix):
    net = M.HSRRef(124, 128, hidden_dims=32, layers=2)
    sd = {k[len(prefix):]: torch.from_numpy(hsr[k]) for k in hsr.files if k.startswith(prefix)}
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net


def test_hsr_forward_losses_grads(hsr):
    net = _load_hsr(hsr, "init::")
    x, y = torch.from_numpy(hsr["x0"]), torch.from_numpy(hsr["y0"])
    mu, lp = net(x)
    np.testing.assert_allclose(mu.detach().numpy(), hsr["mu"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(lp.detach().numpy(), hsr["logprec"], rtol=1e-6, atol=1e-7)
    for mode in ("mse", "mle"):
        net.zero_grad()
        mu, lp = net(x)
        loss = M.HSRRef.loss(mu, lp, y, mle=(mode == "mle"))
        loss.backward()
        assert loss.item() == pytest.approx(float(hsr[f"loss_{mode}"]), rel=1e-6)
        for k, p in net.named_parameters():
            key = f"grad_{mode}::{k}"
            if key in hsr.files:
                np.testing.assert_allclose(p.grad.numpy(), hsr[key], rtol=1e-5, atol=1e-8)


def test_hsr_training_loop_matches_reference_trainer(hsr):
    """3 epochs x 2 batches of the reference's own ``trainer`` (MSE phase, then NLL phase; Adam + per-group L2)."""
    net = _load_hsr(hsr, "init::")
    alpha, beta = M.HSRRef.weight_decays(gamma=0.022)
    groups = [(list(net.mean.parameters()), alpha), (list(net.logprec.parameters()), beta)]
    state = [([torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]) for ps, _ in groups]
    batches = [(torch.from_numpy(hsr[f"x{i}"]), torch.from_numpy(hsr[f"y{i}"])) for i in range(2)]
    epochs, steps = 3, [0, 0]
    for epoch in range(epochs):
        for x, y in batches:
            for p in net.parameters():
                p.grad = None
            mu, lp = net(x)
            M.HSRRef.loss(mu, lp, y, mle=not (epoch < epochs / 3)).backward()
            for gi, ((ps, wd), (m, v)) in enumerate(zip(groups, state)):
                if ps[0].grad is None:      # MSE phase: the log-precision net gets no gradient and torch.optim.Adam
                    continue                # skips it entirely (its per-parameter step counter does not advance)
                steps[gi] += 1
                M.torch_adam_step(ps, [p.grad for p in ps], m, v, steps[gi], lr=1e-3, weight_decay=wd)
    final = {k[len("final::"):]: hsr[k] for k in hsr.files if k.startswith("final::")}
    for k, v in net.state_dict().items():
        np.testing.assert_allclose(v.numpy(), final[k], rtol=2e-5, atol=2e-7, err_msg=k)


@pytest.mark.skipif(not os.path.exists("/root/reference/baseline_models/HSR/model/final_hsr.cp"),
                    reason="reference checkpoint only exists in the build container")
def test_hsr_shipped_checkpoint_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "hsr_final_cp_outputs.npz"))
    net = M.HSRRef(124, 128, hidden_dims=1024, layers=4)
    sd = torch.load("/root/reference/baseline_models/HSR/model/final_hsr.cp", map_location="cpu", weights_only=True)
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        mu, lp = net(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(mu.numpy(), g["mu"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lp.numpy(), g["logprec"], rtol=1e-5, atol=1e-6)


def _online_ref(g):
    net = M.OnlineMLPRef(557, 368, [int(h) for h in g["hidden"]], 3, dropout=0.0, output_prune=True, strato_lev_out=15)
    net.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    return net.eval()


def test_online_mlp_matches_reference_class(golden_dir):
    """oracle OnlineMLPRef == the reference's own mlp.MLP (output_prune, last-eight ReLU) on the golden batch."""
    g = np.load(os.path.join(golden_dir, "online_mlp.npz"))
    net = _online_ref(g)
    with torch.no_grad():
        y_plain = net(torch.from_numpy(g["x_raw"]) * 0.1)
        y_net = net(torch.from_numpy(g["pre"]))
    np.testing.assert_allclose(y_plain.numpy(), g["y_plain"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(y_net.numpy(), g["y_net"], rtol=1e-6, atol=1e-6)
    assert (y_net.numpy()[:, 60:75] == 0).all() and (y_net.numpy()[:, -8:] >= 0).all()


def test_online_inference_wrapper_matches_notebook_class(golden_dir):
    """oracle OnlineWrapperRef == `NewModel` of v2_nn_wrapper.ipynb cell 5 (executed from the notebook by make_golden.py):
    exp transforms, normalisation with 0-width columns, nan / inf -> 0, pruning, RH clipping, output pruning, / out_scale."""
    g = np.load(os.path.join(golden_dir, "online_mlp.npz"))
    w = M.OnlineWrapperRef(_online_ref(g), g["input_sub"], g["input_div"], g["out_scale"], g["lbd_qc"], g["lbd_qi"])
    x = torch.from_numpy(g["x_raw"])
    with torch.no_grad():
        pre = w.preprocessing(x)
        y = w(x)
    np.testing.assert_array_equal(pre.numpy(), g["pre"])                      # element-wise fp32 arithmetic: bit-exact
    np.testing.assert_allclose(y.numpy(), g["y_wrapped"], rtol=1e-6, atol=1e-9)
    assert np.isfinite(g["pre"]).all() and (g["pre"][:, [7, 300]] == 0).all()   # the max == min columns came out as 0
    assert g["pre"][:, 60:120].min() >= 0 and g["pre"][:, 60:120].max() <= np.float32(1.2)


# ---------------------------------------------------------------- internal consistency of the unpinned restatements
def test_keras_adam_first_step_is_sign_step():
    p = [torch.tensor([1.0, -2.0, 3.0])]
    g = [torch.tensor([0.5, -0.25, 0.1])]
    m, v = [torch.zeros(3)], [torch.zeros(3)]
    M.keras_adam_step(p, g, m, v, t=1, lr=0.1)
    np.testing.assert_allclose(p[0].numpy(), [0.9, -1.9, 2.9], rtol=1e-4)


def test_weighted_mse_reduces_to_keras_mse_and_cnn_weights():
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(5, 128, generator=g), torch.randn(5, 128, generator=g)
    assert M.weighted_mse(a, b, torch.ones(128)).item() == pytest.approx(M.mse(a, b).item(), rel=1e-6)
    a, b = torch.randn(4, 60, 10, generator=g), torch.randn(4, 60, 10, generator=g)
    w = M.cnn_loss_weights()
    assert (w * (a - b) ** 2).sum(dim=(1, 2)).mean().item() == pytest.approx(M.mse_adjusted(a, b).item(), rel=1e-5)
    assert (w * (a - b).abs()).sum(dim=(1, 2)).mean().item() == pytest.approx(M.mae_adjusted(a, b).item(), rel=1e-5)


def test_conv1d_same_matches_direct_sum():
    g = torch.Generator().manual_seed(1)
    x, w, b = torch.randn(2, 7, 3, generator=g), torch.randn(3, 3, 4, generator=g), torch.randn(4, generator=g)
    y = M.conv1d_same_cl(x, w, b)
    ref = torch.zeros(2, 7, 4)
    for l in range(7):
        for t in range(3):
            s = l + t - 1
            if 0 <= s < 7:
                ref[:, l] += x[:, s] @ w[t]
    ref += b
    torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-6)


def test_manual_backward_matches_autograd():
    """The explicit backward used to emulate the bf16 rounding points is, without rounding, plain backprop."""
    for act in ("leakyrelu", "relu", "elu"):
        ref = M.MLPRef(units=(96, 64), act=act, seed=2)
        ref.randomize_biases(3)
        g = torch.Generator().manual_seed(4)
        x, y = 0.2 * torch.randn(50, 124, generator=g), 0.1 * torch.randn(50, 128, generator=g)
        w = torch.linspace(0.5, 2.0, 128)
        loss = M.weighted_mse(y, ref(x), w)
        loss.backward()
        l2, grads = ref.manual_train_step(x, y, w=w)
        assert l2.item() == pytest.approx(loss.item(), rel=1e-6)
        for p, gm in zip(ref.params, grads):
            torch.testing.assert_close(gm, p.grad, rtol=1e-4, atol=1e-9)


def test_keras_adam_algebra_against_torch_adam_with_eps_zero():
    """The Keras form (step size lr*sqrt(1-b2^t)/(1-b1^t), epsilon outside the bias correction) and torch.optim.Adam's form
    ((m/bc1)/(sqrt(v/bc2)+eps)) are the same update when epsilon = 0: an independent implementation (PyTorch's own optimizer) pins
    the moment updates and the bias-correction algebra of the restatement; only the epsilon placement is Keras-specific."""
    gen = torch.Generator().manual_seed(0)
    w0 = torch.randn(257, generator=gen, dtype=torch.float64)
    pk, m, v = [w0.clone()], [torch.zeros(257, dtype=torch.float64)], [torch.zeros(257, dtype=torch.float64)]
    pt = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([pt], lr=3e-3, betas=(0.9, 0.999), eps=0.0)
    for t in range(1, 8):
        g = torch.randn(257, generator=gen, dtype=torch.float64)
        M.keras_adam_step(pk, [g], m, v, t, lr=3e-3, eps=0.0)
        pt.grad = g.clone()
        opt.step()
        torch.testing.assert_close(pk[0], pt.detach(), rtol=1e-12, atol=1e-14)
    # with the Keras default epsilon the two differ by the epsilon placement only: tiny where |g| >> eps
    pk2, m2, v2 = [w0.clone()], [torch.zeros(257, dtype=torch.float64)], [torch.zeros(257, dtype=torch.float64)]
    M.keras_adam_step(pk2, [torch.ones(257, dtype=torch.float64)], m2, v2, 1, lr=3e-3, eps=1e-7)
    assert (pk2[0] - (w0 - 3e-3)).abs().max().item() < 1e-8


def test_conv1d_same_and_activations_against_torch_functional():
    """conv1d_same_cl (Keras Conv1D(padding='same'), channels-last, kernel (k, Cin, Cout)) == torch.nn.functional.conv1d with
    padding='same' on the transposed layout; ELU(alpha=1) / LeakyReLU(0.15) as the oracle models use them == torch's."""
    gen = torch.Generator().manual_seed(2)
    for k in (1, 3, 5):
        x, w, b = torch.randn(3, 60, 6, generator=gen), torch.randn(k, 6, 9, generator=gen), torch.randn(9, generator=gen)
        want = torch.nn.functional.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), b, padding="same").transpose(1, 2)
        torch.testing.assert_close(M.conv1d_same_cl(x, w, b), want, rtol=1e-5, atol=1e-5)
    ref = M.MLPRef(units=(32,), act="leakyrelu", seed=0)
    x = torch.randn(4, 124, generator=gen)
    h = torch.nn.functional.leaky_relu(x @ ref.params[0] + ref.params[1], 0.15)
    h = torch.nn.functional.leaky_relu(h @ ref.params[2] + ref.params[3], 0.15)
    want = torch.cat([h @ ref.params[4] + ref.params[5], torch.relu(h @ ref.params[6] + ref.params[7])], dim=1)
    torch.testing.assert_close(ref(x), want, rtol=1e-5, atol=1e-6)


def test_tf32_rounding_of_the_emulating_oracle():
    """``_tf32_round`` (the storage rounding of the CSB_TF32 engine restated): results have at most 10 explicit mantissa bits, the error
    is at most half a TF32 ulp (2^-11 relative), the map is idempotent and odd, exactly representable values are fixed points."""
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.randn(4096, generator=g) * 10.0 ** torch.randint(-6, 6, (4096,), generator=g).float(),
                   torch.tensor([0.0, 1.0, -1.0, 1.0 + 2.0 ** -10, 1.0 + 2.0 ** -11, 3.0, 0.15])])
    r = M._tf32_round(x)
    assert ((r.view(torch.int32) & 0x1FFF) == 0).all()                         # low 13 mantissa bits cleared
    nz = x != 0
    assert ((r[nz] - x[nz]).abs() / x[nz].abs()).max().item() <= 2.0 ** -11 + 1e-9
    assert torch.equal(M._tf32_round(r), r)
    assert torch.equal(M._tf32_round(-x), -r)
    assert r[-7:-4].tolist() == [0.0, 1.0, -1.0] and r[-4].item() == 1.0 + 2.0 ** -10 and r[-3].item() == 1.0 + 2.0 ** -10   # tie rounds up


def test_dropout_masks_of_ones_change_nothing_and_scale_like_torch_dropout():
    """The explicit ``masks=`` path of the oracles (the engine's keep decisions replayed): multipliers of 1 reproduce the dropout-free
    forward; a 0 / (1/(1-p)) mask is ``torch.nn.functional.dropout``'s own arithmetic (inverted dropout) for the same keep pattern,
    placed where the reference puts it -- behind LayerNorm and before the ReLU in hsr.MLP (hsr.py:20-25)."""
    torch.manual_seed(1)
    ref = M.MLPRef(units=(32, 16), seed=2)
    x = torch.randn(7, 124)
    ones = [torch.ones(7, n) for n in (32, 16, 128)]
    assert torch.equal(ref.forward(x, masks=ones), ref.forward(x))
    h = M.HSRMLPRef(124, 128, 24, 2)
    assert torch.allclose(h(x, masks=[torch.ones(7, 24)] * 2), h(x), atol=0, rtol=0)
    keep = (torch.rand(7, 24) > 0.3).float()
    lin, ln = h.linear0[0], h.linear0[1]
    u = ln(lin(x))
    want = torch.relu(u * keep / 0.7)                                          # Dropout(p=0.3) with this keep pattern, then ReLU
    got = torch.relu(h.linear0[1](h.linear0[0](x)) * (keep / 0.7))
    assert torch.allclose(got, want)
    full = h(x, masks=[keep / 0.7, torch.ones(7, 24)])
    assert torch.allclose(full, h.final_linear(torch.relu(h.linear1(want))), atol=1e-6)


def test_graphs_and_training_setups_match_the_reference_saved_keras_models(golden_dir):
    """TensorFlow cannot run here, but the reference SHIPS its trained Keras models, and their ``.h5`` files carry ``model_config`` and
    ``training_config`` (tests/golden/make_golden.py::make_keras_configs reads the JSON out of the files -> keras_configs.json):
    baseline_models/MLP/model/backup_phase-7_retrained_models_step2_lot-147_trial_0027.best.h5 and
    baseline_models/ED/model/ED_ClimSIM_1_3_model.h5.  The oracle's graphs, the engine presets and the optimizer / metric / schedule
    defaults are held to what those artefacts say -- the structure of the 'parity unpinned' Keras models is pinned by the
    reference's own files, not only by reading its scripts."""
    import inspect
    import json
    from climsim_b200 import trainer as T
    cfg = json.load(open(os.path.join(golden_dir, "keras_configs.json")))

    # ---- MLP_v1: Dense(768) LeakyReLU(0.15) ... Dense(128) LeakyReLU -> Dense(120, linear) || Dense(8, relu) -> Concatenate
    mlp = cfg["mlp_v1"]
    layers = mlp["layers"]
    assert layers[0]["class"] == "InputLayer" and layers[0]["input_shape"] == [None, 124]
    dense = [l for l in layers if l["class"] == "Dense"]
    leaky = [l for l in layers if l["class"] == "LeakyReLU"]
    ref = M.MLPRef()                                                             # the oracle's defaults = the shipped best trial
    assert [d["units"] for d in dense[:-2]] == list(ref.units) + [ref.out_lin + ref.out_relu] == [768, 640, 512, 640, 640, 128]
    assert all(d["activation"] == "linear" and d["use_bias"] for d in dense[:-2]) and len(leaky) == len(dense) - 2
    assert ref.act == "leakyrelu" and all(abs(l["alpha"] - ref.alpha) < 1e-7 for l in leaky)       # 0.15 stored as float32
    # every hidden Dense feeds its own LeakyReLU, which feeds the next Dense: a plain chain
    for i, d in enumerate(dense[1:-2], start=1):
        assert d["inbound"] == [leaky[i - 1]["name"]] and leaky[i]["inbound"] == [d["name"]]
    lin, rel, cat = dense[-2], dense[-1], layers[-1]
    assert (lin["units"], lin["activation"], rel["units"], rel["activation"]) == (ref.out_lin, "linear", ref.out_relu, "relu") == (120, "linear", 8, "relu")
    assert lin["inbound"] == rel["inbound"] == [leaky[-1]["name"]]              # both heads read the last hidden activation
    assert cat["class"] == "Concatenate" and cat["inbound"] == [lin["name"], rel["name"]]          # [120 linear | 8 relu], in this order
    assert all(d["kernel_initializer"] == "GlorotUniform" and d["bias_initializer"] == "Zeros" for d in dense)   # trainer.glorot_uniform_flat
    assert ref.num_parameters() == 1753472
    # training set-up: loss 'mse', metrics mse / mae / accuracy (= categorical accuracy: argmax agreement, csb_batch_metrics)
    assert mlp["loss"] == "mse" and mlp["metrics"] == ["mean_squared_error", "mean_absolute_error", "categorical_accuracy"]
    opt = mlp["optimizer"]
    assert opt["class_name"] == "Addons>RectifiedAdam"                           # the best trial's optimizer: the engine's CSB_OPT_RADAM rule
    oc = opt["config"]
    sig = inspect.signature(M.radam_step).parameters
    assert abs(oc["beta_1"] - sig["beta1"].default) < 1e-7 and abs(oc["beta_2"] - sig["beta2"].default) < 1e-7
    assert oc["epsilon"] == sig["eps"].default == 1e-7 and oc["sma_threshold"] == sig["sma_threshold"].default == 5.0
    assert oc["total_steps"] == 0 and oc["weight_decay"] == 0.0 and not oc["amsgrad"]          # no warm-up schedule inside the optimizer
    clr = oc["learning_rate"]
    assert clr["class_name"] == "Addons>CyclicalLearningRate" and clr["config"]["scale_mode"] == "cycle"
    d = inspect.signature(T.cyclical_lr).parameters
    assert (clr["config"]["initial_learning_rate"], clr["config"]["maximal_learning_rate"]) == (d["initial_lr"].default, d["max_lr"].default) == (2.5e-4, 2.5e-3)
    assert inspect.signature(M.keras_adam_step).parameters["eps"].default == 1e-7

    # ---- ED: 124 -> 463 -> 463 -> 231 -> 115 -> 57 -> 28 -> 5 | 28 -> 57 -> 115 -> 231 -> 463 -> 463 -> 128 (ReLU, ELU output)
    ed = cfg["ed"]
    dense = [l for l in ed["layers"] if l["class"] == "Dense"]
    assert [d["units"] for d in dense] == M.ed_widths() == [463, 463, 231, 115, 57, 28, 5, 28, 57, 115, 231, 463, 463, 128]
    assert [d["activation"] for d in dense] == ["relu"] * 13 + ["elu"]
    assert M.EDRef().num_parameters() == sum(k * n + n for k, n in zip([124] + M.ed_widths()[:-1], M.ed_widths()))
    assert ed["loss"] == "mean_squared_error" and ed["metrics"] == mlp["metrics"]
    eo = ed["optimizer"]
    assert eo["class_name"] == "Adam" and eo["config"]["epsilon"] == 1e-7 and not eo["config"]["amsgrad"]
    # the saved learning rate is where the divide-by-5-every-7-epochs schedule stands at the end of the 40-epoch run: 1e-4 / 5^5
    assert abs(eo["config"]["learning_rate"] - T.ed_step_lr(39)) <= 1e-6 * T.ed_step_lr(39)
    assert abs(T.ed_step_lr(39) - 1e-4 / 3125) < 1e-15 and T.ed_step_lr(0) == 1e-4 and abs(T.ed_step_lr(7) - 2e-5) < 1e-12

    # ---- CNN: the shipped SavedModel graph (no variables) still lists its layers: 12 residual blocks of
    # [Conv1D, Activation, Dropout, Conv1D, Activation, Dropout, Conv1D(1x1 on the block input), Add] + the output Conv1D + two Dense heads
    cnn = cfg["cnn"]
    ref = M.CNNRef()
    depth = ref.depth
    assert depth == 12
    assert cnn["layers"] == {"conv1d": 3 * depth + 1, "dense": 2, "dropout": 2 * depth, "activation": 2 * depth, "add": depth, "concatenate": 1}
    assert len(ref.params) == 2 * (cnn["layers"]["conv1d"] + cnn["layers"]["dense"])          # (kernel, bias) per Conv1D / Dense
    assert cnn["has_relu"] and cnn["has_elu"] and "SAME" in cnn["padding"] and cnn["padding_same_count"] >= cnn["layers"]["conv1d"]
