"""Device kernels behind the data_utils mirror and the nn.Module shims, against the reference's golden vectors / the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "data_utils.npz"))


def test_cnn_reshape_kernels_bit_exact(g):
    from climsim_b200.data_utils import data_utils
    x, y, p = (torch.from_numpy(g[k]).cuda() for k in ("x_norm", "target", "cnn_pred"))
    np.testing.assert_array_equal(data_utils.reshape_input_for_cnn(x).cpu().numpy(), g["cnn_in"])
    np.testing.assert_array_equal(data_utils.reshape_target_for_cnn(y).cpu().numpy(), g["cnn_tgt"])
    # mean over 60 levels: the reference reduces a strided fp32 slice with NumPy's pairwise sum; a different summation
    # order moves the result by at most a few ulp
    got = data_utils.reshape_target_from_cnn(p).cpu().numpy()
    np.testing.assert_array_equal(got[:, :120], g["cnn_pred_flat"][:, :120])
    np.testing.assert_allclose(got[:, 120:], g["cnn_pred_flat"][:, 120:], rtol=2e-6, atol=1e-7)


def test_normalize_kernel_matches_reference_rule(g):
    from climsim_b200 import _lib
    lib = _lib.load()
    x_raw = torch.from_numpy(g["x_raw"].astype(np.float32)).cuda()
    sub = torch.from_numpy(g["inp_sub"].astype(np.float32)).cuda()
    div = torch.from_numpy(g["inp_div"].astype(np.float32)).cuda()
    out = torch.empty_like(x_raw)
    _lib.check(lib.csb_normalize(x_raw.data_ptr(), sub.data_ptr(), div.data_ptr(), out.data_ptr(), x_raw.shape[0], 124, None), "csb_normalize")
    got = out.cpu().numpy()
    assert np.all(got[:, 60] == 0)                               # max == min level -> inf/nan -> 0
    # fp32 arithmetic on fp32 inputs vs the reference's fp64 arithmetic cast to fp32: raw values ~1e5 with spans ~1e4
    # lose ~3 digits in the subtraction, so compare in units of the input's own fp32 resolution
    want = g["x_renorm"]
    scale = np.abs(g["x_raw"]).max(axis=0) / np.where(g["inp_div"] == 0, 1, np.abs(g["inp_div"]))
    assert np.all(np.abs(got - want) <= 4 * np.finfo(np.float32).eps * (scale + 1))


def test_nn_module_autograd_matches_oracle():
    from climsim_b200.baseline_models import MLP
    from climsim_b200.synthetic import synthetic_batch
    from oracle import models as M
    units, B = (192, 128, 64), 300
    ref = M.MLPRef(units=units, seed=0)
    ref.randomize_biases(1)
    net = MLP(units=units, dtype="fp32", max_batch=512)
    net.load_keras_weights([p.detach().numpy() for p in ref.params])
    x, y = synthetic_batch(B, 0)
    loss_ref = M.mse(y, ref(x))
    loss_ref.backward()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    out = net(x.cuda())
    loss = ((out - y.cuda()) ** 2).mean()          # any torch loss works on top of the engine's forward
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * loss_ref.item()
    from climsim_b200 import MLPEngine
    g_ref = MLPEngine.keras_to_flat([p.grad.numpy() for p in ref.params])
    g_got = net.flat.grad.cpu().numpy()
    assert np.abs(g_got - g_ref).max() <= 1e-5 * np.abs(g_ref).max()
    opt.step()                                      # in-place update -> the module re-uploads on the next forward
    with torch.no_grad():
        for p in ref.params:
            p -= 0.1 * p.grad
        out2 = net(x.cuda()).cpu().numpy()
    want2 = ref(x).detach().numpy()
    assert np.abs(out2 - want2).max() <= 1e-5 * np.abs(want2).max()
    ws = net.keras_weights()
    assert [w.shape for w in ws][-4:] == [(128, 120), (120,), (128, 8), (8,)]


def test_ed_preset_matches_oracle():
    from climsim_b200.baseline_models import ED
    from climsim_b200.synthetic import synthetic_batch
    from oracle import models as M
    ref = M.EDRef(seed=0)
    assert ref.num_parameters() == 829_032 + sum(M.ed_widths())    # MACs + biases
    net = ED(dtype="fp32", max_batch=256)
    net.load_flat(np.concatenate([p.detach().numpy().reshape(-1) for p in ref.params]))
    x, _ = synthetic_batch(200, 1)
    want = ref(x).detach().numpy()
    with torch.no_grad():
        got = net(x.cuda()).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    net16 = ED(dtype="bf16", max_batch=256)
    net16.load_flat(np.concatenate([p.detach().numpy().reshape(-1) for p in ref.params]))
    with torch.no_grad():
        got16 = net16(x.cuda()).cpu().numpy()
    assert np.abs(got16 - want).max() <= 5e-2 * np.abs(want).max()


@pytest.mark.parametrize("dtype,tol_out,tol_g", [("fp32", 2e-5, 1e-4), ("bf16", 5e-2, 1e-1)])
def test_hsr_layernorm_model_against_reference_golden(golden_dir, dtype, tol_out, tol_g):
    """The LayerNorm MLP pair with the weights, inputs and expected outputs / gradients of the REFERENCE's own
    HeteroskedasticRegression (tests/golden/hsr_small.npz, produced by running hsr.py): forward, MSE loss and NLL loss."""
    from climsim_b200.baseline_models import HSR
    g = np.load(os.path.join(golden_dir, "hsr_small.npz"))
    sd = {k[len("init::"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init::")}
    net = HSR(124, 128, hidden_dims=32, layers=2, dtype=dtype, max_batch=64)
    net.load_reference_state_dict(sd)
    x, y = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["y0"]).cuda()
    mu, lp = net(x)
    scale = lambda a: np.abs(a).max()
    assert np.abs(mu.detach().cpu().numpy() - g["mu"]).max() <= tol_out * scale(g["mu"])
    assert np.abs(lp.detach().cpu().numpy() - g["logprec"]).max() <= tol_out * scale(g["logprec"])
    for mode in ("mse", "mle"):
        net.zero_grad()
        mu, lp = net(x)
        loss = ((y - mu) ** 2).mean() if mode == "mse" else (torch.exp(lp) * (y - mu) ** 2 - lp).mean()
        torch.clip(loss, min=-1e5, max=1e5).backward()
        assert abs(loss.item() - float(g[f"loss_{mode}"])) <= max(tol_out, 1e-5) * abs(float(g[f"loss_{mode}"]))
        for name, sub in (("mean", net.mean), ("logprec", net.logprec)):
            if sub.flat.grad is None:
                continue
            views, off, got = sub.layer_views(), 0, sub.flat.grad.cpu().numpy()
            keys = []
            for i in range(2):
                keys += [(f"{name}.linear{i}.0.weight", True), (f"{name}.linear{i}.0.bias", False),
                         (f"{name}.linear{i}.1.weight", False), (f"{name}.linear{i}.1.bias", False)]
            keys += [(f"{name}.final_linear.weight", True), (f"{name}.final_linear.bias", False)]
            for (key, transpose), v in zip(keys, views):
                want = g[f"grad_{mode}::{key}"]
                want = want.T if transpose else want
                part = got[off:off + v.numel()].reshape(want.shape)
                off += v.numel()
                err = np.linalg.norm(part - want) / max(np.linalg.norm(want), 1e-12)
                assert err <= tol_g, (mode, key, err)


def test_hsr_trainer_matches_reference_trainer_end_state(golden_dir, capsys, monkeypatch):
    """HSR.trainer (same arguments as hsr.py:83-142) from the reference's initial weights on the reference's two batches: after
    3 epochs x 2 batches (MSE phase, then NLL; Adam with the per-group weight decay) the parameters equal the end state of the
    REFERENCE's own trainer run (tests/golden/hsr_small.npz) and every per-step loss equals the one the reference computed.
    fp32 engine: Adam divides by sqrt(v), so elements whose gradient is cancellation noise may move by O(lr) either way -- the bound
    is per tensor in relative L2.  The whole step (both networks, loss, clip, per-group L2 Adam) is csb_hsr_train_step: torch's
    optimizers and autograd are made to raise if anything touches them."""
    from climsim_b200.baseline_models import HSR
    g = np.load(os.path.join(golden_dir, "hsr_small.npz"))
    sd = {k[len("init::"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init::")}
    net = HSR(124, 128, hidden_dims=32, layers=2, dtype="fp32", max_batch=64)
    net.load_reference_state_dict(sd)
    batches = [{"x": torch.from_numpy(g[f"x{i}"]), "y": torch.from_numpy(g[f"y{i}"])} for i in range(2)]

    def forbidden(*a, **k):
        raise AssertionError("HSR.trainer must not fall back to torch optimizers / autograd")

    monkeypatch.setattr(torch.optim, "Adam", forbidden)
    monkeypatch.setattr(torch.optim, "SGD", forbidden)
    monkeypatch.setattr(torch.Tensor, "backward", forbidden)
    losses = net.trainer(batches, epochs=3, save=os.devnull, plot=False, lr=1e-3, gamma=0.022)
    monkeypatch.undo()
    assert len(losses) == 6 and "alpha:" in capsys.readouterr().out
    np.testing.assert_allclose(losses, g["trainer_losses"], rtol=2e-4)       # epochs 0: MSE; 1, 2: Gaussian NLL (hsr.py:128-136)
    got = net.reference_state_dict()
    final = {k[len("final::"):]: g[k] for k in g.files if k.startswith("final::")}
    assert sorted(got) == sorted(final)
    for k, want in final.items():
        have = got[k].detach().cpu().numpy()
        assert have.shape == want.shape, k
        err = np.linalg.norm(have - want) / max(np.linalg.norm(want), 1e-12)
        assert err <= 2e-4, (k, err)
        assert np.abs(have - want).max() <= 3e-3, k                     # nothing moved further than three Adam steps of lr


@pytest.mark.parametrize("hidden,layers,B", [(1024, 2, 333), (200, 1, 1000), (512, 3, 129)])
def test_layernorm_mlp_bf16_wide_rows(hidden, layers, B):
    """bf16 LayerNorm kernels (a row per warp held in registers, fused du -> dz + parameter gradients) at the widths the small golden
    model does not reach: 1024 (four 256-column passes, the shipped HSR width), 200 (statistics over N = 200 of 256 padded columns),
    512; forward and every gradient tensor against the fp32 torch restatement of hsr.MLP."""
    from climsim_b200.baseline_models import HSRMLP
    from oracle import models as M
    torch.manual_seed(3)
    ref = M.HSRMLPRef(124, 128, hidden, layers)
    with torch.no_grad():
        for i in range(layers):                                   # non-trivial gamma / beta
            ln = getattr(ref, f"linear{i}")[1]
            ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.3, 0.3)
    net = HSRMLP(124, 128, hidden_dims=hidden, layers=layers, dtype="bf16", max_batch=1024)
    net.load_reference_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(4)
    x, y = 0.5 * torch.randn(B, 124, generator=g), 0.3 * torch.randn(B, 128, generator=g)
    want = ref(x)
    ((want - y) ** 2).mean().backward()
    got = net(x.cuda())
    ((got - y.cuda()) ** 2).mean().backward()
    assert (got.detach().cpu() - want.detach()).abs().max().item() <= 5e-2 * want.abs().max().item()
    want_g = []
    for i in range(layers):
        seq = getattr(ref, f"linear{i}")
        want_g += [seq[0].weight.grad.t().reshape(-1), seq[0].bias.grad, seq[1].weight.grad, seq[1].bias.grad]
    want_g += [ref.final_linear.weight.grad.t().reshape(-1), ref.final_linear.bias.grad]
    off, got_g = 0, net.flat.grad.cpu()
    for i, w in enumerate(want_g):
        part = got_g[off:off + w.numel()]
        off += w.numel()
        err = (part - w).norm().item() / max(w.norm().item(), 1e-12)
        assert err <= 1e-1, (i, err)
    assert off == got_g.numel()


@pytest.mark.parametrize("dtype,tol,tol_g", [("fp32", 2e-5, 1e-4), ("bf16", 3e-2, 1e-1)])
@pytest.mark.parametrize("loss", ["huber", "mse", "mae"])
def test_online_mlp_surface(dtype, tol, tol_g, loss):
    """The online MLP (557 -> [384, 1024, 640] -> 368, output_prune, last-8 ReLU) with Huber / MSE / L1 criteria
    (train_mlp_h5loader.py:222-232) through the fused train step, against the torch restatement of mlp.py."""
    from climsim_b200.baseline_models import OnlineMLP
    from oracle import models as M
    torch.manual_seed(0)
    ref = M.OnlineMLPRef(557, 368, [384, 1024, 640], 3, output_prune=True)
    net = OnlineMLP(557, 368, [384, 1024, 640], 3, output_prune=True, dtype=dtype, max_batch=512)
    net.load_reference_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(1)
    x, y = 0.5 * torch.randn(300, 557, generator=g), 2.0 * torch.randn(300, 368, generator=g)      # |d| spans both Huber branches
    want = ref(x)
    with torch.no_grad():
        got = net(x.cuda()).cpu()
    assert (got[:, 60:75] == 0).all() and (got[:, 240:255] == 0).all() and (got[:, -8:] >= 0).all()
    assert (got - want).abs().max().item() <= tol * want.abs().max().item()
    crit = {"huber": torch.nn.HuberLoss(), "mse": torch.nn.MSELoss(), "mae": torch.nn.L1Loss()}[loss]
    l_ref = crit(want, y)
    l_ref.backward()
    # fused path: the engine's own loss (criterion selected at creation)
    from climsim_b200 import MLPEngine
    eng = MLPEngine(557, [(384, "relu", 0.0), (1024, "relu", 0.0), (640, "relu", 0.0), (368, "none", 0.0)], head_relu_from=360,
                    dtype=dtype, loss=loss, max_batch=512)
    eng.set_params_flat(net.flat.detach().cpu().numpy())
    mask = np.ones(368, np.float32)
    for s0 in (60, 120, 180, 240):
        mask[s0:s0 + 15] = 0
    eng.set_output_mask(mask)
    l_got = eng.train_step(x.cuda(), y.cuda()).item()
    assert abs(l_got - l_ref.item()) <= max(tol, 1e-5) * abs(l_ref.item())
    want_g = torch.cat([torch.cat([l[0].weight.grad.t().reshape(-1), l[0].bias.grad]) for l in ref.linears] +
                       [ref.final_linear.weight.grad.t().reshape(-1), ref.final_linear.bias.grad]).numpy()
    got_g = eng.get_grads_flat()
    tg = tol_g * (2.5 if (loss == "mae" and dtype == "bf16") else 1.0)       # sign(d) flips under bf16 rounding
    assert np.linalg.norm(got_g - want_g) / np.linalg.norm(want_g) <= tg


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("bf16", 3e-2)])
def test_online_inference_wrapper_matches_reference_golden(golden_dir, dtype, tol):
    """OnlineInferenceWrapper (prologue kernel with the exp transforms / pruning / clipping -> GEMM chain -> output epilogue with mask
    and 1/out_scale, one engine call) against the outputs of the REFERENCE's `NewModel` (v2_nn_wrapper.ipynb cell 5) wrapped around
    the reference's own `mlp.MLP` (tests/golden/online_mlp.npz)."""
    from climsim_b200.baseline_models import OnlineInferenceWrapper, OnlineMLP
    g = np.load(os.path.join(golden_dir, "online_mlp.npz"))
    hidden = [int(h) for h in g["hidden"]]
    net = OnlineMLP(557, 368, hidden, 3, output_prune=True, strato_lev_out=15, dtype=dtype, max_batch=256)
    net.load_reference_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")})
    wrap = OnlineInferenceWrapper(net, g["input_sub"], g["input_div"], g["out_scale"], g["lbd_qc"], g["lbd_qi"])
    x = torch.from_numpy(g["x_raw"]).cuda()
    x_before = x.clone()
    got = wrap(x).cpu().numpy()
    assert torch.equal(x, x_before)                              # unlike the reference module, the caller's buffer is left alone
    # compare in scaled space (the physical outputs span eight decades through 1 / out_scale)
    got_s, want_s = got * g["out_scale"], g["y_wrapped"] * g["out_scale"]
    assert np.abs(got_s - want_s).max() <= tol * np.abs(want_s).max()
    for a, b in ((60, 75), (120, 148), (180, 195), (240, 255), (300, 315)):
        assert (got[:, a:b] == 0).all()
    assert (got[:, -8:] >= 0).all() and np.isfinite(got).all()
    # the wrapped module gets its own configuration back: bare network on normalised inputs == the reference network
    with torch.no_grad():
        y_plain = net((torch.from_numpy(g["x_raw"]) * 0.1).cuda()).cpu().numpy()
    assert np.abs(y_plain - g["y_plain"]).max() <= tol * np.abs(g["y_plain"]).max()
    # and the wrapper can be used again afterwards
    got2 = wrap(x).cpu().numpy()
    np.testing.assert_array_equal(got, got2)


def test_input_transform_prologue_bit_exact(golden_dir):
    """The prologue kernel alone (fp32 engine, identity 557 -> 557 'network' is not available, so the first layer's input buffer is
    read back through a one-layer engine with an identity weight): pre-processed inputs equal the reference's bit for bit except for
    the exp() columns (libdevice expf against the host libm, <= 2 ulp of 1 - exp(.) before normalisation)."""
    from climsim_b200 import MLPEngine
    g = np.load(os.path.join(golden_dir, "online_mlp.npz"))
    eng = MLPEngine(557, [(557, "none", 0.0)], head_relu_from=-1, dtype="fp32", max_batch=64)
    flat = np.concatenate([np.eye(557, dtype=np.float32).reshape(-1), np.zeros(557, np.float32)])
    eng.set_params_flat(flat)
    lam = np.zeros(557, np.float32); lam[120:180] = g["lbd_qc"]; lam[180:240] = g["lbd_qi"]
    keep = np.ones(557, np.float32); keep[120:135] = 0; keep[180:195] = 0
    lo, hi = np.full(557, -np.inf, np.float32), np.full(557, np.inf, np.float32)
    lo[60:120], hi[60:120] = 0.0, 1.2
    eng.set_norm(inp_sub=g["input_sub"], inp_div=g["input_div"])
    eng.set_input_transform(lam, keep, lo, hi)
    got = eng.forward(torch.from_numpy(g["x_raw"]).cuda(), normalize_in=True).cpu().numpy()
    want = g["pre"]
    plain = np.ones(557, bool); plain[120:240] = False
    np.testing.assert_array_equal(got[:, plain], want[:, plain])
    np.testing.assert_allclose(got[:, ~plain], want[:, ~plain], rtol=0, atol=4e-7 / np.abs(g["input_div"][~plain]).min())
    eng.set_input_transform()                                     # back to the plain normalisation
    got0 = eng.forward(torch.from_numpy(g["x_raw"]).cuda(), normalize_in=True).cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (g["x_raw"] - g["input_sub"]) / g["input_div"]
    z[~np.isfinite(z)] = 0
    np.testing.assert_array_equal(got0, z.astype(np.float32))


def test_fused_gpu_metrics_match_reference_golden(golden_dir):
    """csb_eval_metrics (weighting + MAE/RMSE/R2/bias + grid mean in one pass) against the per-index metrics the REFERENCE's
    own data_utils produced (tests/golden/data_utils.npz)."""
    from test_data_utils_cpu import V1_IN, V1_OUT
    from climsim_b200.data_utils import data_utils
    g = np.load(os.path.join(golden_dir, "data_utils.npz"))
    ncol = int(g["ncol"])
    def split(vec, names, ls):
        out, off = {}, 0
        for n, l in zip(names, ls):
            out[n] = vec[off:off + l] if l > 1 else vec[off]
            off += l
        return out
    mean = split(g["inp_sub"], V1_IN, [60, 60, 1, 1, 1, 1])
    vmax, vmin = split(g["inp_div"], V1_IN, [60, 60, 1, 1, 1, 1]), split(np.zeros(124), V1_IN, [60, 60, 1, 1, 1, 1])
    mean["state_ps"], vmax["state_ps"], vmin["state_ps"] = g["ps_mean"], g["ps_max"], g["ps_min"]
    du = data_utils({"lev": np.arange(60), "ncol": np.arange(ncol), "area": g["area"], "hyai": g["hyai"], "hybi": g["hybi"], "P0": 1e5},
                    mean, vmax, vmin, split(g["out_scale"], V1_OUT, [60, 60] + [1] * 8))
    du.set_to_v1_vars()
    df = du.gpu_metrics(torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["target"]).cuda(), torch.from_numpy(g["x_norm"]).cuda())
    col = 0
    for v in V1_OUT:
        n = 60 if v in ("ptend_t", "ptend_q0001") else 1
        for m in ("MAE", "RMSE", "R2", "bias"):
            want = np.atleast_1d(g[f"{m}_{v}"])
            got = df[m].values[col:col + n]
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(want).max()), err_msg=f"{m} {v}")
        col += n


def test_gpu_crps_matches_reference_golden(golden_dir):
    """csb_eval_crps (a warp per (time, column, level) element, bitonic sort of the ensemble in registers) against the CRPS the
    REFERENCE's data_utils.calc_CRPS produced (tests/golden/data_utils.npz), and against the oracle on a 32-member ensemble."""
    from climsim_b200.data_utils import data_utils
    from oracle import data_utils_ref as R
    g = np.load(os.path.join(golden_dir, "data_utils.npz"))
    du = data_utils.__new__(data_utils)
    du.num_latlon = int(g["ncol"])
    got = du.calc_CRPS(torch.from_numpy(g["crps_samples"]).cuda(), torch.from_numpy(g["tw_ptend_t"]).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, g["crps"], rtol=1e-12, atol=1e-300)
    rng = np.random.default_rng(3)
    T, ncol = 37, int(g["ncol"])
    for shape_l in ((60,), ()):                                       # profile variable / scalar variable
        samples = rng.normal(size=(T, ncol) + shape_l + (32,))
        samples[0, 0] = samples[0, 0, ..., :1]                         # ties: an ensemble of identical members has zero spread
        target = rng.normal(size=(T, ncol) + shape_l)
        want = R.calc_crps(samples, target)
        got64 = du.calc_CRPS(torch.from_numpy(samples).cuda(), torch.from_numpy(target).cuda()).cpu().numpy()
        np.testing.assert_allclose(got64, want, rtol=1e-12)
        got32 = du.calc_CRPS(torch.from_numpy(samples).float().cuda(), torch.from_numpy(target).float().cuda()).cpu().numpy()
        np.testing.assert_allclose(got32, want, rtol=2e-6)
    with pytest.raises(NotImplementedError):
        du.calc_CRPS(torch.from_numpy(g["crps_samples"]).cuda(), torch.from_numpy(g["tw_ptend_t"]).cuda(), avg_grid=False)


def test_hsr_trainer_bf16_tracks_the_reference_losses(golden_dir):
    """The same run on the tensor-core engine (bf16 operands, fp32 accumulation / master weights / optimizer, LayerNorm parameters
    updated by the fused optimizer launch): the loss trajectory follows the reference's within bf16 rounding."""
    from climsim_b200.baseline_models import HSR
    g = np.load(os.path.join(golden_dir, "hsr_small.npz"))
    sd = {k[len("init::"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("init::")}
    net = HSR(124, 128, hidden_dims=32, layers=2, dtype="bf16", max_batch=64)
    net.load_reference_state_dict(sd)
    batches = [{"x": torch.from_numpy(g[f"x{i}"]), "y": torch.from_numpy(g[f"y{i}"])} for i in range(2)]
    losses = net.trainer(batches, epochs=3, save=os.devnull, plot=False, lr=1e-3, gamma=0.022)
    np.testing.assert_allclose(losses, g["trainer_losses"], rtol=3e-2)
    got = net.reference_state_dict()
    for k in got:
        want = g["final::" + k]
        err = np.linalg.norm(got[k].cpu().numpy() - want) / max(np.linalg.norm(want), 1e-12)
        assert err <= 5e-2, (k, err)                                    # six Adam steps of lr 1e-3 on bf16-rounded gradients


def test_hsr_step_loss_clip_blocks_the_gradient():
    """torch.clip(loss, -1e5, 1e5).backward(): outside the interval the gradient is zero -- parameters then move by the weight decay
    alone (hsr.py:138).  Targets of 1e4 make the MSE 1e8."""
    from climsim_b200.baseline_models import HSR
    net = HSR(124, 128, hidden_dims=64, layers=1, dtype="fp32", max_batch=32)
    x, y = 0.3 * torch.randn(32, 124), torch.full((32, 128), 1.0e4)
    before = net.mean.flat.detach().clone()
    losses = net.trainer([{"x": x, "y": y}], epochs=3, save=os.devnull, plot=False, optimizer="sgd", lr=1e-2, gamma=0.0)   # alpha = 0
    assert losses[0] > 1e5
    assert torch.equal(net.mean.flat.detach(), before)                  # SGD without decay and a zero gradient: nothing moves


@pytest.mark.parametrize("B,F", [(1, 128), (777, 128), (4096, 368), (33, 10)])
def test_batch_metrics_kernel_against_torch(B, F):
    """csb_batch_metrics = the sufficient statistics of Keras' metrics=['mse','mae','accuracy'] (hpo_baseline_v1.py:127-129)."""
    from climsim_b200 import MLPEngine
    eng = MLPEngine(8, [(8, "none", 0.0)], dtype="fp32", max_batch=8)
    g = torch.Generator().manual_seed(B)
    p, y = torch.randn(B, F, generator=g).cuda(), torch.randn(B, F, generator=g).cuda()
    p[0, 3] = p[0].max() + 1; p[0, 7] = p[0, 3]                       # a tie: the first maximum wins (tf.argmax)
    for _ in range(2):                                                  # the scratch ticket re-arms itself
        got = eng.batch_metrics(p, y).cpu().numpy()
    d = (p - y).double()
    want = [float((d * d).sum()), float(d.abs().sum()), float((p.argmax(1) == y.argmax(1)).sum()), B * F, B]
    np.testing.assert_allclose(got, want, rtol=1e-6)
    assert int(p[0].argmax()) == 3


def test_hsr_mlp_dropout_module_training_mode():
    """hsr.MLP with Dropout(p) (hsr.py:20-25: Linear -> LayerNorm -> Dropout -> ReLU) as a torch module: in ``train()`` mode the
    autograd forward drops (the engine's masks, replayed in the torch restatement through ``masks=``), ``eval()`` does not; outputs
    and every gradient tensor of the fp32 engine against torch autograd."""
    from climsim_b200.baseline_models import HSRMLP
    from oracle import models as M
    torch.manual_seed(5)
    hidden, layers, B, rate = 200, 2, 97, 0.25
    ref = M.HSRMLPRef(124, 128, hidden, layers)
    with torch.no_grad():
        for i in range(layers):
            ln = getattr(ref, f"linear{i}")[1]
            ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.3, 0.3)
    net = HSRMLP(124, 128, hidden_dims=hidden, layers=layers, dropout=rate, dtype="fp32", max_batch=128)
    net.load_reference_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(6)
    x, y = 0.5 * torch.randn(B, 124, generator=g), 0.3 * torch.randn(B, 128, generator=g)
    net.train()
    got = net(x.cuda())
    ((got - y.cuda()) ** 2).mean().backward()
    masks = [net.engine.dropout_mask(l, B).cpu() for l in range(layers)]
    assert all(abs((m == 0).float().mean().item() - rate) < 0.03 for m in masks)
    want = ref(x, masks=masks)
    ((want - y) ** 2).mean().backward()
    assert (got.detach().cpu() - want.detach()).abs().max().item() <= 2e-5 * want.abs().max().item()
    want_g = []
    for i in range(layers):
        seq = getattr(ref, f"linear{i}")
        want_g += [seq[0].weight.grad.t().reshape(-1), seq[0].bias.grad, seq[1].weight.grad, seq[1].bias.grad]
    want_g += [ref.final_linear.weight.grad.t().reshape(-1), ref.final_linear.bias.grad]
    off, got_g = 0, net.flat.grad.cpu()
    for i, w in enumerate(want_g):
        part = got_g[off:off + w.numel()]
        off += w.numel()
        assert (part - w).abs().max().item() <= 5e-5 * max(w.abs().max().item(), 1e-12), (i, (part - w).abs().max().item())
    net.eval()
    with torch.no_grad():
        assert (net(x.cuda()).cpu() - ref(x)).abs().max().item() <= 2e-5 * want.abs().max().item()


def test_layernorm_mlp_tf32_mode():
    """The HSR network (Linear -> LayerNorm -> ReLU) in the CSB_TF32 mode: GEMMs on the tensor cores with kind::tf32, the fp32 LayerNorm
    kernels in between.  Outputs 2e-3 of the largest, gradients 6e-2 relative L2 against fp32 torch (ReLU kinks: see
    tests/test_mlp_gpu.py::test_train_step_tf32_mode)."""
    from climsim_b200.baseline_models import HSRMLP
    from oracle import models as M
    torch.manual_seed(3)
    hidden, layers, B = 512, 2, 333
    ref = M.HSRMLPRef(124, 128, hidden, layers)
    with torch.no_grad():
        for i in range(layers):
            ln = getattr(ref, f"linear{i}")[1]
            ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.3, 0.3)
    net = HSRMLP(124, 128, hidden_dims=hidden, layers=layers, dtype="tf32", max_batch=512)
    net.load_reference_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(4)
    x, y = 0.5 * torch.randn(B, 124, generator=g), 0.3 * torch.randn(B, 128, generator=g)
    want = ref(x)
    ((want - y) ** 2).mean().backward()
    got = net(x.cuda())
    ((got - y.cuda()) ** 2).mean().backward()
    assert (got.detach().cpu() - want.detach()).abs().max().item() <= 2e-3 * want.abs().max().item()
    want_g = []
    for i in range(layers):
        seq = getattr(ref, f"linear{i}")
        want_g += [seq[0].weight.grad.t().reshape(-1), seq[0].bias.grad, seq[1].weight.grad, seq[1].bias.grad]
    want_g += [ref.final_linear.weight.grad.t().reshape(-1), ref.final_linear.bias.grad]
    off, got_g = 0, net.flat.grad.cpu()
    for i, w in enumerate(want_g):
        part = got_g[off:off + w.numel()]
        off += w.numel()
        err = (part - w).norm().item() / max(w.norm().item(), 1e-12)
        assert err <= 6e-2, (i, err)
