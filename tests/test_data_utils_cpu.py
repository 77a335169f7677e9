"""The ``climsim_b200.data_utils`` mirror against golden vectors produced by the REFERENCE's own class
(tests/golden/make_golden.py -> data_utils.npz): same method names, same results."""
import os

import numpy as np
import pytest

from climsim_b200.data_utils import data_utils

V1_OUT = ["ptend_t", "ptend_q0001", "cam_out_NETSW", "cam_out_FLWDS", "cam_out_PRECSC", "cam_out_PRECC", "cam_out_SOLS",
          "cam_out_SOLL", "cam_out_SOLSD", "cam_out_SOLLD"]
V1_IN = ["state_t", "state_q0001", "state_ps", "pbuf_SOLIN", "pbuf_LHFLX", "pbuf_SHFLX"]


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "data_utils.npz"))


@pytest.fixture(scope="module")
def du(g):
    ncol = int(g["ncol"])
    lens = [60, 60, 1, 1, 1, 1]
    def split(vec, names, ls):
        out, off = {}, 0
        for n, l in zip(names, ls):
            out[n] = vec[off:off + l] if l > 1 else vec[off]
            off += l
        return out
    sub, div, scale = g["inp_sub"], g["inp_div"], g["out_scale"]
    mean = split(sub, V1_IN, lens)
    vmin = split(np.zeros(124), V1_IN, lens)
    vmax = split(div, V1_IN, lens)                         # max - min == div with min = 0
    # state_ps needs its true (mean, max, min): the golden file carries them
    mean["state_ps"], vmax["state_ps"], vmin["state_ps"] = g["ps_mean"], g["ps_max"], g["ps_min"]
    oscale = split(scale, V1_OUT, [60, 60] + [1] * 8)
    grid = {"lev": np.arange(60), "ncol": np.arange(ncol), "area": g["area"], "hyai": g["hyai"], "hybi": g["hybi"], "P0": 1e5}
    d = data_utils(grid_info=grid, input_mean=mean, input_max=vmax, input_min=vmin, output_scale=oscale)
    d.set_to_v1_vars()
    return d


def test_save_norm(du, g):
    sub, div, scale = du.save_norm()
    np.testing.assert_array_equal(sub, g["inp_sub"])
    np.testing.assert_allclose(div, g["inp_div"], rtol=1e-15)
    np.testing.assert_array_equal(scale, g["out_scale"])


def test_normalize_input_host(du, g):
    # inp_div from the fixture (max - 0) equals the golden one except state_ps; compare all other columns
    xn = du.normalize_input(g["x_raw"])
    cols = [c for c in range(124) if c != 120]
    np.testing.assert_array_equal(xn[:, cols], g["x_renorm"][:, cols])
    assert np.all(xn[:, 60] == 0)


def test_pressure_grid_weighting_metrics(du, g):
    du.input_val, du.target_val = g["x_norm"], g["target"]
    du.set_pressure_grid("val")
    np.testing.assert_array_equal(du.dp_val, g["dp_val"])
    du.model_names, du.metrics_names = ["m"], ["MAE", "RMSE", "R2", "bias"]
    du.preds_val = {"m": g["pred"]}
    du.reweight_target("val")
    du.reweight_preds("val")
    for v in V1_OUT:
        np.testing.assert_array_equal(du.target_weighted_val[v], g["tw_" + v])
        np.testing.assert_array_equal(du.preds_weighted_val["m"][v], g["pw_" + v])
    np.testing.assert_array_equal(du.output_weighting(g["target"], "val", just_weights=True), g["just_weights"])
    du.create_metrics_df("val")
    df = du.metrics_var_val["m"]
    for v in V1_OUT:
        for m in du.metrics_names:
            assert float(df.loc[v, m]) == pytest.approx(float(np.mean(g[f"{m}_{v}"])), rel=1e-12)
    assert du.metrics_idx_val["m"].shape == (128, 4)
    np.testing.assert_allclose(du.calc_CRPS(g["crps_samples"], g["tw_ptend_t"]), g["crps"], rtol=1e-13)


def test_cnn_reshapes_host(g):
    np.testing.assert_array_equal(data_utils.reshape_input_for_cnn(g["x_norm"]), g["cnn_in"])
    np.testing.assert_array_equal(data_utils.reshape_target_for_cnn(g["target"]), g["cnn_tgt"])
    np.testing.assert_array_equal(data_utils.reshape_target_from_cnn(g["cnn_pred"]), g["cnn_pred_flat"])


def test_assert_messages_match_reference(du):
    with pytest.raises(AssertionError, match="Provided data_split is not valid"):
        du.set_pressure_grid("nope")


def test_reshape_npy(du):
    a = np.arange(du.num_latlon * 3 * 60, dtype=np.float32).reshape(du.num_latlon * 3, 60)
    r = du.reshape_npy(a, 60)
    assert r.shape == (3, du.num_latlon, 60) and r[1, 2, 5] == a[du.num_latlon + 2, 5]
