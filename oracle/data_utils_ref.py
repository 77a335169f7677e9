"""CPU oracle: NumPy restatement of the hot-path arithmetic of ``climsim_utils/data_utils.py``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED: every function here is checked against golden vectors
produced by the reference's own class (tests/golden/make_golden.py -> tests/golden/data_utils.npz) in
tests/test_oracle_pinning.py.  V1 variable set only (124 inputs / 128 targets), like the reference's
``output_weighting`` with ``full_vars=False``.

All functions are vectorised re-derivations of the reference's per-variable code; line cites point at the
statement being restated.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

# V1 feature layout (data_utils.py:172-188, 392-..., 558-568)
V1_INPUTS = ["state_t", "state_q0001", "state_ps", "pbuf_SOLIN", "pbuf_LHFLX", "pbuf_SHFLX"]
V1_OUTPUTS = ["ptend_t", "ptend_q0001", "cam_out_NETSW", "cam_out_FLWDS", "cam_out_PRECSC", "cam_out_PRECC",
              "cam_out_SOLS", "cam_out_SOLL", "cam_out_SOLSD", "cam_out_SOLLD"]
VAR_LENS = {"state_t": 60, "state_q0001": 60, "state_ps": 1, "pbuf_SOLIN": 1, "pbuf_LHFLX": 1, "pbuf_SHFLX": 1,
            "ptend_t": 60, "ptend_q0001": 60, "cam_out_NETSW": 1, "cam_out_FLWDS": 1, "cam_out_PRECSC": 1,
            "cam_out_PRECC": 1, "cam_out_SOLS": 1, "cam_out_SOLL": 1, "cam_out_SOLSD": 1, "cam_out_SOLLD": 1}
PS_INDEX = 120
GRAV, CP, LV, RHO_H2O = 9.80616, 1.00464e3, 2.501e6, 1.0e3          # data_utils.py:128-138


def flatten_vars(per_var: Dict[str, np.ndarray], names: Sequence[str]) -> np.ndarray:
    """Concatenate per-variable vectors in list order (data_utils.py:815-820 ``to_stacked_array``; :954-988)."""
    return np.concatenate([np.atleast_1d(np.asarray(per_var[v], dtype=np.float64)).reshape(-1) for v in names])


def save_norm(input_mean, input_max, input_min, output_scale) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """data_utils.py:954-988: inp_sub = mean, inp_div = max - min, out_scale, flattened to 124 / 124 / 128."""
    sub = flatten_vars(input_mean, V1_INPUTS)
    div = flatten_vars(input_max, V1_INPUTS) - flatten_vars(input_min, V1_INPUTS)
    return sub, div, flatten_vars(output_scale, V1_OUTPUTS)


def normalize_input(x_raw: np.ndarray, inp_sub: np.ndarray, inp_div: np.ndarray) -> np.ndarray:
    """data_utils.py:806-809 then :894-897,:906 -- (x - mean)/(max - min) in fp64, inf -> 0, nan -> 0, cast fp32."""
    with np.errstate(divide="ignore", invalid="ignore"):
        xn = (np.asarray(x_raw, dtype=np.float64) - inp_sub) / inp_div
    xn[np.isinf(xn)] = 0
    xn[np.isnan(xn)] = 0
    return np.float32(xn)


def scale_target(y_raw: np.ndarray, out_scale: np.ndarray) -> np.ndarray:
    """data_utils.py:809 (``ds_target*self.output_scale``) and :917 (fp32 cast)."""
    return np.float32(np.asarray(y_raw, dtype=np.float64) * out_scale)


def pressure_thickness(x_norm: np.ndarray, ps_mean: float, ps_max: float, ps_min: float, hyai: np.ndarray,
                       hybi: np.ndarray, p0: float, num_latlon: int, normalize: bool = True) -> np.ndarray:
    """data_utils.py:1037-1086 ``set_pressure_grid``: de-normalise surface pressure, build the 61 interface
    pressures P0*hyai + hybi*ps and difference them.  Returns dp with shape (T, num_latlon, 60)."""
    ps = x_norm[:, PS_INDEX].astype(np.float64)   # fp32 column x 0-d fp64 ``.values`` promotes to fp64 in the reference
    if normalize:
        ps = ps * (ps_max - ps_min) + ps_mean
    ps = np.reshape(ps, (-1, num_latlon))
    p_int = (p0 * hyai)[:, None, None] + hybi[:, None, None] * ps[None, :, :]
    return (p_int[1:61] - p_int[0:60]).transpose((1, 2, 0))


def energy_conv() -> np.ndarray:
    """Length-128 unit-conversion vector (data_utils.py:480-494): cp for dT/dt, lv for dq/dt,
    lv*rho_h2o for PRECSC/PRECC, 1 for the radiative fluxes."""
    conv = np.ones(128)
    conv[0:60] = CP
    conv[60:120] = LV
    conv[122] = LV * RHO_H2O
    conv[123] = LV * RHO_H2O
    return conv


def output_weights(dp: np.ndarray, out_scale: np.ndarray, area_wgt: np.ndarray, normalize: bool = True) -> np.ndarray:
    """The (N,128) multiplier ``output_weighting(..., just_weights=True)`` returns (data_utils.py:1112-1344):
    1/out_scale  *  dp/g (profiles only)  *  area_wgt[col]  *  energy conversion."""
    t, ncol, _ = dp.shape
    w = np.ones((t, ncol, 128))
    if normalize:
        w = w / out_scale[None, None, :]
    w[:, :, 0:60] = w[:, :, 0:60] * dp / GRAV
    w[:, :, 60:120] = w[:, :, 60:120] * dp / GRAV
    w = w * area_wgt[None, :, None]
    w = w * energy_conv()[None, None, :]
    return w.reshape(t * ncol, 128)


def output_weighting(output: np.ndarray, dp: np.ndarray, out_scale: np.ndarray, area_wgt: np.ndarray,
                     normalize: bool = True) -> Dict[str, np.ndarray]:
    """data_utils.py:1112-1362 with ``just_weights=False``: dict var -> (T,384,60) profiles / (T,384) scalars.
    The order of the four multiplications follows the reference so that fp64 rounding matches bit for bit."""
    t, ncol, _ = dp.shape
    assert output.shape[0] == t * ncol
    conv = energy_conv()
    res = {}
    col = 0
    for var in V1_OUTPUTS:
        n = VAR_LENS[var]
        if n == 60:
            a = output[:, col:col + 60].reshape((t, ncol, 60))
            if normalize:
                a = a / out_scale[col:col + 60][None, None, :]
            a = a * dp / GRAV
            a = a * area_wgt[None, :, None]
        else:
            a = output[:, col].reshape((t, ncol))
            if normalize:
                a = a / out_scale[col]
            a = a * area_wgt[None, :]
        res[var] = a * conv[col]
        col += n
    return res


# -- metrics (data_utils.py:1432-1524) ------------------------------------------------------------------------

def calc_mae(pred, target, avg_grid=True):
    m = np.abs(pred - target).mean(axis=0)
    return m.mean(axis=0) if avg_grid else m


def calc_rmse(pred, target, avg_grid=True):
    m = np.sqrt(((pred - target) ** 2).mean(axis=0))
    return m.mean(axis=0) if avg_grid else m


def calc_r2(pred, target, avg_grid=True):
    ss_res = ((pred - target) ** 2).sum(axis=0)
    ss_tot = ((target - target.mean(axis=0)[None, ...]) ** 2).sum(axis=0)
    m = 1 - ss_res / ss_tot
    return m.mean(axis=0) if avg_grid else m


def calc_bias(pred, target, avg_grid=True):
    m = pred.mean(axis=0) - target.mean(axis=0)
    return m.mean(axis=0) if avg_grid else m


def calc_crps(samplepreds, target, avg_grid=True):
    """Sorted-sample CRPS identity, data_utils.py:1499-1524."""
    n = samplepreds.shape[-1]
    mae = np.mean(np.abs(samplepreds - target[..., None]), axis=(0, -1))
    s = np.sort(samplepreds, axis=-1)
    diff = s[..., 1:] - s[..., :-1]
    count = np.arange(1, n) * np.arange(n - 1, 0, -1)
    spread = (diff * count).sum(axis=-1).mean(axis=0)
    m = mae - spread / (n * (n - 1))
    return m.mean(axis=0) if avg_grid else m


# -- CNN reshapes (data_utils.py:1693-1760) -------------------------------------------------------------------

def reshape_input_for_cnn(x: np.ndarray) -> np.ndarray:
    """(N,124) -> (N,60,6): two profile channels + four scalars broadcast over the 60 levels (:1693-1712)."""
    n = x.shape[0]
    out = np.empty((n, 60, 6), dtype=x.dtype)
    out[:, :, 0] = x[:, 0:60]
    out[:, :, 1] = x[:, 60:120]
    out[:, :, 2:6] = x[:, None, 120:124]
    return out


def reshape_target_for_cnn(y: np.ndarray) -> np.ndarray:
    """(N,128) -> (N,60,10) (:1715-1738)."""
    n = y.shape[0]
    out = np.empty((n, 60, 10), dtype=y.dtype)
    out[:, :, 0] = y[:, 0:60]
    out[:, :, 1] = y[:, 60:120]
    out[:, :, 2:10] = y[:, None, 120:128]
    return out


def reshape_target_from_cnn(p: np.ndarray) -> np.ndarray:
    """(N,60,10) -> (N,128): channels 0,1 as profiles, channels 2..9 averaged over the levels (:1741-1760)."""
    # one strided (N,60) slice per channel, like the reference: NumPy's summation order (hence the fp32 rounding of
    # the mean) depends on the memory layout of the slice being reduced
    scal = [np.mean(p[:, :, c], axis=1)[:, None] for c in range(2, 10)]
    return np.concatenate([p[:, :, 0], p[:, :, 1]] + scal, axis=1)
