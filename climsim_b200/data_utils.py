"""``data_utils``: the part of ``climsim_utils.data_utils.data_utils`` that sits on the emulator hot path, with the
same names, arguments and error behaviour (bare ``assert`` with the reference's messages), for the V1 variable set.

Differences, by design:
  * no xarray / netCDF4 / tensorflow / h5py import: ``grid_info`` and the normalisation objects may be plain dicts of
    NumPy arrays (anything with ``.values`` such as an ``xarray.Dataset`` works too);
  * reading raw netCDF files (``load_ncdata_with_generator``, ``save_as_npy``), plotting and the v2/v4 variable sets are
    out of scope -- the ``.npy`` arrays those produce are the input contract here;
  * ``normalize`` / the CNN reshapes accept CUDA tensors and then run in the engine's kernels (``csb_normalize``,
    ``csb_reshape_*``); NumPy inputs are handled on the host exactly as the reference does.

Reference: climsim_utils/data_utils.py (class ``data_utils``, :45).  Line cites below are into that file.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _val(a):
    """NumPy view of an xarray-like object / array / scalar."""
    return np.asarray(a.values if hasattr(a, "values") else a, dtype=np.float64)


class data_utils:
    def __init__(self, grid_info, input_mean, input_max, input_min, output_scale, ml_backend="pytorch", normalize=True,
                 input_abbrev="mli", output_abbrev="mlo", save_h5=False, save_npy=True):
        self.input_abbrev, self.output_abbrev = input_abbrev, output_abbrev
        self.grid_info = grid_info
        self.input_mean, self.input_max, self.input_min, self.output_scale = input_mean, input_max, input_min, output_scale
        self.normalize = normalize
        self.ml_backend = ml_backend
        self.num_levels = len(_val(grid_info["lev"])) if "lev" in grid_info else 60
        self.num_latlon = len(_val(grid_info["ncol"])) if "ncol" in grid_info else len(_val(grid_info["area"]))
        area = _val(grid_info["area"])
        self.area_wgt = area / area.mean()                                    # :70-71
        self.hyai, self.hybi = _val(grid_info["hyai"]), _val(grid_info["hybi"])
        self.p0 = float(_val(grid_info["P0"])) if "P0" in grid_info else 1e5
        self.grav, self.cp, self.lv, self.rho_h20 = 9.80616, 1.00464e3, 2.501e6, 1.0e3      # :128-138
        self.v1_inputs = ["state_t", "state_q0001", "state_ps", "pbuf_SOLIN", "pbuf_LHFLX", "pbuf_SHFLX"]
        self.v1_outputs = ["ptend_t", "ptend_q0001", "cam_out_NETSW", "cam_out_FLWDS", "cam_out_PRECSC", "cam_out_PRECC",
                           "cam_out_SOLS", "cam_out_SOLL", "cam_out_SOLSD", "cam_out_SOLLD"]
        self.var_lens = {v: 60 for v in ("state_t", "state_q0001", "ptend_t", "ptend_q0001")}
        self.var_lens.update({v: 1 for v in self.v1_inputs[2:] + self.v1_outputs[2:]})
        self.target_energy_conv = {"ptend_t": self.cp, "ptend_q0001": self.lv, "cam_out_NETSW": 1.0, "cam_out_FLWDS": 1.0,
                                   "cam_out_PRECSC": self.lv * self.rho_h20, "cam_out_PRECC": self.lv * self.rho_h20,
                                   "cam_out_SOLS": 1.0, "cam_out_SOLL": 1.0, "cam_out_SOLSD": 1.0, "cam_out_SOLLD": 1.0}   # :480-494
        self.input_vars, self.target_vars = [], []
        self.input_feature_len = self.target_feature_len = self.ps_index = None
        self.full_vars = False
        for split in ("train", "val", "scoring", "test"):
            setattr(self, f"input_{split}", None)
            setattr(self, f"target_{split}", None)
            setattr(self, f"preds_{split}", None)
            setattr(self, f"samplepreds_{split}", None)
            setattr(self, f"target_weighted_{split}", {})
            setattr(self, f"preds_weighted_{split}", {})
            setattr(self, f"metrics_idx_{split}", {})
            setattr(self, f"metrics_var_{split}", {})
            setattr(self, f"dp_{split}", None)
            setattr(self, f"pressure_grid_{split}", None)
        self.model_names, self.metrics_names = [], []
        self.metrics_dict = {"MAE": self.calc_MAE, "RMSE": self.calc_RMSE, "R2": self.calc_R2, "CRPS": self.calc_CRPS,
                             "bias": self.calc_bias}
        self.num_CRPS = 32

    # ------------------------------------------------------------------------------------------------ variable sets
    def set_to_v1_vars(self):
        """:558-568"""
        self.input_vars, self.target_vars = self.v1_inputs, self.v1_outputs
        self.ps_index, self.input_feature_len, self.target_feature_len, self.full_vars = 120, 124, 128, False

    # ------------------------------------------------------------------------------------------------ normalisation
    def save_norm(self, save_path="", write=False):
        """:954-988 -- (inp_sub, inp_div, out_scale) = (mean, max - min, output_scale) flattened to 124 / 124 / 128."""
        cat = lambda d, names: np.concatenate([np.atleast_1d(_val(d[v])).reshape(-1) for v in names])
        input_sub = cat(self.input_mean, self.input_vars)
        input_div = cat(self.input_max, self.input_vars) - cat(self.input_min, self.input_vars)
        out_scale = cat(self.output_scale, self.target_vars)
        if write:
            fmt = "%.6e"
            np.savetxt(save_path + "/inp_sub.txt", input_sub.reshape(1, -1), fmt=fmt, delimiter=",")
            np.savetxt(save_path + "/inp_div.txt", input_div.reshape(1, -1), fmt=fmt, delimiter=",")
            np.savetxt(save_path + "/out_scale.txt", out_scale.reshape(1, -1), fmt=fmt, delimiter=",")
        return input_sub, input_div, out_scale

    def normalize_input(self, x_raw):
        """The generator's ``(x - mean)/(max - min)`` followed by ``save_as_npy``'s inf/nan -> 0 and fp32 cast
        (:806-809, :894-897, :906).  CUDA tensors run in ``csb_normalize``."""
        sub, div, _ = self.save_norm()
        if torch is not None and isinstance(x_raw, torch.Tensor) and x_raw.is_cuda:
            from . import _lib
            lib = _lib.load()
            x = x_raw.to(torch.float32).contiguous()
            out = torch.empty_like(x)
            s = torch.from_numpy(sub.astype(np.float32)).cuda()
            d = torch.from_numpy(div.astype(np.float32)).cuda()
            _lib.check(lib.csb_normalize(x.data_ptr(), s.data_ptr(), d.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                                         _lib.current_stream_ptr()), "csb_normalize")
            return out
        with np.errstate(divide="ignore", invalid="ignore"):
            xn = (np.asarray(x_raw, dtype=np.float64) - sub) / div
        xn[np.isinf(xn)] = 0
        xn[np.isnan(xn)] = 0
        return np.float32(xn)

    def reshape_npy(self, var_arr, var_arr_dim):
        """:946-952 -- (num_samples, dim) -> (timestep, lat/lon column, dim)."""
        return var_arr.reshape((int(var_arr.shape[0] / self.num_latlon), self.num_latlon, var_arr_dim))

    @staticmethod
    def load_npy_file(load_path=""):
        """:1019-1026"""
        with open(load_path, "rb") as f:
            return np.load(f)

    # ------------------------------------------------------------------------------------------------ pressure grid
    def set_pressure_grid(self, data_split):
        """:1037-1086"""
        assert data_split in ["train", "val", "scoring", "test"], \
            "Provided data_split is not valid. Available options are train, val, scoring, and test."
        inp = getattr(self, f"input_{data_split}")
        assert inp is not None
        state_ps = np.asarray(inp)[:, self.ps_index]
        if self.normalize:
            state_ps = state_ps * (_val(self.input_max["state_ps"]) - _val(self.input_min["state_ps"])) + _val(self.input_mean["state_ps"])
        state_ps = np.reshape(state_ps, (-1, self.num_latlon))
        grid = (self.p0 * self.hyai)[:, None, None] + self.hybi[:, None, None] * state_ps[None, :, :]
        setattr(self, f"pressure_grid_{data_split}", grid)
        setattr(self, f"dp_{data_split}", (grid[1:61] - grid[0:60]).transpose((1, 2, 0)))

    # ------------------------------------------------------------------------------------------------ weighting
    def output_weighting(self, output, data_split, just_weights=False):
        """:1112-1362 (V1): undo the output scaling, weight profiles by dp/g, weight by area, convert to energy units.
        Returns a dict var -> (T, ncol, 60) / (T, ncol), or the (N,128) multiplier array with ``just_weights``."""
        assert data_split in ["train", "val", "scoring", "test"], \
            "Provided data_split is not valid. Available options are train, val, scoring, and test."
        dp = getattr(self, f"dp_{data_split}")
        assert dp is not None
        output = np.asarray(output)
        num_samples = output.shape[0]
        t = int(num_samples / self.num_latlon)
        src = np.ones(output.shape) if just_weights else output
        res, col = {}, 0
        for var in self.target_vars:
            n = self.var_lens[var]
            scale = np.atleast_1d(_val(self.output_scale[var]))
            if n == 60:
                a = src[:, col:col + 60].reshape((t, self.num_latlon, 60))
                if self.normalize:
                    a = a / scale[None, None, :]
                a = a * dp / self.grav
                a = a * self.area_wgt[None, :, None]
            else:
                a = src[:, col].reshape((t, self.num_latlon))
                if self.normalize:
                    a = a / scale[0]
                a = a * self.area_wgt[None, :]
            res[var] = a * self.target_energy_conv[var]
            col += n
        if just_weights:
            return np.concatenate([res[v].reshape((num_samples, self.var_lens[v])) for v in self.target_vars], axis=1)
        return res

    def reweight_target(self, data_split):
        """:1364-1380"""
        assert data_split in ["train", "val", "scoring", "test"], \
            "Provided data_split is not valid. Available options are train, val, scoring, and test."
        target = getattr(self, f"target_{data_split}")
        assert target is not None
        setattr(self, f"target_weighted_{data_split}", self.output_weighting(target, data_split))

    def reweight_preds(self, data_split):
        """:1382-1405"""
        assert data_split in ["train", "val", "scoring", "test"], \
            "Provided data_split is not valid. Available options are train, val, scoring, and test."
        assert self.model_names is not None
        preds = getattr(self, f"preds_{data_split}")
        assert preds is not None
        weighted = getattr(self, f"preds_weighted_{data_split}")
        for name in self.model_names:
            weighted[name] = self.output_weighting(preds[name], data_split)

    # ------------------------------------------------------------------------------------------------ metrics (:1432-1524)
    def calc_MAE(self, pred, target, avg_grid=True):
        assert pred.shape[1] == self.num_latlon
        assert pred.shape == target.shape
        m = np.abs(pred - target).mean(axis=0)
        return m.mean(axis=0) if avg_grid else m

    def calc_RMSE(self, pred, target, avg_grid=True):
        assert pred.shape[1] == self.num_latlon
        assert pred.shape == target.shape
        m = np.sqrt(((pred - target) ** 2).mean(axis=0))
        return m.mean(axis=0) if avg_grid else m

    def calc_R2(self, pred, target, avg_grid=True):
        assert pred.shape[1] == self.num_latlon
        assert pred.shape == target.shape
        ss_res = ((pred - target) ** 2).sum(axis=0)
        ss_tot = ((target - target.mean(axis=0)[np.newaxis, ...]) ** 2).sum(axis=0)
        m = 1 - ss_res / ss_tot
        return m.mean(axis=0) if avg_grid else m

    def calc_bias(self, pred, target, avg_grid=True):
        assert pred.shape[1] == self.num_latlon
        assert pred.shape == target.shape
        m = pred.mean(axis=0) - target.mean(axis=0)
        return m.mean(axis=0) if avg_grid else m

    def calc_CRPS(self, samplepreds, target, avg_grid=True):
        assert samplepreds.shape[1] == self.num_latlon
        assert len(samplepreds.shape) == len(target.shape) + 1
        assert len(samplepreds.shape) == 3 or len(samplepreds.shape) == 4
        if torch is not None and isinstance(samplepreds, torch.Tensor) and samplepreds.is_cuda:
            return self._gpu_crps(samplepreds, target, avg_grid)
        n = samplepreds.shape[-1]
        mae = np.mean(np.abs(samplepreds - target[..., np.newaxis]), axis=(0, -1))
        s = np.sort(samplepreds, axis=-1)
        diff = s[..., 1:] - s[..., :-1]
        count = np.arange(1, n) * np.arange(n - 1, 0, -1)
        spread = (diff * count).sum(axis=-1).mean(axis=0)
        m = mae - spread / (n * (n - 1))
        return m.mean(axis=0) if avg_grid else m

    @staticmethod
    def _gpu_crps(samplepreds, target, avg_grid):
        """``csb_eval_crps``: CUDA tensors (T, ncol, [L,] S) / (T, ncol[, L]), fp32 or fp64; returns a CUDA fp64 tensor of length L (or 1)."""
        from . import _lib
        lib = _lib.load()
        if not avg_grid:
            raise NotImplementedError("the device CRPS returns the grid average (avg_grid=True); pass NumPy arrays for the per-column field")
        f64 = samplepreds.dtype == torch.float64
        dt = torch.float64 if f64 else torch.float32
        s, t = samplepreds.to(dt).contiguous(), target.to(samplepreds.device, dt).contiguous()
        S = int(s.shape[-1])
        L = int(s.shape[2]) if s.dim() == 4 else 1
        n_tc = int(s.shape[0] * s.shape[1])
        out = torch.empty(L, dtype=torch.float64, device=s.device)
        scratch = torch.empty(L * 64, dtype=torch.float64, device=s.device)
        _lib.check(lib.csb_eval_crps(s.data_ptr(), t.data_ptr(), 1 if f64 else 0, n_tc, L, S, out.data_ptr(), scratch.data_ptr(),
                                     _lib.current_stream_ptr()), "csb_eval_crps")
        return out if s.dim() == 4 else out[0]

    def create_metrics_df(self, data_split):
        """:1526-1621 -- per-variable and per-output-index metric tables for every model in ``model_names``."""
        assert data_split in ["train", "val", "scoring", "test"], \
            "Provided data_split is not valid. Available options are train, val, scoring, and test."
        assert len(self.model_names) != 0
        assert len(self.metrics_names) != 0
        assert len(self.target_vars) != 0
        assert self.target_feature_len is not None
        preds_w, target_w = getattr(self, f"preds_weighted_{data_split}"), getattr(self, f"target_weighted_{data_split}")
        assert len(preds_w) != 0
        assert len(target_w) != 0
        for model_name in self.model_names:
            df_var = pd.DataFrame(columns=self.metrics_names, index=self.target_vars)
            df_var.index.name = "variable"
            df_idx = pd.DataFrame(columns=self.metrics_names, index=range(self.target_feature_len))
            df_idx.index.name = "output_idx"
            for metric_name in self.metrics_names:
                current_idx = 0
                for target_var in self.target_vars:
                    metric = self.metrics_dict[metric_name](preds_w[model_name][target_var], target_w[target_var])
                    df_var.loc[target_var, metric_name] = np.mean(metric)
                    df_idx.loc[current_idx:current_idx + self.var_lens[target_var] - 1, metric_name] = np.atleast_1d(metric)
                    current_idx += self.var_lens[target_var]
            getattr(self, f"metrics_var_{data_split}")[model_name] = df_var
            getattr(self, f"metrics_idx_{data_split}")[model_name] = df_idx

    def gpu_metrics(self, preds, targets, inputs):
        """Fused device evaluation (``csb_eval_metrics``): what ``set_pressure_grid`` + ``reweight_target`` + ``reweight_preds`` +
        ``create_metrics_df`` compute for the metrics MAE / RMSE / R2 / bias, in one pass over CUDA tensors ``preds`` (N,128),
        ``targets`` (N,128), ``inputs`` (N,124).  Returns a DataFrame indexed by output index (like ``metrics_idx_<split>``)."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        assert preds.is_cuda and targets.is_cuda and inputs.is_cuda
        p, t, x = (a.to(torch.float32).contiguous() for a in (preds, targets, inputs))
        n = p.shape[0]
        out = torch.empty(4, 128, dtype=torch.float64, device=p.device)
        scratch = torch.empty(4 * 128 * self.num_latlon + self.num_latlon + 256, dtype=torch.float64, device=p.device)
        _, _, out_scale = self.save_norm()
        hyai, hybi = np.ascontiguousarray(self.hyai), np.ascontiguousarray(self.hybi)
        aw, osc = np.ascontiguousarray(self.area_wgt), np.ascontiguousarray(out_scale)
        _lib.check(lib.csb_eval_metrics(p.data_ptr(), t.data_ptr(), x.data_ptr(), n, self.num_latlon, hyai.ctypes.data, hybi.ctypes.data,
                                        self.p0, aw.ctypes.data, osc.ctypes.data, float(_val(self.input_mean["state_ps"])),
                                        float(_val(self.input_max["state_ps"])), float(_val(self.input_min["state_ps"])),
                                        1 if self.normalize else 0, out.data_ptr(), scratch.data_ptr(), _lib.current_stream_ptr()),
                   "csb_eval_metrics")
        df = pd.DataFrame(out.cpu().numpy().T, columns=["MAE", "RMSE", "R2", "bias"], index=range(128))
        df.index.name = "output_idx"
        return df

    # ------------------------------------------------------------------------------------------------ CNN layouts (:1693-1760)
    @staticmethod
    def _cuda_reshape(fn_name, t, out_shape):
        from . import _lib
        lib = _lib.load()
        x = t.to(torch.float32).contiguous()
        out = torch.empty(out_shape, dtype=torch.float32, device=x.device)
        _lib.check(getattr(lib, fn_name)(x.data_ptr(), out.data_ptr(), x.shape[0], _lib.current_stream_ptr()), fn_name)
        return out

    @staticmethod
    def reshape_input_for_cnn(npy_input, save_path=""):
        if torch is not None and isinstance(npy_input, torch.Tensor) and npy_input.is_cuda:
            return data_utils._cuda_reshape("csb_reshape_input_for_cnn", npy_input, (npy_input.shape[0], 60, 6))
        x = np.asarray(npy_input)
        out = np.stack([x[:, 0:60], x[:, 60:120]] + [np.repeat(x[:, c][:, np.newaxis], 60, axis=1) for c in range(120, 124)], axis=2)
        if save_path != "":
            with open(save_path + "train_input_cnn.npy", "wb") as f:
                np.save(f, np.float32(out))
        return out

    @staticmethod
    def reshape_target_for_cnn(npy_target, save_path=""):
        if torch is not None and isinstance(npy_target, torch.Tensor) and npy_target.is_cuda:
            return data_utils._cuda_reshape("csb_reshape_target_for_cnn", npy_target, (npy_target.shape[0], 60, 10))
        y = np.asarray(npy_target)
        out = np.stack([y[:, 0:60], y[:, 60:120]] + [np.repeat(y[:, c][:, np.newaxis], 60, axis=1) for c in range(120, 128)], axis=2)
        if save_path != "":
            with open(save_path + "train_target_cnn.npy", "wb") as f:
                np.save(f, np.float32(out))
        return out

    @staticmethod
    def reshape_target_from_cnn(npy_predict_cnn, save_path=""):
        if torch is not None and isinstance(npy_predict_cnn, torch.Tensor) and npy_predict_cnn.is_cuda:
            return data_utils._cuda_reshape("csb_reshape_target_from_cnn", npy_predict_cnn, (npy_predict_cnn.shape[0], 128))
        p = np.asarray(npy_predict_cnn)
        out = np.concatenate([p[:, :, 0], p[:, :, 1]] + [np.mean(p[:, :, c], axis=1)[:, np.newaxis] for c in range(2, 10)], axis=1)
        if save_path != "":
            with open(save_path + "cnn_predict_reshaped.npy", "wb") as f:
                np.save(f, np.float32(out))
        return out
