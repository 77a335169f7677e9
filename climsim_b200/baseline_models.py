"""``torch.nn.Module`` faces of the reference's baseline column emulators, backed by the CUDA engine.

The reference builds MLP_v1 / ED as inline Keras graphs (baseline_models/MLP/training/HPO/baseline_v1/
hpo_baseline_v1.py:75-103, baseline_models/ED/training/ClimSIM_ED_1_3_train.py:56-92) and has no importable model
class; these modules span the same hyper-parameter space with ``forward(x: (B,124)) -> (B,128)``, keep
their parameters in ONE flat fp32 ``nn.Parameter`` (Keras ``get_weights()`` order, kernels (in,out)), and run forward and
backward through ``csb_mlp_forward`` / ``csb_mlp_backward``.  Any torch optimizer and any loss written in torch works on
top.  The fused path (loss + backward + optimizer inside the engine) is ``climsim_b200.trainer.Trainer``.

There is no CPU path: constructing a module without a B200 raises ``CsbError`` (CSB_ENODEV).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .engine import CNNEngine, MLPEngine
from .trainer import glorot_uniform_flat


class _EngineFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flat, module):
        eng = module.engine
        module._sync_params()
        ctx.module, ctx.need_dx = module, x.requires_grad
        return eng.forward(x, keep_activations=True, training=module.training)      # module.train(): the Dropout layers are active

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.module.engine
        dx = eng.backward(dy.contiguous(), need_dx=ctx.need_dx)
        return dx, eng.get_grads_device(), None


class _EngineModule(torch.nn.Module):
    def __init__(self, engine: MLPEngine, seed: int = 0):
        super().__init__()
        self.engine = engine
        self.flat = torch.nn.Parameter(torch.from_numpy(glorot_uniform_flat(engine.layer_dims, seed, engine.layernorm)).cuda())
        self._uploaded_version = -1

    def _sync_params(self) -> None:
        if self.flat._version != self._uploaded_version:       # the optimizer updated the parameter in place
            self.engine.set_params_device(self.flat.detach())
            self._uploaded_version = self.flat._version

    def _apply_own_config(self) -> None:
        """Engine state this module relies on (output mask, no input transform); re-applied after a wrapper borrowed the engine."""
        self.engine.set_input_transform()
        self.engine.set_output_mask(None)

    def _claim_engine(self) -> None:
        if getattr(self.engine, "config_owner", None) is not None:
            self._apply_own_config()
            self.engine.config_owner = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._claim_engine()
        if torch.is_grad_enabled() and (self.flat.requires_grad or x.requires_grad):
            return _EngineFunction.apply(x, self.flat, self)
        self._sync_params()
        return self.engine.forward(x)

    # -- Keras-style weight access -------------------------------------------------------------------------------
    def layer_views(self) -> List[torch.Tensor]:
        """[W0 (in,out), b0, W1, b1, ...] as views of the flat parameter."""
        out, off = [], 0
        for (k, n), ln in zip(self.engine.layer_dims, self.engine.layernorm):
            out.append(self.flat.detach()[off:off + k * n].view(k, n)); off += k * n
            for _ in range(3 if ln else 1):                   # bias (, gamma, beta)
                out.append(self.flat.detach()[off:off + n]); off += n
        return out

    def load_flat(self, flat: np.ndarray) -> None:
        with torch.no_grad():
            self.flat.copy_(torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float32)))

    def load_keras_h5(self, path: str) -> dict:
        """Load a Keras ``.h5`` checkpoint of the reference (``ModelCheckpoint`` / ``model.save``: step2_retrain.py:252-262; the shipped
        ``baseline_models/MLP/model/*.best.h5`` and ``baseline_models/ED/model/ED_ClimSIM_1_3_model.h5``) -- no h5py / TensorFlow
        needed (``climsim_b200.keras_h5``).  Returns what the file holds (weights, optimizer slots, configs)."""
        from .keras_h5 import read_keras_h5
        ck = read_keras_h5(path)
        fused = getattr(self, "out_lin", None) is not None                  # MLP_v1: the two output Dense layers are one fused layer here
        flat = MLPEngine.keras_to_flat(ck["weights"], fused_head=fused)
        assert flat.size == self.flat.numel(), f"{path}: {flat.size} parameters, this module has {self.flat.numel()}"
        self.load_flat(flat)
        return ck


class MLP(_EngineModule):
    """MLP_v1: ``x -> [Dense(u) -> act]* -> Dense(128) -> act -> concat(Dense(120), relu(Dense(8)))``.
    Defaults = the shipped best trial (step1_results.csv lot-147 / trial_0027)."""

    def __init__(self, units: Sequence[int] = (768, 640, 512, 640, 640), activation: str = "leakyrelu", alpha: float = 0.15,
                 in_dim: int = 124, out_lin: int = 120, out_relu: int = 8, dtype: str = "bf16", max_batch: int = 65536,
                 seed: int = 0):
        self.out_lin = out_lin
        super().__init__(MLPEngine.mlp_v1(units=units, act=activation, alpha=alpha, in_dim=in_dim, out_lin=out_lin,
                                          out_relu=out_relu, dtype=dtype, max_batch=max_batch), seed)

    def load_keras_weights(self, weights: Sequence[np.ndarray]) -> None:
        """``keras_model.get_weights()`` (the two output Dense layers separate, as Keras stores them)."""
        self.load_flat(MLPEngine.keras_to_flat(weights, fused_head=True))

    def keras_weights(self) -> List[np.ndarray]:
        return self.engine.flat_to_keras(self.flat.detach().cpu().numpy(), out_lin=self.out_lin)


class ED(_EngineModule):
    """Encoder-decoder MLP 124->463->463->231->115->57->28->5->28->...->463->128, ReLU, ELU output
    (baseline_models/ED/training/ClimSIM_ED_1_3_train.py:56-92; Keras truncates the float widths 463/2**k)."""

    def __init__(self, intermediate_dim: int = 463, latent_dim: int = 5, in_dim: int = 124, out_dim: int = 128,
                 dtype: str = "bf16", max_batch: int = 65536, seed: int = 0):
        d = intermediate_dim
        widths = [d, d, int(d / 2), int(d / 4), int(d / 8), int(d / 16), latent_dim,
                  int(d / 16), int(d / 8), int(d / 4), int(d / 2), d, d, out_dim]
        layers = [(w, "relu", 0.0) for w in widths[:-1]] + [(out_dim, "elu", 0.0)]
        super().__init__(MLPEngine(in_dim, layers, head_relu_from=-1, dtype=dtype, max_batch=max_batch), seed)


class HSRMLP(_EngineModule):
    """``MLP`` of baseline_models/HSR/training/hsr.py:14-35: layers x [Linear -> LayerNorm -> Dropout(p) -> ReLU] -> Linear.
    ``dropout`` > 0 is active in training (``module.train()`` forwards that record a graph, and ``HSR.trainer``'s steps) with the
    engine's own counter-based masks (torch's random stream is not reproduced); the shipped configuration uses p = 0
    (hpo.py:225-238)."""

    def __init__(self, in_dims: int = 124, out_dims: int = 128, hidden_dims: int = 512, layers: int = 1, dropout: float = 0.0,
                 dtype: str = "bf16", max_batch: int = 16384, seed: int = 0):
        spec = [(hidden_dims, "relu", 0.0)] * layers + [(out_dims, "none", 0.0)]
        self.n_hidden = layers
        super().__init__(MLPEngine(in_dims, spec, dtype=dtype, max_batch=max_batch, layernorm=[True] * layers + [False]), seed)
        if dropout > 0:
            self.engine.set_dropout(dropout, seed)

    def load_reference_state_dict(self, sd, prefix: str = "") -> None:
        """Keys of the reference module: ``linear{i}.0.weight`` (out,in), ``linear{i}.0.bias``, ``linear{i}.1.weight`` (gamma),
        ``linear{i}.1.bias`` (beta), ``final_linear.weight``, ``final_linear.bias``."""
        parts = []
        for i in range(self.n_hidden):
            parts += [sd[f"{prefix}linear{i}.0.weight"].t().reshape(-1), sd[f"{prefix}linear{i}.0.bias"],
                      sd[f"{prefix}linear{i}.1.weight"], sd[f"{prefix}linear{i}.1.bias"]]
        parts += [sd[f"{prefix}final_linear.weight"].t().reshape(-1), sd[f"{prefix}final_linear.bias"]]
        self.load_flat(torch.cat([p.detach().float().cpu().reshape(-1) for p in parts]).numpy())

    def reference_state_dict(self, prefix: str = "") -> dict:
        """The parameters under the reference module's keys and layouts (inverse of ``load_reference_state_dict``)."""
        views, sd, i = self.layer_views(), {}, 0
        for l in range(self.n_hidden):
            w, b, g, be = views[i:i + 4]
            i += 4
            sd[f"{prefix}linear{l}.0.weight"], sd[f"{prefix}linear{l}.0.bias"] = w.t().contiguous().clone(), b.clone()
            sd[f"{prefix}linear{l}.1.weight"], sd[f"{prefix}linear{l}.1.bias"] = g.clone(), be.clone()
        sd[f"{prefix}final_linear.weight"], sd[f"{prefix}final_linear.bias"] = views[i].t().contiguous().clone(), views[i + 1].clone()
        return sd


class HSR(torch.nn.Module):
    """``HeteroskedasticRegression`` (hsr.py:38-81): two LayerNorm MLPs estimating mean and log-precision; ``forward`` returns
    ``(mean, logprec)``; the losses of hsr.py:126-138 are ordinary torch expressions on top."""

    def __init__(self, in_dims: int = 124, out_dims: int = 128, hidden_dims: int = 512, layers: int = 1, dropout: float = 0.0,
                 dtype: str = "bf16", max_batch: int = 16384):
        super().__init__()
        self.mean = HSRMLP(in_dims, out_dims, hidden_dims, layers, dropout, dtype, max_batch, seed=0)
        self.logprec = HSRMLP(in_dims, out_dims, hidden_dims, layers, dropout, dtype, max_batch, seed=1)

    def forward(self, x):
        return self.mean(x), self.logprec(x)

    def sample(self, x, random: bool = True):
        mu, logprec = self.forward(x)
        if random:
            return mu + torch.randn_like(mu) * torch.exp(logprec) ** (-0.5)
        return mu, torch.exp(logprec) ** (-0.5)

    def load_reference_state_dict(self, sd) -> None:
        self.mean.load_reference_state_dict(sd, "mean.")
        self.logprec.load_reference_state_dict(sd, "logprec.")

    def reference_state_dict(self) -> dict:
        """``state_dict`` with the reference's keys (``mean.linear0.0.weight`` ... ``logprec.final_linear.bias``): what its
        ``torch.save(self.state_dict(), save)`` checkpoints hold, loadable by ``hsr.HeteroskedasticRegression.load_state_dict``."""
        return {**self.mean.reference_state_dict("mean."), **self.logprec.reference_state_dict("logprec.")}

    def trainer(self, data, epochs: int = 20, save: str = "models/vae.cp", plot: bool = True, loss_type: str = "mle",
                optimizer: str = "adam", lr: float = 0.0001, gamma: float = 0.01, rho: Optional[float] = None,
                checkpoint_every_s: float = 1200.0):
        """``HeteroskedasticRegression.trainer`` (hsr.py:83-142): same arguments, same returned per-step losses, same end state.

        Every batch is ONE engine call (``csb_hsr_train_step``): the forward passes of both networks, the loss -- MSE on the mean
        for the first third of the epochs, the Gaussian negative log-likelihood afterwards, clipped to +-1e5 -- its gradients, both
        backward passes and the optimizer (Adam or SGD with the per-group L2 decay alpha = (1-rho)/rho*gamma on the mean network
        and beta = (1-rho)/rho*(1-gamma) on the log-precision network) run as CUDA kernels of this library.  No torch kernel and
        no host synchronisation per step: the losses accumulate in a device array that is read once per epoch.  A checkpoint with
        the reference's ``state_dict`` keys is written every ``checkpoint_every_s`` seconds (20 minutes in the reference)."""
        import time
        from .engine import hsr_train_step
        from . import _lib
        if loss_type != "mle":
            raise ValueError("Unknown loss")
        if optimizer not in ("adam", "sgd"):
            raise ValueError("Unknown optimizer")
        rho = rho if rho is not None else 1 - gamma
        alpha, beta = (1 - rho) / rho * gamma, (1 - rho) / rho * (1 - gamma)
        print("alpha: %.3f, beta: %.3f" % (alpha, beta))
        rule = "adam_torch" if optimizer == "adam" else "sgd"
        device = self.mean.flat.device
        for m in (self.mean, self.logprec):
            m._claim_engine()
            m._sync_params()
        scratch = torch.zeros(_lib.BATCH_METRICS_SCRATCH, dtype=torch.float64, device=device)
        losses: List[float] = []
        t_ckpt = time.time()
        steps_per_epoch = len(data) if hasattr(data, "__len__") else 0
        buf = torch.zeros(max(steps_per_epoch, 64), dtype=torch.float32, device=device)
        for epoch in range(epochs):
            k = 0
            for batch in data:
                x, y = batch["x"].to(device, non_blocking=True), batch["y"].to(device, non_blocking=True)
                if k == buf.numel():                                    # an iterable without __len__: grow the per-epoch loss array
                    buf = torch.cat([buf, torch.zeros_like(buf)])
                hsr_train_step(self.mean.engine, self.logprec.engine, x, y, mle=not (epoch < epochs / 3), loss_out=buf[k:k + 1],
                               scratch=scratch, rule=rule, lr=lr, wd_mean=alpha, wd_logprec=beta)
                k += 1
                if time.time() - t_ckpt > checkpoint_every_s:
                    self._pull_params()
                    torch.save(self.reference_state_dict(), save)
                    t_ckpt = time.time()
            losses += buf[:k].tolist()                                  # the one D2H read of the epoch
            steps_per_epoch = k
        self._pull_params()
        print("Last-epoch loss: %.2f" % sum(losses[len(losses) - steps_per_epoch:-1]))
        print("Finished Training")
        if plot:
            try:
                import matplotlib.pyplot as plt
                plt.plot(np.array(losses)[:-1])
            except ImportError:
                pass
        return losses

    def _pull_params(self) -> None:
        """Engine-side training updates the parameters inside the engine: mirror them into the modules' ``flat`` parameters."""
        for m in (self.mean, self.logprec):
            with torch.no_grad():
                m.engine.get_params_device(out=m.flat.data)
            m._uploaded_version = m.flat._version


class OnlineMLP(_EngineModule):
    """The online-testing MLP (online_testing/baseline_models/MLP_v2rh/training/mlp.py:24-68): ``layers`` x [Linear -> ReLU] ->
    Linear, optional ``output_prune`` (zero the top ``strato_lev_out`` levels of the q1, q2, q3 and u tendencies), ReLU on the last
    eight outputs.  Same constructor arguments as the reference class (``dropout`` > 0: active in training forwards, with the
    engine's own counter-based masks)."""

    def __init__(self, in_dims: int = 557, out_dims: int = 368, hidden_dims=(384, 1024, 640), layers: int = 3, dropout: float = 0.0,
                 output_prune: bool = False, strato_lev_out: int = 15, dtype: str = "bf16", max_batch: int = 16384, seed: int = 0):
        if isinstance(hidden_dims, (list, tuple)):
            assert len(hidden_dims) == layers, "Length of hidden_dims should be equal to layers"
            hidden = list(hidden_dims)
        else:
            hidden = [hidden_dims] * layers
        spec = [(h, "relu", 0.0) for h in hidden] + [(out_dims, "none", 0.0)]
        super().__init__(MLPEngine(in_dims, spec, head_relu_from=out_dims - 8, dtype=dtype, max_batch=max_batch), seed)
        self.output_prune, self.strato_lev_out = output_prune, strato_lev_out
        if dropout > 0:
            self.engine.set_dropout(dropout, seed)
        self._apply_own_config()

    def _own_mask(self) -> Optional[np.ndarray]:
        if not self.output_prune:
            return None
        mask = np.ones(self.engine.out_dim, np.float32)
        for start in (60, 120, 180, 240):
            mask[start:start + self.strato_lev_out] = 0
        return mask

    def _apply_own_config(self) -> None:
        self.engine.set_input_transform()
        self.engine.set_output_mask(self._own_mask())

    def load_reference_state_dict(self, sd) -> None:
        """Keys ``linears.{i}.0.weight`` (out,in) / ``.bias`` and ``final_linear.*`` of the reference module."""
        parts = []
        for i in range(len(self.engine.layer_dims) - 1):
            parts += [sd[f"linears.{i}.0.weight"].t().reshape(-1), sd[f"linears.{i}.0.bias"]]
        parts += [sd["final_linear.weight"].t().reshape(-1), sd["final_linear.bias"]]
        self.load_flat(torch.cat([p.detach().float().cpu().reshape(-1) for p in parts]).numpy())


class OnlineInferenceWrapper(torch.nn.Module):
    """``NewModel`` of online_testing/model_postprocessing/v2_nn_wrapper.ipynb (cell 5), the module exported for the E3SM coupling:
    raw, un-normalised ``(B, 557)`` in, physical-unit ``(B, 368)`` tendencies out.  ``preprocessing`` (``1 - exp(-lambda q)`` for
    cloud liquid / ice, normalisation, nan / inf -> 0, pruning of the top ``prune_qn_levels`` cloud levels, clipping of relative
    humidity to [0, 1.2]), the network and ``postprocessing`` (zeroing of the stratospheric tendencies, division by ``out_scale``)
    run as ONE engine call: the prologue kernel, the GEMM chain, and the output layer's epilogue carrying mask and scale.

    Same constructor arguments as the reference class; the pruning ranges it hard-codes are keyword arguments with its values."""

    def __init__(self, original_model: "OnlineMLP", input_sub, input_div, out_scale, lbd_qc, lbd_qi, prune_qn_levels: int = 15,
                 rh_clip=(0.0, 1.2), out_prune=((60, 75), (120, 148), (180, 195), (240, 255), (300, 315))):
        super().__init__()
        self.original_model = original_model
        eng = original_model.engine
        n_in, n_out = eng.in_dim, eng.out_dim
        lam = np.zeros(n_in, np.float32)
        lam[120:180] = np.asarray(lbd_qc, np.float32).reshape(-1)
        lam[180:240] = np.asarray(lbd_qi, np.float32).reshape(-1)
        keep = np.ones(n_in, np.float32)
        keep[120:120 + prune_qn_levels] = 0
        keep[180:180 + prune_qn_levels] = 0
        lo, hi = np.full(n_in, -np.inf, np.float32), np.full(n_in, np.inf, np.float32)
        lo[60:120], hi[60:120] = rh_clip
        own = original_model._own_mask()                      # the network's own output_prune (mlp.py:56-61)
        mask = own.copy() if own is not None else np.ones(n_out, np.float32)
        for a, b in out_prune:
            mask[a:b] = 0
        self._cfg = dict(sub=np.asarray(input_sub, np.float32), div=np.asarray(input_div, np.float32),
                         scale=np.asarray(out_scale, np.float32), lam=lam, keep=keep, lo=lo, hi=hi, mask=mask)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        m, c = self.original_model, self._cfg
        eng = m.engine
        m._sync_params()
        if getattr(eng, "config_owner", None) is not self:    # configure the engine once; the wrapped module re-claims it on its next call
            eng.set_norm(inp_sub=c["sub"], inp_div=c["div"], out_scale=c["scale"])
            eng.set_input_transform(c["lam"], c["keep"], c["lo"], c["hi"])
            eng.set_output_mask(c["mask"])
            eng.config_owner = self
        return eng.forward(x, normalize_in=True, denorm_out=True)


class _CNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flat, module):
        module._sync_params()
        y = module.engine.forward(x)
        ctx.module = module
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        eng = ctx.module.engine
        eng.backward(y, dy.contiguous())
        return None, eng.get_grads_device(), None


class CNN(torch.nn.Module):
    """ResNet-1D of baseline_models/CNN/training/hpo_train.py:131-200 (a Keras model in the reference): ``forward(x: (B,60,6))
    -> (B,60,10)``.  Two ways to train it: the fused ``train_step`` (loss = ``mae_adjusted`` or ``mse_adjusted``, Dropout 0.175 behind
    the ReLUs, the optimizer inside the engine -- the way ``model.fit`` drives the reference, hpo_train.py:355-368), or ordinary
    torch autograd: the parameters are ONE flat ``nn.Parameter`` (Keras ``get_weights()`` order, the two Dense heads concatenated),
    ``forward`` records an autograd node whose backward is ``csb_cnn_backward``, so any torch loss and optimizer work on top
    (dropout-free, like ``model(x, training=False)``; no gradient w.r.t. ``x``)."""

    def __init__(self, depth: int = 12, width: int = 406, kernel: int = 3, loss: str = "mae", dtype: str = "bf16", max_batch: int = 4096,
                 dropout: float = 0.175, seed: int = 0):
        super().__init__()
        self.engine = CNNEngine(depth=depth, width=width, kernel=kernel, loss=loss, dtype=dtype, max_batch=max_batch)
        if dropout > 0 and dtype == "bf16":                   # hp_dropout = 0.175 (hpo_train.py:143); active in train_step only
            self.engine.set_dropout(dropout, seed)
        rng = np.random.default_rng(seed)
        parts = []
        for shp in self.engine.shapes():                      # Keras defaults: glorot_uniform kernels, zero biases
            if len(shp) == 1:
                parts.append(np.zeros(shp, np.float32))
            else:
                fan_in = int(np.prod(shp[:-1]))
                fan_out = int(shp[0] * shp[-1]) if len(shp) == 3 else int(shp[-1])
                lim = float(np.sqrt(6.0 / (fan_in + fan_out)))
                parts.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
        self.flat = torch.nn.Parameter(torch.from_numpy(np.concatenate([p.reshape(-1) for p in parts])).cuda())
        self._uploaded_version = -1

    def _sync_params(self) -> None:
        if self.flat._version != self._uploaded_version:
            self.engine.set_params_device(self.flat.detach())
            self._uploaded_version = self.flat._version

    def load_keras_weights(self, weights: Sequence[np.ndarray]) -> None:
        with torch.no_grad():
            self.flat.copy_(torch.from_numpy(CNNEngine.keras_to_flat(weights)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and self.flat.requires_grad:
            return _CNNFunction.apply(x, self.flat, self)
        self._sync_params()
        return self.engine.forward(x)

    def train_step(self, x: torch.Tensor, y: torch.Tensor, lr: float = 1e-4, rule: str = "adam_keras") -> torch.Tensor:
        """Fused step inside the engine (the parameters it updates live in the engine; ``pull_params`` mirrors them into ``flat``)."""
        self._sync_params()
        loss = self.engine.train_step(x, y)
        self.engine.apply_opt(rule, lr=lr)
        return loss

    def pull_params(self) -> None:
        with torch.no_grad():
            self.flat.copy_(torch.from_numpy(self.engine.get_params_flat()).to(self.flat.device))
        self._uploaded_version = self.flat._version
