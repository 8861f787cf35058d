"""B200TrainMixin -- the binding a careless maintainer would add, as importable code.

    class VariationalMergingModel(B200TrainMixin, careless.models.merging.variational.VariationalMergingModel): ...

It replaces ONLY `train_model` (careless/models/merging/variational.py:226-275) and talks to the model through the reference's
own object protocol, never through this repository's mirror classes:
  * `self.surrogate_posterior.trainable_variables`  -> [loc_raw, scale_raw]  (the pretransformed inputs of the two
    `tfp.util.TransformedVariable`s of `TruncatedNormal.from_loc_and_scale`, surrogate_posteriors.py:104-131)
  * `self.scaling_model.trainable_variables`        -> keras order [kernel_0, bias_0, ..., kernel_out, bias_out] (nn.py:55-79)
  * `self.prior.centric / .epsilon / .sigma`        (priors/wilson.py:43-47)
  * `self.likelihood` (StudentT has `.dof`, likelihoods/mono.py:25-37), `self.mc_sample_size`, `self.kl_weight`,
    `self.optimizer.{learning_rate, beta_1, beta_2, epsilon, clipnorm, clipvalue, global_clipnorm}` (io/manager.py:494-501)
  * the input accessors of `BaseModel` (models/base.py:39-121).
A "variable" only needs `.numpy()` and `.assign(value)`; that is all keras / TFP variables are asked for, which is why the
test-suite can drive this file with small stand-in objects on a machine without TensorFlow
(tests/test_gpu_integration.py).  Scope: WilsonPrior, mono or Laue Normal / StudentT likelihood, MLPScaler with the exp or
softplus bijector; anything else raises NotImplementedError instead of silently training a different model.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine, EngineConfig


def _np(x):
    return np.asarray(x.numpy() if hasattr(x, "numpy") else x)


class B200TrainMixin:
    #: the reference applies tfb.Exp / tfb.Chain([Shift(eps), Exp]) (io/manager.py:457-463); override for softplus models
    b200_scale_bijector = "exp"
    b200_epsilon = 1e-7
    b200_seed = 1234
    b200_device = 0

    def _b200_engine(self, data):
        q, prior, scaler, opt = self.surrogate_posterior, self.prior, self.scaling_model, self.optimizer
        if type(prior).__name__ != "WilsonPrior":
            raise NotImplementedError(f"B200TrainMixin binds WilsonPrior models; got {type(prior).__name__} (use careless_b200.models for the rest)")
        mlp_vars = [_np(v) for v in scaler.trainable_variables]
        if len(mlp_vars) < 2 or len(mlp_vars) % 2 or mlp_vars[-2].ndim != 2 or mlp_vars[-2].shape[1] != 2:
            raise NotImplementedError("B200TrainMixin expects an MLPScaler: [kernel, bias] * L + [kernel_out (W, 2), bias_out (2)]")
        metadata = np.asarray(_np(self.get_metadata(data)), dtype=np.float32)
        n_layers = len(mlp_vars) // 2 - 1
        width = int(mlp_vars[0].shape[1]) if n_layers > 0 else int(metadata.shape[1])
        laue = bool(self.is_laue(data))
        lik = self.likelihood
        dof = getattr(lik, "dof", None)

        def hyper(name, default=None):
            v = getattr(opt, name, default)
            return None if v is None else float(_np(v))
        cfg = EngineConfig(
            n_refl=len(np.asarray(prior.centric)), n_meta=int(metadata.shape[1]), mlp_width=width, mlp_layers=n_layers,
            mc_samples=int(self.mc_sample_size), likelihood="normal" if dof is None else "studentt",
            dof=None if dof is None else float(dof), laue=laue, scale_bijector=self.b200_scale_bijector, epsilon=self.b200_epsilon,
            kl_weight=getattr(self, "kl_weight", None), learning_rate=hyper("learning_rate", 1e-3), beta_1=hyper("beta_1", 0.9),
            beta_2=hyper("beta_2", 0.99), adam_epsilon=hyper("epsilon", 1e-7), clipnorm=hyper("clipnorm"), clipvalue=hyper("clipvalue"),
            global_clipnorm=hyper("global_clipnorm"), seed=self.b200_seed, device=self.b200_device)
        eng = Engine(cfg)                                                         # clb_create
        eng.set_observations(_np(self.get_refl_id(data)), _np(self.get_image_id(data)), metadata, _np(self.get_intensities(data)),
                             _np(self.get_uncertainties(data)), harmonic_id=_np(self.get_harmonic_id(data)) if laue else None)
        sigma = np.asarray(prior.sigma, dtype=np.float32)
        eng.set_prior(np.asarray(prior.centric), np.asarray(prior.epsilon),
                      None if sigma.ndim == 0 and float(sigma) == 1.0 else np.broadcast_to(sigma, np.asarray(prior.epsilon).shape),
                      init_scale=-1.0)
        return eng, mlp_vars

    def train_model(self, data, steps, message=None, format_string="{:0.2e}", validation_data=None, validation_frequency=10,
                    progress=True, **kwargs):
        """variational.py:226-275 on the B200: same arguments, same history dict, early stop on a non-finite gradient norm."""
        if validation_data is not None:
            raise NotImplementedError("validation_data: use careless_b200.models.merging.variational.VariationalMergingModel (clb_eval)")
        q, scaler = self.surrogate_posterior, self.scaling_model
        eng, mlp_vars = self._b200_engine(data)
        try:
            qv = list(q.trainable_variables)
            eng.set_params("sf_loc_raw", _np(qv[0]).reshape(-1))
            eng.set_params("sf_scale_raw", _np(qv[1]).reshape(-1))
            eng.set_params("mlp", np.concatenate([w.reshape(-1) for w in mlp_vars]))
            eng.set_trainable("mlp", bool(getattr(scaler, "trainable", True)))
            for g in ("sf_loc_raw", "sf_scale_raw"):
                eng.set_trainable(g, bool(getattr(q, "trainable", True)))
            rows = eng.step(int(steps))                                           # clb_step: the hot loop, no host code between steps
            qv[0].assign(eng.get_params("sf_loc_raw").reshape(_np(qv[0]).shape))   # write back: get_results / save_weights keep working
            qv[1].assign(eng.get_params("sf_scale_raw").reshape(_np(qv[1]).shape))
            flat, off = eng.get_params("mlp"), 0
            for var, w in zip(scaler.trainable_variables, mlp_vars):
                var.assign(flat[off:off + w.size].reshape(w.shape)); off += w.size
        finally:
            eng.close()
        if len(rows) < steps or (rows and not np.isfinite(rows[-1]["Grad Norm"])):
            print("Encountered numerical issues, terminating optimization early!")
        return {k: [float(r[k]) for r in rows] for k in (rows[0] if rows else {})}
