"""VariationalMergingModel: the reference's public training interface on top of the CUDA engine.

Mirror of careless/models/merging/variational.py:11-275.  ``train_model(data, steps)`` uploads the input
tuple once, runs ``steps`` fused ELBO-gradient + Adam steps on the GPU and returns the same history
dict (keys ``loss``, ``NLL``, ``F KLDiv``, ``Grad Norm``), stopping early on a non-finite gradient
norm exactly like the reference loop (:271-274).  Trained parameters are written back into the
surrogate / scaler objects so ``mean() / stddev() / save_weights`` behave as in careless.
"""
from __future__ import annotations

import numpy as np

from ..base import BaseModel
from ..priors.wilson import DoubleWilsonPrior, WilsonPrior
from ..scaling.image import HybridImageScaler, NeuralImageScaler
from ..scaling.nn import MetadataScaler
from ... import parallel
from ...engine import Engine, EngineConfig
from ...optimizers import Adam


class VariationalMergingModel(BaseModel):
    def __init__(self, surrogate_posterior, prior, likelihood, scaling_model, mc_sample_size=1, kl_weight=None,
                 scale_kl_weight=None, scale_prior=None):
        if scale_prior is not None:
            raise NotImplementedError("scale_prior is never enabled by the reference CLI (variational.py:159-163) and is not implemented")
        self.prior = prior
        self.surrogate_posterior = surrogate_posterior
        self.likelihood = likelihood
        self.scaling_model = scaling_model
        self.mc_sample_size = mc_sample_size
        self.kl_weight = kl_weight
        self.optimizer = Adam(beta_2=0.99)
        self.seed = 1234
        self.device = 0
        self._engine = None
        self._engine_key = None
        self._engine_data = None
        self._ranks = None

    def compile(self, optimizer=None, run_eagerly=None, **kwargs):
        if optimizer is not None and not isinstance(optimizer, str):
            self.optimizer = optimizer
        return self

    # ------------------------------------------------------------------ engine plumbing
    def _parts(self):
        sm = self.scaling_model
        if isinstance(sm, NeuralImageScaler):
            return sm.metadata_scaler, None
        if isinstance(sm, HybridImageScaler):
            return sm.mlp_scaler, sm.image_scaler
        if isinstance(sm, MetadataScaler):
            return sm, None
        raise TypeError(f"unsupported scaling model {type(sm).__name__}")

    def _build_engine(self, data, validation_data=None):
        """The engine holding `data` (cached per input tuple).  In a multi-process job (``parallel.context()``: torchrun
        sets WORLD_SIZE > 1) the reflections are partitioned over the ranks: this process's engine holds
        its own surrogate entries and their observations, and the library's NCCL communicator is created for it."""
        mlp, img = self._parts()
        metadata = self.get_metadata(data)
        refl_id = np.asarray(self.get_refl_id(data)).reshape(-1)
        n_meta = metadata.shape[-1]
        mlp.build(n_meta)
        laue = bool(getattr(self.likelihood, "laue", False))
        if laue != self.is_laue(data):
            raise ValueError("Laue likelihood needs the 8-entry Laue input tuple (and vice versa)")
        q, prior, opt = self.surrogate_posterior, self.prior, self.optimizer
        R = q.loc_raw.shape[0]
        dw = isinstance(prior, DoubleWilsonPrior)
        nis = self.scaling_model if isinstance(self.scaling_model, NeuralImageScaler) else None
        ctx = parallel.context()
        cfg = EngineConfig(
            n_refl=R, n_meta=n_meta, mlp_width=mlp.width, mlp_layers=mlp.n_layers,
            n_images=(img.max_images if img is not None else (nis.max_images if nis is not None else 0)), image_scales=img is not None,
            image_layers=(len(nis.image_layers) if nis is not None else 0),
            refine_uncertainties=bool(getattr(self.likelihood, "refine_uncertainties", False)),
            mc_samples=int(self.mc_sample_size), likelihood=self.likelihood.kind, dof=self.likelihood.dof, laue=laue,
            prior="double_wilson" if dw else "wilson", n_asu=(len(prior.r) if dw else 0),
            optimize_dw_r=(prior.optimize_r if dw else False), scale_bijector=mlp.scale_bijector,
            scale_shift=mlp.scale_multiplier, epsilon=q.scale_shift, kl_weight=self.kl_weight,
            learning_rate=opt.learning_rate, beta_1=opt.beta_1, beta_2=opt.beta_2, adam_epsilon=opt.epsilon,
            clipnorm=opt.clipnorm, clipvalue=opt.clipvalue, global_clipnorm=opt.global_clipnorm,
            seed=self.seed, device=(ctx.local_rank if ctx.active else self.device), rank=ctx.rank, world_size=ctx.world)
        # the cache holds a strong reference to the tuple it was built from and compares identity (`is`): an id() alone
        # could be reused by a new tuple after the old one was garbage-collected
        key = tuple(sorted(cfg.__dict__.items()))
        if self._engine is not None and self._engine_key == key and self._engine_data is data:
            return self._engine
        if self._engine is not None:
            self._engine.close()
        self._ranks = None
        if ctx.active:
            extra = ()
            if validation_data is not None:
                extra = ((self.get_refl_id(validation_data), self.get_harmonic_id(validation_data) if laue else None),)
            self._ranks = parallel.partition(R, refl_id, ctx.world, self.get_harmonic_id(data) if laue else None,
                                             prior.dw_parent if dw else None, extra=extra)
        eng = self._engine_for(data, cfg, self._ranks)
        self._engine, self._engine_key, self._engine_data = eng, key, data
        return eng

    def _engine_for(self, data, cfg, ranks):
        """An engine with the rows of `data` (all of them, or this rank's share under the partition `ranks`)."""
        import dataclasses
        ctx = parallel.context()
        laue, prior = bool(cfg.laue), self.prior
        dw = isinstance(prior, DoubleWilsonPrior)
        refl_id = np.asarray(self.get_refl_id(data)).reshape(-1)
        R = self.surrogate_posterior.loc_raw.shape[0]
        if ranks is None:
            eng = Engine(cfg)
            eng.mine = None
            eng.set_observations(refl_id, self.get_image_id(data), self.get_metadata(data), self.get_intensities(data),
                                 self.get_uncertainties(data), harmonic_id=self.get_harmonic_id(data) if laue else None)
            eng.set_prior(prior.centric, prior.epsilon, prior.sigma, dw_parent=prior.dw_parent if dw else None,
                          asu_id=prior.asu_ids if dw else None, r=prior.r if dw else None, init_scale=-1.0)
            return eng
        inputs = {"refl_id": refl_id, "image_id": self.get_image_id(data), "metadata": self.get_metadata(data),
                  "intensities": self.get_intensities(data), "uncertainties": self.get_uncertainties(data),
                  "harmonic_id": self.get_harmonic_id(data) if laue else None}
        sigma = prior.sigma
        tables = {"centric": prior.centric, "multiplicity": prior.epsilon,
                  "sigma": None if sigma is None else np.broadcast_to(np.asarray(sigma, dtype=np.float32), (R,)),
                  "dw_parent": prior.dw_parent if dw else None, "asu_id": prior.asu_ids if dw else None}
        li, lt = parallel.shard(inputs, tables, ranks, ctx.rank, laue=laue)
        mine = lt["refl_index"]
        if len(mine) == 0 or len(li["refl_id"]) == 0:
            raise ValueError(f"rank {ctx.rank} of {ctx.world} received no reflections / observations: use fewer processes for this data set")
        eng = Engine(dataclasses.replace(cfg, n_refl=len(mine), n_refl_total=R))
        eng.mine = mine
        eng.set_observations(li["refl_id"], li.get("image_id"), li["metadata"], li["intensities"], li["uncertainties"],
                             harmonic_id=li.get("harmonic_id"), obs_index=li["obs_index"], n_rows_total=len(refl_id))
        eng.set_prior(lt["centric"], lt["multiplicity"], lt.get("sigma"), dw_parent=lt.get("dw_parent"), asu_id=lt.get("asu_id"),
                      r=prior.r if dw else None, refl_index=mine, init_scale=-1.0)
        parallel.init_engine_comm(eng, ctx)
        return eng

    # per-reflection vectors travel as this rank's slice; gathered back with a host-side sum over ranks
    @staticmethod
    def _local(eng, full):
        full = np.asarray(full)
        return full if eng.mine is None else full[eng.mine]

    @staticmethod
    def _gather(eng, local, n_total):
        if eng.mine is None:
            return local
        full = np.zeros(n_total, dtype=local.dtype)
        full[eng.mine] = local
        return parallel.context().allsum(full)

    def _push(self, eng):
        mlp, img = self._parts()
        q = self.surrogate_posterior
        eng.set_params("sf_loc_raw", self._local(eng, q.loc_raw))
        eng.set_params("sf_scale_raw", self._local(eng, q.scale_raw))
        eng.set_params("mlp", mlp.flat())
        if img is not None:
            eng.set_params("image_scales", img._scales)
        if isinstance(self.scaling_model, NeuralImageScaler) and self.scaling_model.image_layers:
            eng.set_params("image_layers", self.scaling_model.flat())
            eng.set_trainable("image_layers", self.scaling_model.trainable)
        if getattr(self.likelihood, "refine_uncertainties", False):
            eng.set_params("likelihood", self.likelihood.raw)
            eng.set_trainable("likelihood", self.likelihood.trainable)
        eng.set_trainable("sf_loc_raw", q.trainable)
        eng.set_trainable("sf_scale_raw", q.trainable)
        eng.set_trainable("mlp", self.scaling_model.trainable and mlp.trainable)
        if img is not None:
            eng.set_trainable("image_scales", self.scaling_model.trainable and img.trainable)

    def _pull(self, eng):
        mlp, img = self._parts()
        q = self.surrogate_posterior
        R = q.loc_raw.shape[0]
        q.loc_raw = self._gather(eng, eng.get_params("sf_loc_raw"), R)
        q.scale_raw = self._gather(eng, eng.get_params("sf_scale_raw"), R)
        mlp.from_flat(eng.get_params("mlp"))
        if img is not None:
            img._scales = eng.get_params("image_scales")
        if isinstance(self.scaling_model, NeuralImageScaler) and self.scaling_model.image_layers:
            self.scaling_model.from_flat(eng.get_params("image_layers"))
        if getattr(self.likelihood, "refine_uncertainties", False):
            self.likelihood.raw = eng.get_params("likelihood").astype(np.float32)
        if isinstance(self.prior, DoubleWilsonPrior) and self.prior.optimize_r:
            self.prior.r = (1.0 / (1.0 + np.exp(-eng.get_params("dw_r_logit")))).astype(np.float32)

    # ------------------------------------------------------------------ the reference's training entry
    def train_model(self, data, steps, message=None, format_string="{:0.2e}", validation_data=None,
                    validation_frequency=10, progress=True, use_custom_train_step=True, jit_compile=None,
                    reduce_retracing=False, chunk=100):
        """variational.py:226-275.  Returns history: dict[str, list[float]], one entry per step taken.
        With ``validation_data`` the held-out NLL (keras test_on_batch, :257-260) is evaluated after every
        ``validation_frequency``-th step on a second engine and logged as ``NLL_val`` scaled by
        len(train)/len(validation) (:249)."""
        eng = self._build_engine(data, validation_data)
        self._push(eng)
        progress = progress and parallel.context().rank == 0
        veng, val_scale, nll_val = None, None, None
        if validation_data is not None:
            val_scale = len(data[0]) / len(validation_data[0])
            veng = self._validation_engine(validation_data)
        history = {}
        done = 0
        bar = None
        if progress:
            from tqdm import tqdm
            bar = tqdm(total=steps, desc=message)
        stopped = False
        while done < steps and not stopped:
            if veng is not None:
                to_val = (-done) % validation_frequency           # steps until the next validated step
                n = 1 if to_val == 0 else min(chunk, to_val, steps - done)
            else:
                n = min(chunk, steps - done)
            rows = eng.step(n)            # metrics stay on the device for the whole chunk (no per-step host sync)
            if veng is not None and done % validation_frequency == 0 and rows:
                for g in ("sf_loc_raw", "sf_scale_raw", "mlp", "image_scales", "dw_r_logit", "image_layers", "likelihood"):
                    if eng.group_size(g) > 0:
                        veng.set_params(g, eng.get_params(g))
                nll_val = val_scale * veng.eval()["NLL"]
            for row in rows:
                for k, v in row.items():
                    history.setdefault(k, []).append(float(v))
                if veng is not None:
                    history.setdefault("NLL_val", []).append(float(nll_val))
            done += len(rows)
            if bar is not None:
                bar.update(len(rows))
                bar.set_postfix({k: format_string.format(v[-1]) for k, v in history.items()})
            # variational.py:271-274: the loop ends AFTER the step whose gradient norm was non-finite.  clb_step returns
            # fewer rows only when the bad step is not the last of its chunk, so the metric itself is checked too.
            if len(rows) < n or (rows and not np.isfinite(rows[-1]["Grad Norm"])):
                print("Encountered numerical issues, terminating optimization early!")
                stopped = True
        if bar is not None:
            bar.close()
        if veng is not None:
            veng.close()
        self._pull(eng)
        return history

    def _validation_engine(self, validation_data):
        """A second engine holding the held-out rows (same model configuration and partition, its own RNG stream)."""
        import dataclasses
        eng = self._engine
        cfg = dataclasses.replace(eng.cfg, seed=eng.cfg.seed + 0x9E3779B9, n_refl=self.surrogate_posterior.loc_raw.shape[0])
        return self._engine_for(validation_data, cfg, self._ranks)

    # ------------------------------------------------------------------ post-hoc moments ("next" row 2)
    def scale_mean_stddev(self, inputs):
        """variational.py:47-78: moments of the posterior of the scale of every observation (Laue: convolved)."""
        eng = self._build_engine(inputs)
        self._push(eng)
        mean, stddev = self._scale_moments(eng)
        if getattr(self.likelihood, "laue", False):
            hid = self.get_harmonic_id(inputs)
            mean = self.likelihood.convolve(mean, hid)
            stddev = np.sqrt(self.likelihood.convolve(stddev * stddev, hid))
        return mean, stddev

    def prediction_mean_stddev(self, inputs):
        """variational.py:80-121: expected intensity <Sigma><F^2> and its standard deviation per observation."""
        eng = self._build_engine(inputs)
        self._push(eng)
        smean, sstd = self._scale_moments(eng)
        res = self._results(eng)
        refl_id = np.asarray(self.get_refl_id(inputs)).reshape(-1)
        f2 = np.square(res["F"]) + np.square(res["SigF"])
        iexp = smean * f2[refl_id]
        f4 = self.surrogate_posterior.moment_4(method='scipy')
        s2 = np.square(smean) + np.square(sstd)
        ivar = f4[refl_id] * s2 - iexp * iexp
        if getattr(self.likelihood, "laue", False):
            hid = self.get_harmonic_id(inputs)
            iexp = self.likelihood.convolve(iexp, hid)
            ivar = self.likelihood.convolve(ivar, hid)
        return iexp, np.sqrt(ivar)

    def get_results(self, inputs):
        """Numeric part of DataManager.get_results (io/manager.py:188-209): dict of F, SigF, I, SigI, N per reflection."""
        eng = self._build_engine(inputs)
        self._push(eng)
        return self._results(eng)

    def _scale_moments(self, eng):
        mean, std = eng.get_scale_moments()          # original row order; rows of other ranks are 0
        if eng.mine is not None:
            ctx = parallel.context()
            mean, std = ctx.allsum(mean), ctx.allsum(std)
        return mean, std

    def _results(self, eng):
        res = eng.get_results()
        R = self.surrogate_posterior.loc_raw.shape[0]
        return {k: self._gather(eng, v, R) for k, v in res.items()}

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
            self._engine_data = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
