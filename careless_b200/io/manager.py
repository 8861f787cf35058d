"""DataManager: model factory, train/test and half-dataset splits, results and predictions tables.

Host-side mirror of careless/io/manager.py:10-507 (same method names and argument meaning).  The tables are
careless_b200.io.mtz.DataSet column stores instead of rs.DataSet; the numbers in them come from the GPU
(`clb_get_results`, `clb_get_scale_moments`) when a model is passed, else from the host scipy formulas.
"""
from __future__ import annotations

import warnings

import numpy as np

from ..models.base import BaseModel
from ..models.priors.wilson import DoubleWilsonPrior, WilsonPrior
from .mtz import DataSet, read_mtz
from .symmetry import parse_triplet


class DataManager:
    parser = None

    def __init__(self, inputs, asu_collection, parser=None):
        self.inputs = inputs
        self.asu_collection = asu_collection
        self.parser = parser

    @classmethod
    def from_datasets(cls, datasets, formatter):
        inputs, rac = formatter(datasets)
        return cls(inputs, rac)

    @classmethod
    def from_mtz_files(cls, filenames, formatter):
        return cls.from_datasets((read_mtz(i) for i in filenames), formatter)

    @classmethod
    def from_stream_files(cls, filenames, formatter):
        from .crystfel import read_crystfel
        return cls.from_datasets((read_crystfel(i) for i in filenames), formatter)

    # ---- priors -------------------------------------------------------------------
    @staticmethod
    def wilson_sigma(b, dHKL):
        return np.exp(-0.25 * b * np.reciprocal(dHKL * dHKL))

    def get_wilson_sigma(self, b=None):
        if b is None:
            return 1.
        return self.wilson_sigma(b, self.asu_collection.dHKL)

    def get_wilson_prior(self, b=None, k=1.):
        if b is None:
            sigma = 1.
        elif isinstance(b, float):
            sigma = self.get_wilson_sigma(b)
        else:
            raise ValueError(f"parameter b has type{type(b)} but float was expected")
        return WilsonPrior(self.asu_collection.centric, self.asu_collection.multiplicity, sigma * k)

    # ---- outputs ------------------------------------------------------------------
    def get_predictions(self, model, inputs=None, test_value=0):
        """Per-observation (per-spot for Laue) Iobs, SigIobs, Ipred, SigIpred, Scale, SigScale; one table per ASU
        (manager.py:89-161)."""
        if inputs is None:
            inputs = self.inputs
        laue = BaseModel.is_laue(inputs)
        refl_id = BaseModel.get_refl_id(inputs).reshape(-1)
        asu_id, H = self.asu_collection.to_asu_id_and_miller_index(refl_id)
        asu_id = asu_id.reshape(-1)
        file_id = BaseModel.get_file_id(inputs).reshape(-1)
        image_id = BaseModel.get_image_id(inputs).reshape(-1)
        harmonic_id = BaseModel.get_harmonic_id(inputs).reshape(-1) if laue else np.arange(len(refl_id))
        _, idx = np.unique(harmonic_id, return_index=True)
        n = len(idx)
        ipred, sigipred = model.prediction_mean_stddev(inputs)
        scale, sigscale = model.scale_mean_stddev(inputs)
        iobs = BaseModel.get_intensities(inputs).reshape(-1)
        sigiobs = BaseModel.get_uncertainties(inputs).reshape(-1)
        cols = {"H": H[idx, 0].astype(np.int32), "K": H[idx, 1].astype(np.int32), "L": H[idx, 2].astype(np.int32),
                "asu_id": asu_id[idx].astype(np.int32), "image_id": image_id[idx].astype(np.int32),
                "file_id": file_id[idx].astype(np.int32), "test": np.full(n, test_value, dtype=np.int32),
                "Iobs": iobs[:n], "SigIobs": sigiobs[:n], "Ipred": np.asarray(ipred).reshape(-1)[:n],
                "SigIpred": np.asarray(sigipred).reshape(-1)[:n], "Scale": np.asarray(scale).reshape(-1)[:n],
                "SigScale": np.asarray(sigscale).reshape(-1)[:n]}
        types = {"H": "H", "K": "H", "L": "H", "asu_id": "I", "image_id": "I", "file_id": "I", "test": "I", "Iobs": "J",
                 "SigIobs": "Q", "Ipred": "J", "SigIpred": "Q", "Scale": "J", "SigScale": "Q"}
        for i, rasu in enumerate(self.asu_collection):
            sel = cols["asu_id"] == i
            yield DataSet({k: v[sel] for k, v in cols.items()}, types, rasu.cell, rasu.spacegroup, merged=False)

    def get_results(self, surrogate_posterior, inputs=None, output_parameters=True, max_intensity_snr=1e-5, model=None):
        """F, SigF, I, SigI, N (+ loc, scale ...) of the observed reflections, one merged table per ASU
        (manager.py:164-250).  With `model=` the moments are computed on the GPU from the model's engine."""
        if inputs is None:
            inputs = self.inputs
        q = surrogate_posterior
        if model is not None and max_intensity_snr == 1e-5:
            res = model.get_results(inputs)
            F, SigF, I, SigI, N = (res[k] for k in ("F", "SigF", "I", "SigI", "N"))
        else:
            F, SigF = q.mean(), q.stddev()
            I = SigF * SigF + F * F
            f4 = q.moment_4(method="scipy")
            SigI = np.sqrt(np.maximum(np.square(I * max_intensity_snr), f4 - I * I))
            N = np.bincount(BaseModel.get_refl_id(inputs).reshape(-1), minlength=len(F)).astype(np.float32)
        params = None
        if output_parameters:
            params = {k: np.asarray(v, dtype=np.float32).reshape(-1) * np.ones(len(F), dtype=np.float32)
                      for k, v in sorted(q.parameters.items())}
        asu_id, H = self.asu_collection.to_asu_id_and_miller_index(np.arange(len(F)))
        asu_id = asu_id.reshape(-1)
        results = ()
        for i, asu in enumerate(self.asu_collection):
            sel = (asu_id == i) & (N > 0)
            cols = {"H": H[sel, 0].astype(np.int32), "K": H[sel, 1].astype(np.int32), "L": H[sel, 2].astype(np.int32),
                    "F": F[sel].astype(np.float32), "SigF": SigF[sel].astype(np.float32), "I": I[sel].astype(np.float32),
                    "SigI": SigI[sel].astype(np.float32), "N": N[sel].astype(np.float32)}
            types = {"H": "H", "K": "H", "L": "H", "F": "F", "SigF": "Q", "I": "J", "SigI": "Q", "N": "I"}
            if params is not None:
                for k in sorted(params):
                    cols[k], types[k] = params[k][sel], "R"
            out = DataSet(cols, types, asu.cell, asu.spacegroup, merged=True)
            if asu.anomalous:
                out = unstack_anomalous(out)
            results += (out,)
        return results

    # ---- cross-validation splits ----------------------------------------------------
    def split_mono_data_by_mask(self, test_idx):
        test_idx = np.asarray(test_idx).reshape(-1)
        train = tuple(v[~test_idx, ...] for v in self.inputs)
        test = tuple(v[test_idx, ...] for v in self.inputs)
        return train, test

    def split_laue_data_by_mask(self, test_idx):
        """Whole spots go to one side; harmonic_id is re-numbered densely and the padded intensity columns are
        rebuilt for each side (manager.py:299-343)."""
        test_idx = np.asarray(test_idx).reshape(-1)
        harmonic_id = BaseModel.get_harmonic_id(self.inputs).reshape(-1)
        isect = np.intersect1d(harmonic_id[test_idx], harmonic_id[~test_idx])
        if len(isect) > 0:
            raise ValueError(f"test_idx splits harmonic observations with harmonic_id : {isect}")

        def split(idx):
            uni, inv = np.unique(harmonic_id[idx], return_inverse=True)
            out = ()
            for i, v in enumerate(self.inputs):
                name = BaseModel.get_name_by_index(i)
                if name in ("intensities", "uncertainties"):
                    v = np.pad(v[uni], [[0, len(inv) - len(uni)], [0, 0]], constant_values=1.)
                elif name == "harmonic_id":
                    v = inv.reshape(-1, 1).astype(np.int64)
                else:
                    v = v[idx, ...]
                out += (v,)
            return out

        return split(~test_idx), split(test_idx)

    def split_data_by_refl(self, test_fraction=0.5):
        if BaseModel.is_laue(self.inputs):
            harmonic_id = BaseModel.get_harmonic_id(self.inputs).reshape(-1)
            test_idx = (np.random.random(harmonic_id.max() + 1) <= test_fraction)[harmonic_id]
            return self.split_laue_data_by_mask(test_idx)
        test_idx = np.random.random(len(self.inputs[0])) <= test_fraction
        return self.split_mono_data_by_mask(test_idx)

    def split_data_by_image(self, test_fraction=0.5):
        image_id = BaseModel.get_image_id(self.inputs).reshape(-1)
        test_idx = np.random.random(image_id.max() + 1) <= test_fraction
        if True not in test_idx:
            test_idx[0] = True
        elif False not in test_idx:
            test_idx[0] = False
        test_idx = test_idx[image_id]
        if BaseModel.is_laue(self.inputs):
            return self.split_laue_data_by_mask(test_idx)
        return self.split_mono_data_by_mask(test_idx)

    # ---- model factory ------------------------------------------------------------------
    def _prior_from(self, opt):
        """Wilson prior, or the DoubleWilson prior when --double-wilson-parents is given (manager.py:405-429)."""
        if opt.parents is None:
            return self.get_wilson_prior(opt.wilson_prior_b)
        parents = [None if tok == "None" else int(tok) for tok in opt.parents.split(",")]
        r_values = [float(tok) for tok in opt.dwr.split(",")]
        bad = [r for r in r_values if not (-1. < r < 1.)]
        if bad:
            raise ValueError(f"Supplied --double-wilson-r value {bad[0]} outside of allowed range (-1, 1)")
        for r in r_values:
            if r < 0:
                warnings.warn(f"Supplied --double-wilson-r value {r} is negative")
        ops = None
        if opt.reindexing_ops is not None:
            ops = [parse_triplet(_hkl_to_xyz(tok)) for tok in opt.reindexing_ops.split(";")]
        return DoubleWilsonPrior.from_asu_collection(self.asu_collection, parents, r_values, ops,
                                                     sigma=self.get_wilson_sigma(opt.wilson_prior_b),
                                                     optimize_r=opt.optimize_double_wilson_r)

    def _likelihood_from(self, opt):
        """Normal / Student-t, mono / Laue, plain / Ev11 (manager.py:392-403, 438-444)."""
        if opt.type not in ("mono", "poly"):
            raise ValueError(f"unknown mode {opt.type!r}")
        if opt.type == "poly":
            from ..models.likelihoods import laue as family
        else:
            from ..models.likelihoods import mono as family
        suffix = "Ev11Likelihood" if opt.refine_uncertainties else "Likelihood"
        dof = opt.studentt_likelihood_dof
        return getattr(family, "Normal" + suffix)() if dof is None else getattr(family, "StudentT" + suffix)(dof)

    def _scaler_from(self, opt):
        """MLP (+ image scales | image layers) scale model (manager.py:446-490)."""
        from ..models.scaling.image import HybridImageScaler, ImageScaler, NeuralImageScaler
        from ..models.scaling.nn import MLPScaler
        bijector = opt.scale_bijector.lower()
        if bijector not in ("exp", "softplus"):
            raise ValueError(f"Unsupported scale bijector type, {opt.scale_bijector}")
        # softplus models work in units of the intensity spread (additive shift of the scale distribution)
        shift = float(BaseModel.get_intensities(self.inputs).std()) if bijector == "softplus" else None
        width = opt.mlp_width if opt.mlp_width is not None else BaseModel.get_metadata(self.inputs).shape[-1]
        n_images = int(np.max(BaseModel.get_image_id(self.inputs))) + 1
        common = dict(epsilon=opt.epsilon, scale_bijector=bijector, scale_multiplier=shift)
        if opt.image_layers > 0:
            return NeuralImageScaler(opt.image_layers, n_images, opt.mlp_layers, width, **common)
        mlp = MLPScaler(opt.mlp_layers, width, **common)
        return HybridImageScaler(mlp, ImageScaler(n_images)) if opt.use_image_scales else mlp

    def build_model(self, parser=None, surrogate_posterior=None, prior=None, likelihood=None, scaling_model=None, mc_sample_size=None):
        """The model `parser` describes (manager.py:380-507); any of the four parts can be overridden."""
        from ..models.merging.surrogate_posteriors import TruncatedNormal
        from ..models.merging.variational import VariationalMergingModel
        from ..optimizers import Adam
        opt = self.parser if parser is None else parser
        if opt is None:
            raise ValueError("No parser supplied, but self.parser is unset")
        likelihood = likelihood if likelihood is not None else self._likelihood_from(opt)
        prior = prior if prior is not None else self._prior_from(opt)
        if surrogate_posterior is None:
            # start at the prior's moments; acentric reflections are kept off zero (manager.py:431-436)
            lower = np.where(self.asu_collection.centric, 0., 1e-32).astype("float32")
            surrogate_posterior = TruncatedNormal.from_loc_and_scale(prior.mean(), prior.stddev() * opt.structure_factor_init_scale,
                                                                     lower, scale_shift=opt.epsilon)
        scaling_model = scaling_model if scaling_model is not None else self._scaler_from(opt)
        model = VariationalMergingModel(surrogate_posterior, prior, likelihood, scaling_model,
                                        opt.mc_samples if mc_sample_size is None else mc_sample_size, kl_weight=opt.kl_weight)
        model.seed, model.device = getattr(opt, "seed", 1234), getattr(opt, "gpu_id", 0)
        model.compile(Adam(opt.learning_rate, opt.beta_1, opt.beta_2, clipnorm=opt.clipnorm, clipvalue=opt.clipvalue,
                           global_clipnorm=opt.global_clipnorm))
        return model


def _hkl_to_xyz(op):
    """Reindexing operators are written in h,k,l ('k,h,-l'); the triplet parser speaks x,y,z."""
    return op.lower().replace("h", "x").replace("k", "y").replace("l", "z")


def unstack_anomalous(ds, columns=("F", "SigF", "I", "SigI", "N")):
    """Anomalous ASU rows -> one row per Friedel pair with X(+) / X(-) columns (rs.DataSet.unstack_anomalous as used at
    manager.py:236-246): centric reflections fill both, a missing mate is NaN, the PHENIX column order first."""
    sg = ds.spacegroup
    hkl = ds.get_hkls()
    asu, isym = sg.hkl_to_asu(hkl)
    centric = sg.is_centric(asu)
    plus = (isym % 2 == 1) | centric
    minus = (isym % 2 == 0) | centric
    uni, inv = np.unique(asu, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cols = {"H": uni[:, 0].astype(np.int32), "K": uni[:, 1].astype(np.int32), "L": uni[:, 2].astype(np.int32)}
    types = {"H": "H", "K": "H", "L": "H"}
    other = [k for k in ds.keys() if k not in ("H", "K", "L") + tuple(columns)]
    order = ["F", "SigF"], ["I", "SigI"], ["N"]
    for group in order:
        for sign, mask in (("(+)", plus), ("(-)", minus)):
            for k in group:
                if k not in ds:
                    continue
                v = np.full(len(uni), np.nan, dtype=np.float32)
                v[inv[mask]] = ds[k][mask]
                cols[k + sign] = v
                types[k + sign] = {"F": "G", "SigF": "L", "I": "K", "SigI": "M", "N": "I"}[k]
    # PHENIX order: F(+) SigF(+) F(-) SigF(-) I(+) SigI(+) I(-) SigI(-) N(+) N(-)
    anom = ["F(+)", "SigF(+)", "F(-)", "SigF(-)", "I(+)", "SigI(+)", "I(-)", "SigI(-)", "N(+)", "N(-)"]
    out = {k: cols[k] for k in ("H", "K", "L")}
    out.update({k: cols[k] for k in anom if k in cols})
    for k in other:
        for sign, mask in (("(+)", plus), ("(-)", minus)):
            v = np.full(len(uni), np.nan, dtype=np.float32)
            v[inv[mask]] = ds[k][mask]
            out[k + sign] = v
            types[k + sign] = "R"
    return DataSet(out, types, ds.cell, ds.spacegroup, merged=True)
