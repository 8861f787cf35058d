"""GPU tests of the careless-shaped public API (careless_b200.models), written like the reference's own
tests/models/merging/test_variational_{mono,laue}.py: every likelihood x prior x scaler combination trains
and stays finite -- plus what the reference never checks: the history matches the oracle."""
import numpy as np
import pytest

from careless_b200 import synth
from careless_b200.models.likelihoods import laue as laue_lik
from careless_b200.models.likelihoods import mono as mono_lik
from careless_b200.models.merging.surrogate_posteriors import TruncatedNormal
from careless_b200.models.merging.variational import VariationalMergingModel
from careless_b200.models.priors.wilson import DoubleWilsonPrior, WilsonPrior
from careless_b200.models.scaling.image import HybridImageScaler, ImageScaler, NeuralImageScaler
from careless_b200.models.scaling.nn import MLPScaler
from careless_b200.optimizers import Adam

pytestmark = pytest.mark.gpu


def _inputs(p, laue):
    col = lambda a: np.asarray(a).reshape(-1, 1)
    t = (col(p["refl_id"]), col(p["image_id"]), col(p["file_id"]), p["metadata"], col(p["intensities"]), col(p["uncertainties"]))
    if laue:
        t = t + (col(p["wavelength"]), col(p["harmonic_id"]))
    return t


def _model(p, laue, likelihood, prior_kind, scaler_kind, mc_samples):
    if prior_kind == "wilson":
        prior = WilsonPrior(p["centric"], p["multiplicity"])
    else:
        prior = DoubleWilsonPrior(p["centric"], p["multiplicity"], p["asu_id"], p["reflids"], p["root"], p["r"],
                                  optimize_r=(prior_kind == "dw_opt"))
    loc, scale = prior.mean(), prior.stddev()
    low = (1e-32 * ~np.asarray(p["centric"], dtype=bool)).astype("float32")          # manager.py:434
    q = TruncatedNormal.from_loc_and_scale(loc, scale, low)
    mod = laue_lik if laue else mono_lik
    lik = mod.NormalLikelihood() if likelihood == "normal" else mod.StudentTLikelihood(4.0)
    mlp = MLPScaler(3, 6, scale_bijector="exp")
    scaler = mlp if scaler_kind == "mlp" else HybridImageScaler(mlp, ImageScaler(int(p["n_images"])))
    model = VariationalMergingModel(q, prior, lik, scaler, mc_samples)
    model.compile(Adam(1e-2, 0.9, 0.99))
    return model


@pytest.mark.parametrize("laue", [False, True])
@pytest.mark.parametrize("likelihood", ["normal", "studentt"])
@pytest.mark.parametrize("scaler_kind", ["mlp", "hybrid"])
@pytest.mark.parametrize("mc_samples", [1, 3])
def test_train_model_runs_and_improves(laue, likelihood, scaler_kind, mc_samples):
    p = synth.make_laue(3000, 300, d=3, n_images=12, seed=4) if laue else synth.make_mono(3000, 300, d=3, n_images=12, seed=4)
    model = _model(p, laue, likelihood, "wilson", scaler_kind, mc_samples)
    data = _inputs(p, laue)
    hist = model.train_model(data, 40, progress=False)
    assert set(hist) >= {"loss", "NLL", "F KLDiv", "Grad Norm"}
    assert all(len(v) == 40 for v in hist.values())
    assert np.all(np.isfinite([hist[k] for k in hist]))
    assert np.mean(hist["loss"][-5:]) < np.mean(hist["loss"][:5])
    q = model.surrogate_posterior
    assert np.all(np.isfinite(q.mean())) and np.all(q.stddev() > 0) and np.all(np.isfinite(q.moment_4()))
    model.close()


@pytest.mark.parametrize("prior_kind", ["dw", "dw_opt"])
def test_double_wilson_model(prior_kind):
    p = synth.make_double_wilson(1500, 200, n_datasets=3, d=3, n_images=5, r=0.9, seed=5)
    model = _model(p, False, "normal", prior_kind, "hybrid", 1)
    r0 = model.prior.r.copy()
    hist = model.train_model(_inputs(p, False), 30, progress=False)
    assert np.all(np.isfinite(hist["loss"]))
    if prior_kind == "dw_opt":
        assert not np.allclose(model.prior.r[1:], r0[1:])
    else:
        assert np.allclose(model.prior.r, r0)
    with pytest.raises(ValueError):
        DoubleWilsonPrior(p["centric"], p["multiplicity"], p["asu_id"], p["reflids"], p["root"], [0.0, 1.0, 0.5])
    model.close()


def test_freezing_and_weight_roundtrip(tmp_path):
    """careless.py:48-56,79-80,104: save/load weights, frozen scaler for the half-dataset re-merge."""
    p = synth.make_mono(2000, 200, d=3, n_images=8, seed=6)
    model = _model(p, False, "normal", "wilson", "hybrid", 1)
    data = _inputs(p, False)
    model.train_model(data, 10, progress=False)
    mlp = model.scaling_model.mlp_scaler
    w_before = [w.copy() for w in mlp.get_weights()]
    s_before = model.scaling_model.image_scaler._scales.copy()
    q_before = model.surrogate_posterior.loc_raw.copy()
    mlp.save_weights(tmp_path / "scale"); model.surrogate_posterior.save_weights(tmp_path / "sf")
    model.scaling_model.trainable = False                     # careless.py:104
    model.train_model(data, 5, progress=False)
    assert all(np.array_equal(a, b) for a, b in zip(w_before, mlp.get_weights()))
    assert np.array_equal(s_before, model.scaling_model.image_scaler._scales)
    assert not np.array_equal(q_before, model.surrogate_posterior.loc_raw)
    model.surrogate_posterior.load_weights(tmp_path / "sf"); mlp.load_weights(tmp_path / "scale")
    assert np.array_equal(q_before, model.surrogate_posterior.loc_raw)
    model.close()


def test_history_matches_oracle_through_public_api():
    import torch
    from oracle import model as om
    from oracle import philox
    p = synth.make_mono(2500, 250, d=3, n_images=9, seed=7)
    model = _model(p, False, "studentt", "wilson", "hybrid", 2)
    model.seed = 4242
    hist = model.train_model(_inputs(p, False), 3, progress=False)
    ocfg = om.ModelConfig(n_refl=250, n_meta=3, mlp_width=6, mlp_layers=3, likelihood="studentt", dof=4.0,
                          mc_samples=2, image_scales=True, n_images=12)
    ocfg.n_images = int(p["n_images"])
    oprior = om.PriorData(p["centric"], p["multiplicity"])
    params = om.init_params(ocfg, oprior)
    draws = [(philox.refl_uniforms(4242, s, 2, np.arange(250)), philox.obs_normals(4242, s, 2, np.arange(2500))) for s in range(3)]
    _, ohist, _ = om.train(params, p, oprior, ocfg, om.AdamConfig(lr=1e-2), draws)
    for i in range(3):
        for k in ("loss", "NLL", "F KLDiv", "Grad Norm"):
            assert abs(hist[k][i] - ohist[i][k]) <= 2e-4 * abs(ohist[i][k]) + 1e-6, (i, k, hist[k][i], ohist[i][k])
    model.close()


def test_validation_nll_follows_the_reference_loop():
    """variational.py:249,257-260: NLL_val = len(train)/len(val) * test_on_batch NLL, refreshed every
    validation_frequency steps; checked against the oracle's forward pass with the validation engine's own draws."""
    from oracle import model as om
    from oracle import philox
    p = synth.make_mono(3000, 250, d=3, n_images=9, seed=8)
    rng = np.random.default_rng(1)
    test = rng.random(3000) < 0.2
    split = lambda m: {k: (v[m] if isinstance(v, np.ndarray) and v.shape[:1] == (3000,) else v) for k, v in p.items()}
    ptr, pte = split(~test), split(test)
    model = _model(p, False, "normal", "wilson", "mlp", 1)
    model.seed = 77
    hist = model.train_model(_inputs(ptr, False), 25, validation_data=_inputs(pte, False), validation_frequency=10, progress=False)
    assert len(hist["NLL_val"]) == 25 and np.all(np.isfinite(hist["NLL_val"]))
    v = np.array(hist["NLL_val"])
    assert np.all(v[0:10] == v[0]) and np.all(v[10:20] == v[10]) and np.all(v[20:25] == v[20]) and v[0] != v[10]
    # oracle: forward on the held-out rows with the trained-at-step-21 parameters is not reproducible here, so check step 0:
    model2 = _model(p, False, "normal", "wilson", "mlp", 1)
    model2.seed = 77
    h2 = model2.train_model(_inputs(ptr, False), 1, validation_data=_inputs(pte, False), validation_frequency=10, progress=False)
    ocfg = om.ModelConfig(n_refl=250, n_meta=3, mlp_width=6, mlp_layers=3)
    oprior = om.PriorData(p["centric"], p["multiplicity"])
    import torch
    q = model2.surrogate_posterior
    params = om.init_params(ocfg, oprior)
    params["sf_loc_raw"] = torch.as_tensor(q.loc_raw.astype(np.float64)); params["sf_scale_raw"] = torch.as_tensor(q.scale_raw.astype(np.float64))
    ws = model2.scaling_model.get_weights()
    names = [f"mlp.{k}.{n}" for k in range(3) for n in ("kernel", "bias")] + ["mlp.out.kernel", "mlp.out.bias"]
    for n, w in zip(names, ws):
        params[n] = torch.as_tensor(w.astype(np.float64))
    vseed = 77 + 0x9E3779B9
    nte = int(test.sum())
    out = om.forward(params, pte, oprior, ocfg, philox.refl_uniforms(vseed, 0, 1, np.arange(250)), philox.obs_normals(vseed, 0, 1, np.arange(nte)))
    expect = (3000 - nte) / nte * float(out["nll"])
    assert abs(h2["NLL_val"][0] - expect) <= 2e-4 * abs(expect), (h2["NLL_val"][0], expect)
    model.close(); model2.close()


@pytest.mark.parametrize("laue", [False, True])
def test_neural_image_scaler_trains(laue):
    """careless --image-layers=2 (tests/test_cli.py:211-228 exercises it through the CLI)."""
    p = synth.make_laue(3000, 300, d=3, n_images=6, seed=9) if laue else synth.make_mono(3000, 300, d=3, n_images=6, seed=9)
    prior = WilsonPrior(p["centric"], p["multiplicity"])
    low = (1e-32 * ~np.asarray(p["centric"], dtype=bool)).astype("float32")
    q = TruncatedNormal.from_loc_and_scale(prior.mean(), prior.stddev(), low)
    lik = (laue_lik if laue else mono_lik).NormalLikelihood()
    scaler = NeuralImageScaler(2, int(p["n_images"]), 3, 6, scale_bijector="exp")
    model = VariationalMergingModel(q, prior, lik, scaler, 1)
    model.compile(Adam(1e-2, 0.9, 0.99))
    w0 = scaler.image_layers[0].w.copy()
    hist = model.train_model(_inputs(p, laue), 30, progress=False)
    assert np.all(np.isfinite(hist["loss"])) and np.mean(hist["loss"][-5:]) < np.mean(hist["loss"][:5])
    assert scaler.image_layers[0].w.shape == (6, 6, 6) and not np.allclose(scaler.image_layers[0].w, w0)
    model.close()


@pytest.mark.parametrize("laue", [False, True])
def test_results_and_predictions(laue):
    """DataManager.get_results / get_predictions numerics (io/manager.py:188-209, variational.py:47-121):
    merged F/SigF/I/SigI vs scipy.stats.truncnorm, scale moments vs the oracle's scale network."""
    import torch
    from scipy.stats import truncnorm
    from oracle import model as om
    p = synth.make_laue(3000, 300, d=3, n_images=8, seed=10) if laue else synth.make_mono(3000, 300, d=3, n_images=8, seed=10)
    model = _model(p, laue, "normal", "wilson", "hybrid", 1)
    data = _inputs(p, laue)
    model.train_model(data, 20, progress=False)
    res = model.get_results(data)
    q = model.surrogate_posterior
    loc, scale = q.loc.astype(np.float64), q.scale.astype(np.float64)
    low = q.low.astype(np.float64)
    a, b = (low - loc) / scale, (1e10 - loc) / scale
    F, SigF = truncnorm.mean(a, b, loc, scale), truncnorm.std(a, b, loc, scale)
    assert np.allclose(res["F"], F, rtol=1e-5) and np.allclose(res["SigF"], SigF, rtol=1e-5)
    I = SigF ** 2 + F ** 2
    f4 = truncnorm.moment(4, a, np.inf, loc, scale)
    SigI = np.sqrt(np.maximum((I * 1e-5) ** 2, f4 - I * I))
    assert np.allclose(res["I"], I, rtol=1e-5) and np.allclose(res["SigI"], SigI, rtol=2e-4)
    assert np.array_equal(res["N"], np.bincount(p["refl_id"], minlength=300).astype(np.float32))
    # scale moments against the oracle's network with the trained weights
    ocfg = om.ModelConfig(n_refl=300, n_meta=3, mlp_width=6, mlp_layers=3, image_scales=True, n_images=int(p["n_images"]), laue=laue)
    mlp = model.scaling_model.mlp_scaler
    names = [f"mlp.{k}.{n}" for k in range(3) for n in ("kernel", "bias")] + ["mlp.out.kernel", "mlp.out.bias"]
    params = {n: torch.as_tensor(w.astype(np.float64)) for n, w in zip(names, mlp.get_weights())}
    params["image_scales"] = torch.as_tensor(model.scaling_model.image_scaler._scales.astype(np.float64))
    mu_s, sig_s, shift = om.scale_network(params, p, ocfg, torch.float64)
    aimg = om.image_scale_vector(params, p, ocfg, torch.float64)
    smean, sstd = (aimg * mu_s).numpy(), (aimg.abs() * sig_s).numpy()
    iexp = smean * (F ** 2 + SigF ** 2)[p["refl_id"]]
    ivar = f4[p["refl_id"]] * (smean ** 2 + sstd ** 2) - iexp ** 2
    if laue:
        conv = lambda v: np.bincount(p["harmonic_id"], weights=v, minlength=len(v))
        smean, sstd, iexp, ivar = conv(smean), np.sqrt(conv(sstd ** 2)), conv(iexp), conv(ivar)
    m, sd = model.scale_mean_stddev(data)
    assert np.allclose(m, smean, rtol=2e-4, atol=1e-5) and np.allclose(sd, sstd, rtol=2e-4, atol=1e-6)
    ip, sip = model.prediction_mean_stddev(data)
    assert np.allclose(ip, iexp, rtol=5e-4, atol=1e-4) and np.allclose(sip, np.sqrt(ivar), rtol=2e-3, atol=1e-4)
    model.close()


@pytest.mark.parametrize("with_validation", [False, True])
@pytest.mark.parametrize("bad_step", [3, 4, 7])
def test_nonfinite_gradient_norm_ends_training_at_any_chunk_position(monkeypatch, with_validation, bad_step):
    """variational.py:271-274: the loop ends AFTER the first step whose gradient norm is not finite -- also when that step is the last
    one of a chunk of steps handed to the library (clb_step then returns every row of the chunk) or a one-step chunk in front of a
    validation step.  The non-finite norm is injected into the metrics the engine returns."""
    from careless_b200.engine import Engine
    p = synth.make_mono(1500, 120, d=3, n_images=5, seed=3)
    model = _model(p, False, "normal", "wilson", "mlp", 1)
    real_step = Engine.step
    seen = {"n": 0}

    def step(self, n, *a, **kw):
        rows = real_step(self, n, *a, **kw)
        for r in rows:
            if seen["n"] == bad_step:
                r["Grad Norm"] = float("nan")
                del rows[rows.index(r) + 1:]           # the library stops after the bad step: later rows of the chunk do not exist
                seen["n"] += 1
                break
            seen["n"] += 1
        return rows

    monkeypatch.setattr(Engine, "step", step)
    kw = {}
    if with_validation:
        test = np.random.default_rng(0).random(1500) < 0.25
        split = lambda m: {k: (v[m] if isinstance(v, np.ndarray) and v.shape[:1] == (1500,) else v) for k, v in p.items()}
        kw = dict(validation_data=_inputs(split(test), False), validation_frequency=4)
        data = _inputs(split(~test), False)
    else:
        data = _inputs(p, False)
    hist = model.train_model(data, 20, progress=False, chunk=4, **kw)
    assert len(hist["loss"]) == bad_step + 1, (len(hist["loss"]), bad_step)
    assert not np.isfinite(hist["Grad Norm"][-1]) and np.all(np.isfinite(hist["Grad Norm"][:-1]))
    model.close()
