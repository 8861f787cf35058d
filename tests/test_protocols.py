"""The reference's object protocols on the mirror classes: surrogate .sample/.log_prob/.parameter_properties
(surrogate_posteriors.py:11-37), likelihood(inputs).log_prob / .convolve (likelihoods/mono.py:16-37, laue.py:9-34) and the
callable scaler -> dist with mean/stddev/sample (variational.py:67-69, 156-157).  Host-side helpers are checked against
closed forms here; on the GPU they are tied to what the kernels compute."""
import numpy as np
import pytest
from scipy import stats

from careless_b200 import synth
from careless_b200.models.likelihoods import laue as ll, mono as lm
from careless_b200.models.merging.surrogate_posteriors import TruncatedNormal


def _tuple(p, laue=False):
    col = lambda a, t: np.asarray(a).reshape(-1, 1).astype(t)
    base = (col(p["refl_id"], np.int64), col(p["image_id"], np.int64), col(np.zeros(len(p["refl_id"])), np.int64),
            p["metadata"].astype(np.float32), col(p["intensities"], np.float32), col(p["uncertainties"], np.float32))
    if laue:
        base += (col(p["wavelength"], np.float32), col(p["harmonic_id"], np.int64))
    return base


def test_surrogate_protocol():
    rng = np.random.default_rng(0)
    loc, scale = rng.uniform(0.5, 3.0, 50), rng.uniform(0.1, 1.0, 50)
    low = np.where(rng.random(50) < 0.3, 0.0, 1e-32)
    q = TruncatedNormal.from_loc_and_scale(loc, scale, low)
    z = q.sample(4000, seed=3)
    assert z.shape == (4000, 50) and np.all(z >= low)
    assert np.allclose(z.mean(0), q.mean(), rtol=0.05, atol=0.02)            # draws follow the distribution whose moments we report
    lp = q.log_prob(z[:5])
    a, b = (low - q.loc) / q.scale, (1e10 - q.loc) / q.scale
    assert np.allclose(lp, stats.truncnorm.logpdf(z[:5].astype(np.float64), a, b, q.loc, q.scale))
    props = q.parameter_properties()
    assert set(props) == {"loc", "scale", "low", "high"} and props["loc"]["bijector"] == "Exp"
    assert set(q.parameters) == {"loc", "scale", "low", "high"}


def test_mono_likelihood_protocol():
    p = synth.make_mono(300, 40, d=2, n_images=3, seed=5)
    inputs = _tuple(p)
    x = np.random.default_rng(1).normal(p["intensities"], 1.0, size=(2, 300))
    i, s = p["intensities"].astype(np.float64), p["uncertainties"].astype(np.float64)
    assert np.allclose(lm.NormalLikelihood()(inputs).log_prob(x), stats.norm.logpdf(x, i, s))          # tests/models/likelihoods/test_mono.py:12-28
    assert np.allclose(lm.StudentTLikelihood(7.0)(inputs).log_prob(x), stats.t.logpdf(x, 7.0, i, s))   # :38-51
    ev = lm.NormalEv11Likelihood()
    sp = np.logaddexp(0, x)
    sig = ev.Sdfac * np.sqrt(s * s + ev.SdB * sp + ev.Sdadd * sp * sp)                                   # mono.py:46-59
    assert np.allclose(ev(inputs).log_prob(x), stats.norm.logpdf(x, i, sig))


def test_laue_likelihood_protocol():
    """tests/models/likelihoods/test_laue.py:11-36: convolving I / multiplicity gives I back on the first n_spots entries."""
    p = synth.make_laue(400, 60, d=2, n_images=4, seed=6)
    inputs = _tuple(p, laue=True)
    hid, n_spots = p["harmonic_id"], p["n_spots"]
    counts = np.bincount(hid, minlength=len(hid))
    lik = ll.NormalLikelihood()(inputs)
    ipred = (p["intensities"][hid] / counts[hid])[None, :]
    conv = lik.convolve(ipred)
    assert np.allclose(conv[0, :n_spots], p["intensities"][:n_spots], rtol=1e-6) and np.all(conv[0, n_spots:] == 0)
    ref = stats.norm.logpdf(conv, p["intensities"].astype(np.float64), p["uncertainties"].astype(np.float64))
    assert np.allclose(lik.log_prob(ipred), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["mlp", "hybrid", "image_layers"])
def test_callable_scaler_returns_gpu_moments(kind):
    """scaling_model(inputs) -> dist with mean / stddev / sample: the moments come from the CUDA forward pass and equal the oracle's."""
    import torch
    from careless_b200.models.scaling.image import HybridImageScaler, ImageScaler, NeuralImageScaler
    from careless_b200.models.scaling.nn import MLPScaler
    from oracle import model as om
    p = synth.make_mono(3000, 300, d=3, n_images=7, seed=8)
    p["image_id"] = np.sort(p["image_id"])
    inputs = _tuple(p)
    rng = np.random.default_rng(2)
    if kind == "image_layers":
        scaler = NeuralImageScaler(2, 7, 3, 10, scale_bijector="exp")
        mlp = scaler.metadata_scaler
    else:
        mlp = MLPScaler(4, 10, scale_bijector="exp")
        scaler = mlp if kind == "mlp" else HybridImageScaler(mlp, ImageScaler(7))
    mlp.build(3)
    mlp.set_weights([w + 0.05 * rng.standard_normal(w.shape).astype(np.float32) for w in mlp.get_weights()])
    if kind == "hybrid":
        scaler.image_scaler._scales = (1.0 + 0.1 * rng.standard_normal(6)).astype(np.float32)
    if kind == "image_layers":
        for l in scaler.image_layers:
            l.w = l.w + 0.05 * rng.standard_normal(l.w.shape).astype(np.float32)
            l.b = l.b + 0.05 * rng.standard_normal(l.b.shape).astype(np.float32)
    dist = scaler(inputs)
    # the oracle's scale network on the same weights
    cfg = om.ModelConfig(n_refl=300, n_meta=3, mlp_width=10, mlp_layers=mlp.n_layers, image_scales=(kind == "hybrid"), n_images=7,
                         image_layers=2 if kind == "image_layers" else 0)
    prior = om.PriorData(p["centric"], p["multiplicity"])
    params = om.init_params(cfg, prior)
    ws = mlp.get_weights()
    for k in range(mlp.n_layers):
        params[f"mlp.{k}.kernel"], params[f"mlp.{k}.bias"] = torch.as_tensor(ws[2 * k].astype(np.float64)), torch.as_tensor(ws[2 * k + 1].astype(np.float64))
    params["mlp.out.kernel"], params["mlp.out.bias"] = torch.as_tensor(ws[-2].astype(np.float64)), torch.as_tensor(ws[-1].astype(np.float64))
    if kind == "hybrid":
        params["image_scales"] = torch.as_tensor(scaler.image_scaler._scales.astype(np.float64))
    if kind == "image_layers":
        for k, l in enumerate(scaler.image_layers):
            params[f"image_layer.{k}.kernel"], params[f"image_layer.{k}.bias"] = torch.as_tensor(l.w.astype(np.float64)), torch.as_tensor(l.b.astype(np.float64))
    mu_s, sig_s, shift = om.scale_network(params, p, cfg, torch.float64)
    aimg = om.image_scale_vector(params, p, cfg, torch.float64)
    aimg = 1.0 if aimg is None else aimg
    mean, std = (aimg * (mu_s + shift)).numpy(), (abs(aimg) * sig_s).numpy()
    assert np.allclose(dist.mean(), mean, rtol=1e-4, atol=1e-6) and np.allclose(dist.stddev(), std, rtol=1e-4, atol=1e-6)
    assert dist.sample(3, seed=0).shape == (3, 3000)


@pytest.mark.gpu
def test_likelihood_protocol_agrees_with_the_kernel():
    """-sum likelihood(inputs).log_prob(ipred) / S (variational.py:169-178) == the NLL the CUDA step reports for the same ipred."""
    import _util as U
    p = synth.make_mono(4000, 500, d=3, n_images=9, seed=9)
    ocfg, oprior, eng = U.build(p, mlp_width=10, mlp_layers=5, likelihood="studentt", dof=9.0, mc_samples=2)
    try:
        eng.enable_ipred(True)
        hist = eng.step(1)
        ip = eng.get_ipred()
        nll = -lm.StudentTLikelihood(9.0)(_tuple(p)).log_prob(ip.astype(np.float64)).sum() / 2
        assert abs(nll - hist[0]["NLL"]) <= 1e-5 * abs(nll)
    finally:
        eng.close()
