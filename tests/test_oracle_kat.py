"""Pin the CPU oracle to the reference's own closed-form known-answer tests.

Each test restates a test of /root/reference/tests with scipy/numpy closed forms in place of
the TensorFlow-Probability objects (which are not installable here).
"""
import math

import numpy as np
import pytest
import scipy.special
import scipy.stats as st
import torch

from oracle import model as om
from oracle import philox

T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64))


def test_centric_pdf():
    """reference tests/models/priors/test_wilson.py:13-20."""
    E = np.linspace(0.1, 3.0, 100)
    p = (2.0 / np.pi) ** 0.5 * np.exp(-0.5 * E ** 2)
    lp = om.wilson_log_prob(T(E), torch.ones(100, dtype=torch.bool), T(np.ones(100)), T(np.ones(100))).numpy()
    assert np.allclose(np.log(p), lp)
    assert np.allclose(p, np.exp(lp))


def test_acentric_pdf():
    """reference tests/models/priors/test_wilson.py:22-29."""
    E = np.linspace(0.1, 3.0, 100)
    p = 2.0 * E * np.exp(-E ** 2)
    lp = om.wilson_log_prob(T(E), torch.zeros(100, dtype=torch.bool), T(np.ones(100)), T(np.ones(100))).numpy()
    assert np.allclose(np.log(p), lp)


def test_wilson_vs_scipy_with_epsilon_sigma():
    rng = np.random.default_rng(0)
    eps = rng.choice([1.0, 2.0, 3.0, 4.0, 6.0], 200)
    sig = 0.2 + rng.random(200)
    z = 0.05 + 3 * rng.random(200)
    c = rng.random(200) < 0.5
    lp = om.wilson_log_prob(T(z), torch.as_tensor(c), T(eps), T(sig)).numpy()
    s = np.sqrt(eps * sig)
    ref = np.where(c, st.halfnorm.logpdf(z, scale=s), st.weibull_min.logpdf(z, 2.0, scale=s))
    assert np.allclose(lp, ref, rtol=1e-12, atol=1e-12)
    mean, std = om.wilson_mean_stddev(c, eps, sig)
    assert np.allclose(mean, np.where(c, st.halfnorm.mean(scale=s), st.weibull_min.mean(2.0, scale=s)))
    assert np.allclose(std, np.where(c, st.halfnorm.std(scale=s), st.weibull_min.std(2.0, scale=s)))


def test_truncated_normal_moment4_and_logprob():
    """reference tests/models/merging/test_truncated_normal.py:29-42 (scipy is the reference there too)."""
    rng = np.random.default_rng(1)
    loc, scale = rng.random((2, 100))
    scale = scale + 1e-3
    mean, std, m4 = om.tn_moments(loc, scale, 0.0, np.inf)
    a, b = (0.0 - loc) / scale, np.full(100, np.inf)
    assert np.allclose(m4, st.truncnorm.moment(4, a, b, loc, scale), rtol=1e-5)
    z = loc + scale * np.abs(rng.standard_normal(100))
    lp = om.tn_log_prob(T(z)[None], T(loc), T(scale), T(np.zeros(100)), T(np.full(100, 1e10))).numpy()[0]
    assert np.allclose(lp, st.truncnorm.logpdf(z, a, (1e10 - loc) / scale, loc, scale), rtol=1e-10, atol=1e-10)
    assert np.all(np.isneginf(om.tn_log_prob(T(-np.ones(100))[None], T(loc), T(scale), T(np.zeros(100)), T(np.full(100, 1e10))).numpy()))


def test_truncated_normal_sampler_distribution_and_gradient():
    """Inverse-CDF draws follow scipy's truncnorm; the TFP custom gradient equals d/dparam of the inverse CDF."""
    rng = np.random.default_rng(2)
    loc, scale = np.array([0.7]), np.array([0.9])
    u = (np.arange(2000) + 0.5) / 2000
    z = om.tn_sample(T(loc), T(scale), T([0.0]), T([1e10]), T(u)[:, None]).numpy()[:, 0]
    a = (0.0 - loc) / scale
    assert np.allclose(z, st.truncnorm.ppf(u, a, np.inf, loc, scale), rtol=1e-9, atol=1e-12)
    # gradient vs finite differences at fixed u
    for uu in (0.03, 0.37, 0.9, 0.999):
        l = T(loc).requires_grad_(True); s = T(scale).requires_grad_(True)
        zz = om.tn_sample(l, s, T([0.0]), T([1e10]), T([[uu]]))
        gl, gs = torch.autograd.grad(zz.sum(), [l, s])
        h = 1e-6
        f = lambda L, S: st.truncnorm.ppf(uu, (0 - L) / S, np.inf, L, S)
        assert abs(float(gl) - (f(loc + h, scale) - f(loc - h, scale))[0] / (2 * h)) < 1e-6
        assert abs(float(gs) - (f(loc, scale + h) - f(loc, scale - h))[0] / (2 * h)) < 1e-6


@pytest.mark.parametrize("dof", [1.0, 2.0, 4.0, 12.0])
def test_likelihood_definitions(dof):
    """reference tests/models/likelihoods/test_mono.py:12-51: Normal(I, sigma), StudentT(dof, I, sigma)."""
    rng = np.random.default_rng(3)
    loc = 10 * rng.standard_normal(50); scale = 0.5 + rng.random(50); x = loc + scale * rng.standard_normal(50) * 3
    assert np.allclose(om.normal_log_prob(T(x), T(loc), T(scale)).numpy(), st.norm.logpdf(x, loc, scale))
    assert np.allclose(om.studentt_log_prob(T(x), dof, T(loc), T(scale)).numpy(), st.t.logpdf(x, dof, loc, scale))


def test_laue_convolution_identity():
    """reference tests/models/likelihoods/test_laue.py:11-36: convolving I/count per harmonic reproduces I
    on the first n_spots entries; the padded tail is zero."""
    from careless_b200 import synth
    p = synth.make_laue(500, 60, d=3, n_images=5, seed=4)
    hid = p["harmonic_id"]; n_spots = p["n_spots"]
    iobs = p["intensities"].astype(np.float64)
    fake = iobs[hid] / np.bincount(hid)[hid]
    conv = om.laue_convolve(T(fake)[None], hid).numpy()[0]
    assert np.allclose(conv[:n_spots], iobs[:n_spots])
    assert np.all(conv[n_spots:] == 0)
    cfg = om.ModelConfig(n_refl=60, n_meta=3, mlp_width=3, mlp_layers=1, laue=True)
    ll = om.likelihood_log_prob(T(fake)[None].repeat(3, 1), p, cfg).numpy()
    expect = st.norm.logpdf(iobs, iobs, p["uncertainties"].astype(np.float64))
    assert np.allclose(ll[:, :n_spots], expect[None, :n_spots])


def test_rice_and_folded_normal_vs_scipy():
    """careless/utils/distributions.py:278-283 and :300-335."""
    rng = np.random.default_rng(5)
    nu = 3 * rng.random(100); sig = 0.2 + rng.random(100); x = 0.05 + 4 * rng.random(100)
    assert np.allclose(om.rice_log_prob(T(x), T(nu), T(sig)).numpy(), st.rice.logpdf(x, nu / sig, scale=sig), rtol=1e-10)
    assert np.allclose(om.folded_normal_log_prob(T(x), T(nu), T(sig)).numpy(), st.foldnorm.logpdf(x, nu / sig, scale=sig), rtol=1e-10)
    # large arguments stay finite (log I0 via the exponentially scaled Bessel function)
    assert np.isfinite(om.rice_log_prob(T([1.0]), T([1.0]), T([0.01])).numpy()).all()


def test_double_wilson_prior_reduces_to_conditionals():
    """doc/double_wilson.md:31-59: root entries follow Wilson, children Rice/Woolfson around r*z_parent."""
    rng = np.random.default_rng(6)
    R0 = 20
    centric = np.tile(rng.random(R0) < 0.4, 2); mult = np.tile(rng.choice([1.0, 2.0], R0), 2)
    prior = om.PriorData(centric, mult, 1.0, reflids=np.concatenate([np.arange(R0), np.arange(R0)]),
                         root=np.arange(2 * R0) < R0, asu_ids=np.repeat([0, 1], R0), r=np.array([0.0, 0.9]))
    cfg = om.ModelConfig(n_refl=2 * R0, n_meta=1, mlp_width=1, mlp_layers=0, prior="double_wilson")
    z = 0.1 + rng.random((1, 2 * R0))
    lp = om.prior_log_prob(T(z), {}, prior, cfg).numpy()[0]
    s = np.sqrt(mult)
    root_ref = np.where(centric, st.halfnorm.logpdf(z[0], scale=s), st.weibull_min.logpdf(z[0], 2.0, scale=s))
    assert np.allclose(lp[:R0], root_ref[:R0])
    loc = 0.9 * z[0, :R0]
    sc_c = np.sqrt(mult[R0:] * (1 - 0.81)); sc_a = np.sqrt(0.5 * mult[R0:] * (1 - 0.81))
    child_ref = np.where(centric[R0:], st.foldnorm.logpdf(z[0, R0:], loc / sc_c, scale=sc_c),
                         st.rice.logpdf(z[0, R0:], loc / sc_a, scale=sc_a))
    assert np.allclose(lp[R0:], child_ref, rtol=1e-9)


def test_adam_matches_closed_form_first_step():
    """[3P] tf_keras Adam: after one step from m=v=0 the update is lr * g/(|g| + eps*sqrt(1-b2)...)."""
    p = {"w": T([1.0, -2.0, 3.0])}
    g = {"w": T([0.5, -4.0, 0.0])}
    state = om.adam_init(p)
    opt = om.AdamConfig(lr=1e-3, beta1=0.9, beta2=0.99, eps=1e-7)
    new = om.adam_apply(p, g, state, opt)
    alpha = 1e-3 * math.sqrt(1 - 0.99) / (1 - 0.9)
    m = np.array([0.05, -0.4, 0.0]); v = np.array([0.0025, 0.16, 0.0])
    assert np.allclose(new["w"].numpy(), np.array([1.0, -2.0, 3.0]) - alpha * m / (np.sqrt(v) + 1e-7))
    # non-finite elements are zeroed before the update (variational.py:208)
    new2 = om.adam_apply(p, {"w": T([np.nan, 1.0, np.inf])}, om.adam_init(p), opt)
    assert np.isfinite(new2["w"].numpy()).all() and new2["w"][0] == 1.0 and new2["w"][2] == 3.0


def test_elbo_gradients_match_finite_differences():
    from careless_b200 import synth
    rng = np.random.default_rng(7)
    p = synth.make_mono(300, 40, d=3, n_images=5, seed=8)
    cfg = om.ModelConfig(n_refl=40, n_meta=3, mlp_width=5, mlp_layers=3, likelihood="studentt", dof=5.0,
                         mc_samples=2, image_scales=True, n_images=5)
    prior = om.PriorData(p["centric"], p["multiplicity"])
    params = om.init_params(cfg, prior)
    params = {k: v + 0.05 * T(rng.standard_normal(tuple(v.shape))) for k, v in params.items()}
    u, e = rng.random((2, 40)), rng.standard_normal((2, 300))
    _, g, _ = om.loss_and_grads(params, p, prior, cfg, u, e)
    f = lambda pp: float(om.forward(pp, p, prior, cfg, u, e)["loss"])
    for k in ("sf_loc_raw", "sf_scale_raw", "mlp.1.kernel", "mlp.out.bias", "image_scales"):
        idx = tuple(int(rng.integers(0, s)) for s in params[k].shape)
        h = 1e-6
        pp = {kk: v.clone() for kk, v in params.items()}
        pp[k][idx] += h; fp = f(pp); pp[k][idx] -= 2 * h; fm = f(pp)
        fd = (fp - fm) / (2 * h)
        assert abs(float(g[k][idx]) - fd) <= 1e-5 * max(1.0, abs(fd)), (k, float(g[k][idx]), fd)


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    out = philox.philox4x32_10(np.uint32(0), np.uint32(0), np.uint32(0), np.uint32(0), 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = philox.philox4x32_10(np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff), np.uint32(0xffffffff),
                               0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = philox.philox4x32_10(np.uint32(0x243f6a88), np.uint32(0x85a308d3), np.uint32(0x13198a2e), np.uint32(0x03707344),
                               0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = philox.refl_uniforms(1234, 0, 2, np.arange(1000))
    assert u.shape == (2, 1000) and u.min() > 0 and u.max() < 1
    e = philox.obs_normals(1234, 3, 1, np.arange(200000))
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1) < 0.01


def test_uniform_grid_is_exact_in_float32_and_never_hits_the_ends():
    """The in-kernel draw u = ((x >> 9) + 0.5) / 2^23 needs 24 significant bits: float32 holds it exactly, so the
    CUDA kernels and this float64 restatement see the same numbers and u can never round to 0 or 1
    (a 24-bit grid would: (2^24 - 0.5) / 2^24 rounds to 1.0f and the inverse CDF returns +inf)."""
    x = np.array([0, 1, 511, 512, 2 ** 31, 2 ** 32 - 512, 2 ** 32 - 1], dtype=np.uint64).astype(np.uint32)
    u = philox.u01(x)
    assert np.array_equal(u.astype(np.float32).astype(np.float64), u)
    assert u.min() == 2.0 ** -24 and u.max() == 1.0 - 2.0 ** -24
    assert np.float32(u.max()) < np.float32(1.0)
    # the oracle sampler clamps injected draws like the kernel does
    z = om.tn_sample(T([1.0]), T([0.5]), T([0.0]), T([1e10]), T([[1.0], [0.0]]))
    assert np.all(np.isfinite(z.numpy()))


def test_oracle_reproduces_committed_golden_vectors():
    """tests/golden/oracle_vectors.npz (made by make_oracle_vectors.py): the checker itself must not drift."""
    import importlib.util, os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_oracle_vectors", os.path.join(here, "golden", "make_oracle_vectors.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    gold = np.load(os.path.join(here, "golden", "oracle_vectors.npz"))
    for name in mod.CASES:
        now = mod.run_case(name)
        for k, v in now.items():
            assert np.allclose(v, gold[k], rtol=1e-9, atol=1e-12), k
