"""CPU restatement of careless's variational merging ELBO step (TEST INFRASTRUCTURE).

Follows, citing ``/root/reference`` file:line:

* ``careless/models/merging/variational.py:123-224``  -- ELBO assembly, MC KL, custom train step
* ``careless/models/merging/surrogate_posteriors.py:45-131`` -- TruncatedNormal surrogate
* ``careless/models/scaling/nn.py:10-120``  -- MLP scale model + NormalLayer
* ``careless/models/scaling/image.py:9-125`` -- image scales / hybrid / image layers
* ``careless/models/likelihoods/mono.py:10-37`` and ``laue.py:9-100`` -- likelihoods
* ``careless/models/priors/wilson.py:13-175`` -- Wilson and DoubleWilson priors
* ``careless/utils/distributions.py:228-348`` -- Rice and FoldedNormal log-densities
* ``careless/io/manager.py:380-507`` -- model construction / initialisation

Third-party arithmetic that is NOT under ``/root/reference`` (tensorflow 2.18,
tensorflow-probability 0.25, tf_keras -- pins in ``pyproject.toml:14-22``) is restated from
the published algorithms and marked [3P]:  ``tfd.TruncatedNormal`` (reparameterised sampler
with the custom gradient of ``_std_samples_with_gradients``; ``log_prob``),
``tfd.Normal/StudentT/HalfNormal/Weibull`` log-densities, ``tfb.AbsoluteValue`` transformed
log-density, ``tf.math.bessel_i0e``, keras ``Dense``/``LeakyReLU``, and the tf_keras
``optimizers.Adam`` update.

Everything is written for torch tensors of a caller-chosen dtype: float64 is the oracle,
float32 (same code) is the timed "port" CPU baseline of bench.py.  Randomness is injected:
``u_f`` in (0,1) of shape (S, R) drives an inverse-CDF truncated-normal sampler, ``eps_s``
~ N(0,1) of shape (S, N) drives the scale sample.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

FLT_TINY = float(np.finfo(np.float32).tiny)
FLT_EPS = float(np.finfo(np.float32).eps)
LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------
# configuration / containers
# ----------------------------------------------------------------------------------------
@dataclass
class ModelConfig:
    n_refl: int
    n_meta: int
    mlp_width: int
    mlp_layers: int
    likelihood: str = "normal"          # 'normal' | 'studentt'   (mono.py:16-37, laue.py:68-100)
    dof: Optional[float] = None
    laue: bool = False
    prior: str = "wilson"               # 'wilson' | 'double_wilson'
    mc_samples: int = 1                 # args/common.py:11-15
    kl_weight: Optional[float] = None   # args/prior.py:7-15
    scale_bijector: str = "exp"         # manager.py:450-463
    scale_shift: Optional[float] = None  # nn.py:84-87 tfb.Shift(scale_multiplier) -- additive
    eps: float = 1e-7                   # args/common.py:38-42
    leakiness: float = 0.01             # nn.py:32
    image_scales: bool = False          # image.py:9-63 (HybridImageScaler)
    n_images: int = 0
    image_layers: int = 0               # image.py:66-125 (NeuralImageScaler)
    refine_uncertainties: bool = False  # Ev11 error model, likelihoods/mono.py:39-73, laue.py:49-65
    optimize_dw_r: bool = False         # wilson.py:105-110
    high: float = 1e10                  # surrogate_posteriors.py:105


@dataclass
class PriorData:
    centric: np.ndarray                 # (R,) bool
    multiplicity: np.ndarray            # (R,) float  (epsilon)
    sigma: np.ndarray | float = 1.0     # Wilson Sigma, manager.py:43-68
    # DoubleWilson (wilson.py:83-138)
    reflids: Optional[np.ndarray] = None   # (R,) int, parent surrogate index, -1 = absent
    root: Optional[np.ndarray] = None      # (R,) bool
    asu_ids: Optional[np.ndarray] = None   # (R,) int
    r: Optional[np.ndarray] = None         # (n_asu,) float


@dataclass
class AdamConfig:
    lr: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.99                 # args/optimizer.py:11-27
    eps: float = 1e-7                   # [3P] tf_keras Adam default epsilon
    clipnorm: Optional[float] = None
    clipvalue: Optional[float] = None
    global_clipnorm: Optional[float] = None


# ----------------------------------------------------------------------------------------
# special functions
# ----------------------------------------------------------------------------------------
def ndtr(x):
    return torch.special.ndtr(x)


def ndtri(p):
    return torch.special.ndtri(p)


def normal_pdf(x):
    return torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


class _StdTruncatedNormal(torch.autograd.Function):
    """Standardised truncated-normal draw e in [alpha, beta] by inverse CDF at injected u.

    Gradient: [3P] TFP ``TruncatedNormal._std_samples_with_gradients`` custom gradient
    (reached from ``surrogate_posteriors.py:50-53``):
        d e / d alpha = exp(0.5 (e^2 - alpha^2)) (1 - u'),   d e / d beta = exp(0.5 (e^2 - beta^2)) u',
    with u' = clip(cdf(e), FLT_MIN, 1 - FLT_EPS).  TFP recomputes u' from the sample; for an
    inverse-CDF sampler that is the injected u up to round-off, which is what is used here.
    """

    @staticmethod
    def forward(ctx, alpha, beta, u):
        u = torch.clamp(u, 2.0 ** -24, 1.0 - 2.0 ** -24)     # draws strictly inside (0,1), as in the CUDA sampler
        Pa = ndtr(alpha)
        Z = ndtr(beta) - Pa
        p = Pa + u * Z
        q = (1.0 - u) * Z + ndtr(-beta)          # 1 - p without cancellation
        e = torch.where(p < 0.5, ndtri(p), -ndtri(q))
        ctx.save_for_backward(alpha, beta, u, e)
        return e

    @staticmethod
    def backward(ctx, dy):
        alpha, beta, u, e = ctx.saved_tensors
        uc = torch.clamp(u, FLT_TINY, 1.0 - FLT_EPS)
        du = torch.exp(0.5 * (e * e - beta * beta) + torch.log(uc))
        dl = torch.exp(0.5 * (e * e - alpha * alpha) + torch.log1p(-uc))
        grad_u = (dy * du).sum(0)
        grad_l = (dy * dl).sum(0)
        return grad_l, grad_u, None


def surrogate_loc_scale(params, cfg: ModelConfig):
    """surrogate_posteriors.py:104-131: loc = Exp(v), scale = Shift(eps)(Exp(v))."""
    loc = torch.exp(params["sf_loc_raw"])
    scale = torch.exp(params["sf_scale_raw"]) + cfg.eps
    return loc, scale


def tn_sample(loc, scale, low, high, u):
    """[3P] tfd.TruncatedNormal._sample_n + surrogate_posteriors.py:50-53 ``maximum(low, s)``."""
    alpha = (low - loc) / scale
    beta = (high - loc) / scale
    e = _StdTruncatedNormal.apply(alpha, beta, u)
    s = e * scale[None, :] + loc[None, :]
    # tf.maximum(low, s): gradient reaches s only where s > low
    return torch.where(s > low[None, :], s, low[None, :].expand_as(s))


def tn_log_prob(z, loc, scale, low, high):
    """[3P] tfd.TruncatedNormal._log_prob."""
    alpha = (low - loc) / scale
    beta = (high - loc) / scale
    logZ = torch.log(ndtr(beta) - ndtr(alpha))
    lp = -(0.5 * ((z - loc) / scale) ** 2 + 0.5 * LOG_2PI + torch.log(scale) + logZ)
    ninf = torch.full_like(lp, -math.inf)
    return torch.where((z > high) | (z < low), ninf, lp)


# ----------------------------------------------------------------------------------------
# priors
# ----------------------------------------------------------------------------------------
def _as(x, like):
    return torch.as_tensor(np.array(x, copy=True), dtype=like.dtype)


def wilson_log_prob(z, centric, eps_mult, sigma):
    """wilson.py:13-57.  centric: HalfNormal(sqrt(eps*Sigma)); acentric: Weibull(2, sqrt(eps*Sigma))."""
    s = torch.sqrt(eps_mult * sigma)
    lp_c = 0.5 * math.log(2.0 / math.pi) - torch.log(s) - 0.5 * (z / s) ** 2
    zz = torch.where(centric, torch.ones_like(z), z)   # unselected branch kept finite
    lp_a = math.log(2.0) + torch.log(zz) - 2.0 * torch.log(s) - (zz / s) ** 2
    return torch.where(centric, lp_c, lp_a)


def wilson_mean_stddev(centric, eps_mult, sigma):
    """[3P] HalfNormal / Weibull(k=2) moments used for initialisation (manager.py:432-433)."""
    s = np.sqrt(np.asarray(eps_mult, dtype=np.float64) * np.asarray(sigma, dtype=np.float64))
    mean_c = s * math.sqrt(2.0 / math.pi)
    std_c = s * math.sqrt(1.0 - 2.0 / math.pi)
    mean_a = s * math.sqrt(math.pi) / 2.0          # lambda * Gamma(1.5)
    std_a = s * math.sqrt(1.0 - math.pi / 4.0)
    c = np.asarray(centric, dtype=bool)
    return np.where(c, mean_c, mean_a), np.where(c, std_c, std_a)


def log_i0(x):
    """distributions.py:260-261: log(bessel_i0e(x)) + |x|."""
    return torch.log(torch.special.i0e(x)) + torch.abs(x)


def rice_log_prob(x, nu, sigma):
    """distributions.py:278-283."""
    return torch.log(x) - 2.0 * torch.log(sigma) - (x * x + nu * nu) / (2.0 * sigma * sigma) \
        + log_i0(x * nu / (sigma * sigma))


def folded_normal_log_prob(x, loc, scale):
    """distributions.py:300-335: Normal(loc, scale) pushed through |.| ([3P] AbsoluteValue bijector:
    log p = logsumexp(log N(x), log N(-x))), NaN for x < 0."""
    a = -0.5 * ((x - loc) / scale) ** 2
    b = -0.5 * ((-x - loc) / scale) ** 2
    lp = torch.logsumexp(torch.stack([a, b]), dim=0) - 0.5 * LOG_2PI - torch.log(scale)
    return torch.where(x < 0, torch.full_like(lp, math.nan), lp)


def prior_log_prob(z, params, prior: PriorData, cfg: ModelConfig):
    centric = torch.as_tensor(np.asarray(prior.centric, dtype=bool))
    mult = _as(prior.multiplicity, z)
    sigma = _as(np.broadcast_to(np.asarray(prior.sigma, dtype=np.float64), np.shape(prior.multiplicity)), z)
    p_wilson = wilson_log_prob(z, centric, mult, sigma)
    if cfg.prior == "wilson":
        return p_wilson
    # DoubleWilson, wilson.py:146-175
    if cfg.optimize_dw_r:
        r_all = torch.sigmoid(params["dw_r_logit"])
    else:
        r_all = _as(prior.r, z)
    asu = torch.as_tensor(np.asarray(prior.asu_ids, dtype=np.int64))
    r = r_all[asu]
    reflids = np.asarray(prior.reflids, dtype=np.int64)
    mask = torch.as_tensor(reflids >= 0)
    safe = torch.as_tensor(np.where(reflids >= 0, reflids, 0))
    z_parent = torch.where(mask[None, :], z[:, safe], torch.zeros_like(z))
    loc = z_parent * r
    r2 = r * r
    scale = torch.where(centric, torch.sqrt(mult * sigma * (1.0 - r2)), torch.sqrt(0.5 * mult * sigma * (1.0 - r2)))
    root = torch.as_tensor(np.asarray(prior.root, dtype=bool))
    zc = torch.where(centric | root, torch.ones_like(z), z)
    p_rice = rice_log_prob(zc, loc, scale)
    p_fold = folded_normal_log_prob(z, loc, scale)
    p_dw = torch.where(centric, p_fold, p_rice)
    return torch.where(root, p_wilson, p_dw)


# ----------------------------------------------------------------------------------------
# scale model
# ----------------------------------------------------------------------------------------
def leaky_relu(x, alpha):
    return torch.where(x > 0, x, alpha * x)


def scale_network(params, data, cfg: ModelConfig, dtype):
    """nn.py:92-120 (+ image.py:116-125 when image layers are present) -> (mu_s, sigma_s, shift)."""
    h = torch.as_tensor(np.asarray(data["metadata"]), dtype=dtype)
    for k in range(cfg.mlp_layers):
        h = leaky_relu(h @ params[f"mlp.{k}.kernel"] + params[f"mlp.{k}.bias"], cfg.leakiness)
    if cfg.image_layers > 0:
        img = torch.as_tensor(np.asarray(data["image_id"], dtype=np.int64))
        for k in range(cfg.image_layers):
            w = params[f"image_layer.{k}.kernel"][img]          # (N, units, in)  image.py:93
            b = params[f"image_layer.{k}.bias"][img]
            h = leaky_relu(torch.einsum("noi,ni->no", w, h) + b, cfg.leakiness)
    out = h @ params["mlp.out.kernel"] + params["mlp.out.bias"]
    mu_s, raw = out[:, 0], out[:, 1]
    if cfg.scale_bijector == "exp":
        sigma_s = torch.exp(raw) + cfg.eps
    elif cfg.scale_bijector == "softplus":
        sigma_s = torch.nn.functional.softplus(raw) + cfg.eps
    else:
        raise ValueError(cfg.scale_bijector)
    shift = 0.0 if cfg.scale_shift is None else float(cfg.scale_shift)
    return mu_s, sigma_s, shift


def image_scale_vector(params, data, cfg: ModelConfig, dtype):
    """image.py:23-42: scales = concat([1], _scales)[image_id]."""
    if not cfg.image_scales:
        return None
    one = torch.ones(1, dtype=dtype)
    w = torch.cat([one, params["image_scales"]])
    return w[torch.as_tensor(np.asarray(data["image_id"], dtype=np.int64))]


# ----------------------------------------------------------------------------------------
# likelihoods
# ----------------------------------------------------------------------------------------
def normal_log_prob(x, loc, scale):
    return -0.5 * ((x - loc) / scale) ** 2 - torch.log(scale) - 0.5 * LOG_2PI


def studentt_log_prob(x, dof, loc, scale):
    """[3P] tfd.StudentT._log_prob."""
    y = (x - loc) / scale
    return (-0.5 * (dof + 1.0) * torch.log1p(y * y / dof) - torch.log(scale) - 0.5 * math.log(dof)
            - 0.5 * math.log(math.pi) - math.lgamma(0.5 * dof) + math.lgamma(0.5 * (dof + 1.0)))


def laue_convolve(ipred, harmonic_id):
    """laue.py:17-25: scatter_nd(harmonic_id, ipred^T, shape) -- sum rows per spot, 0 elsewhere."""
    hid = torch.as_tensor(np.asarray(harmonic_id, dtype=np.int64))
    out = torch.zeros_like(ipred)
    return out.index_add(1, hid, ipred)


def ev11_sigma(x, sig, lik_raw):
    """mono.py:52-59 corrected_sigiobs: Sdfac sqrt(sigma^2 + SdB softplus(x) + Sdadd softplus(x)^2) with
    (Sdfac, Sdadd, SdB) = softplus(raw) (TransformedVariable(1., Softplus), mono.py:42-44)."""
    sdfac, sdadd, sdb = torch.nn.functional.softplus(lik_raw)
    p = torch.nn.functional.softplus(x)
    return sdfac * torch.sqrt(sig * sig + sdb * p + sdadd * p * p)


def likelihood_log_prob(ipred, data, cfg: ModelConfig, params=None):
    dtype = ipred.dtype
    iobs = torch.as_tensor(np.asarray(data["intensities"]), dtype=dtype)
    sig = torch.as_tensor(np.asarray(data["uncertainties"]), dtype=dtype)
    x = laue_convolve(ipred, data["harmonic_id"]) if cfg.laue else ipred
    if cfg.refine_uncertainties:
        sig = ev11_sigma(x, sig, params["likelihood"])
    if cfg.likelihood == "normal":
        return normal_log_prob(x, iobs, sig)
    if cfg.likelihood == "studentt":
        return studentt_log_prob(x, float(cfg.dof), iobs, sig)
    raise ValueError(cfg.likelihood)


# ----------------------------------------------------------------------------------------
# the ELBO
# ----------------------------------------------------------------------------------------
def forward(params, data, prior: PriorData, cfg: ModelConfig, u_f, eps_s):
    """variational.py:141-183.  Returns dict(loss, nll, kl, ipred, z_f, z_scale)."""
    dtype = params["sf_loc_raw"].dtype
    S = cfg.mc_samples
    u_f = torch.as_tensor(np.asarray(u_f), dtype=dtype).reshape(S, -1)
    eps_s = torch.as_tensor(np.asarray(eps_s), dtype=dtype).reshape(S, -1)
    centric = np.asarray(prior.centric, dtype=bool)
    low = torch.as_tensor((1e-32 * ~centric).astype(np.float32), dtype=dtype)   # manager.py:434
    high = torch.full_like(low, cfg.high)

    loc, scale = surrogate_loc_scale(params, cfg)
    z_f = tn_sample(loc, scale, low, high, u_f)                                   # :154

    mu_s, sigma_s, shift = scale_network(params, data, cfg, dtype)                # :156
    z_scale = mu_s[None, :] + sigma_s[None, :] * eps_s + shift                    # :157
    a = image_scale_vector(params, data, cfg, dtype)
    if a is not None:
        z_scale = z_scale * a[None, :]                                            # image.py:58-63

    refl_id = torch.as_tensor(np.asarray(data["refl_id"], dtype=np.int64))
    ipred = z_scale * z_f[:, refl_id] ** 2                                        # :167
    ll = likelihood_log_prob(ipred, data, cfg, params)                                    # :169-171

    kl_terms = tn_log_prob(z_f, loc, scale, low, high) - prior_log_prob(z_f, params, prior, cfg)  # :123-128
    if cfg.kl_weight is None:                                                     # :172-174
        kl = kl_terms.sum() / S
        kl_loss = kl
        ll_red = ll.sum() / S
    else:                                                                         # :175-177
        kl = kl_terms.mean()
        kl_loss = cfg.kl_weight * kl
        ll_red = ll.mean()
    loss = kl_loss - ll_red                                                       # :180 + add_loss
    return dict(loss=loss, nll=-ll_red, kl=kl, ipred=ipred, z_f=z_f, z_scale=z_scale)


def trainable_names(params, frozen=()):
    return [k for k in params if not any(k == f or k.startswith(f + ".") for f in frozen)]


def loss_and_grads(params, data, prior, cfg, u_f, eps_s, frozen=()):
    """variational.py:197-205: loss, grads and the (pre-filter) global gradient norm."""
    names = trainable_names(params, frozen)
    leaves = {k: (v.detach().clone().requires_grad_(k in names)) for k, v in params.items()}
    out = forward(leaves, data, prior, cfg, u_f, eps_s)
    grads = torch.autograd.grad(out["loss"], [leaves[k] for k in names], allow_unused=True)
    g = {k: (torch.zeros_like(leaves[k]) if gi is None else gi) for k, gi in zip(names, grads)}
    gn = math.sqrt(sum(float((gi.double() ** 2).sum()) for gi in g.values()))
    metrics = {"loss": float(out["loss"].detach()), "NLL": float(out["nll"].detach()),
               "F KLDiv": float(out["kl"].detach()), "Grad Norm": gn}
    return metrics, g, out


def adam_init(params):
    return {"t": 0, "m": {k: torch.zeros_like(v) for k, v in params.items()},
            "v": {k: torch.zeros_like(v) for k, v in params.items()}}


def adam_apply(params, grads, state, opt: AdamConfig):
    """variational.py:208-209 + [3P] tf_keras optimizers.Adam.update_step (epsilon outside sqrt):
        alpha = lr sqrt(1-b2^t)/(1-b1^t);  m += (g-m)(1-b1);  v += (g^2-v)(1-b2);  x -= alpha m/(sqrt(v)+eps)
    Non-finite gradient elements are zeroed first (:208); optional clipping in keras order."""
    g = {k: torch.where(torch.isfinite(v), v, torch.zeros_like(v)) for k, v in grads.items()}
    if opt.clipnorm is not None:        # per-variable tf.clip_by_norm
        for k in g:
            # "likelihood" packs three scalar keras variables (Sdfac, Sdadd, SdB): each is clipped on its own
            n = torch.abs(g[k]) if k == "likelihood" else torch.sqrt((g[k] ** 2).sum())
            g[k] = g[k] * opt.clipnorm / torch.clamp(n, min=opt.clipnorm)
    if opt.global_clipnorm is not None:  # tf.clip_by_global_norm
        n = torch.sqrt(sum((v ** 2).sum() for v in g.values()))
        for k in g:
            g[k] = g[k] * opt.global_clipnorm / torch.clamp(n, min=opt.global_clipnorm)
    if opt.clipvalue is not None:
        for k in g:
            g[k] = torch.clamp(g[k], -opt.clipvalue, opt.clipvalue)
    state["t"] += 1
    t = state["t"]
    alpha = opt.lr * math.sqrt(1.0 - opt.beta2 ** t) / (1.0 - opt.beta1 ** t)
    new = dict(params)
    for k in g:
        m, v = state["m"][k], state["v"][k]
        m += (g[k] - m) * (1.0 - opt.beta1)
        v += (g[k] * g[k] - v) * (1.0 - opt.beta2)
        new[k] = params[k] - alpha * m / (torch.sqrt(v) + opt.eps)
    return new


def train(params, data, prior, cfg, opt: AdamConfig, draws, frozen=()):
    """variational.py:226-275 for len(draws) steps; draws = [(u_f, eps_s), ...]."""
    state = adam_init(params)
    history = []
    for u_f, eps_s in draws:
        metrics, g, _ = loss_and_grads(params, data, prior, cfg, u_f, eps_s, frozen)
        history.append(metrics)
        if not math.isfinite(metrics["Grad Norm"]):
            break
        params = adam_apply(params, g, state, opt)
    return params, history, state


# ----------------------------------------------------------------------------------------
# construction (manager.py:380-507) and posterior moments (manager.py:164-250)
# ----------------------------------------------------------------------------------------
def init_params(cfg: ModelConfig, prior: PriorData, dtype=torch.float64, init_scale=1.0):
    mean, std = wilson_mean_stddev(prior.centric, prior.multiplicity,
                                   np.broadcast_to(np.asarray(prior.sigma, dtype=np.float64), np.shape(prior.multiplicity)))
    loc = mean.astype(np.float32).astype(np.float64)
    scale = (std * init_scale).astype(np.float32).astype(np.float64)
    p = {}
    # TransformedVariable stores the inverse-bijected value (float32 in the reference)
    p["sf_loc_raw"] = torch.as_tensor(np.log(loc).astype(np.float32), dtype=dtype)
    p["sf_scale_raw"] = torch.as_tensor(np.log(scale - cfg.eps).astype(np.float32), dtype=dtype)
    fan_in = cfg.n_meta
    for k in range(cfg.mlp_layers):                                  # nn.py:55-68 identity init
        p[f"mlp.{k}.kernel"] = torch.eye(fan_in, cfg.mlp_width, dtype=dtype)
        p[f"mlp.{k}.bias"] = torch.zeros(cfg.mlp_width, dtype=dtype)
        fan_in = cfg.mlp_width
    for k in range(cfg.image_layers):                                # image.py:73-88
        p[f"image_layer.{k}.kernel"] = torch.eye(cfg.mlp_width, fan_in, dtype=dtype).repeat(cfg.n_images, 1, 1)
        p[f"image_layer.{k}.bias"] = torch.zeros(cfg.n_images, cfg.mlp_width, dtype=dtype)
        fan_in = cfg.mlp_width
    p["mlp.out.kernel"] = torch.eye(fan_in, 2, dtype=dtype)          # nn.py:72-79
    p["mlp.out.bias"] = torch.zeros(2, dtype=dtype)
    if cfg.image_scales:                                             # image.py:21
        p["image_scales"] = torch.ones(cfg.n_images - 1, dtype=dtype)
    if cfg.prior == "double_wilson" and cfg.optimize_dw_r:           # wilson.py:105-110 (Sigmoid bijector)
        r = np.asarray(prior.r, dtype=np.float64)
        with np.errstate(divide="ignore"):
            p["dw_r_logit"] = torch.as_tensor(np.log(r) - np.log1p(-r), dtype=dtype)
    if cfg.refine_uncertainties:                                     # mono.py:42-44: softplus^-1(1) three times
        p["likelihood"] = torch.full((3,), float(np.float32(math.log(math.e - 1.0))), dtype=dtype)
    return p


def tn_moments(loc, scale, low, high=np.inf):
    """Mean, stddev and 4th raw moment of TruncatedNormal(loc, scale, low, high) via scipy
    (surrogate_posteriors.py:74-102; manager.py:188-197)."""
    from scipy.stats import truncnorm
    loc = np.asarray(loc, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    a, b = (low - loc) / scale, (high - loc) / scale
    mean = truncnorm.mean(a, b, loc, scale)
    std = truncnorm.std(a, b, loc, scale)
    m4 = truncnorm.moment(4, a, b, loc, scale)
    return mean, std, m4
