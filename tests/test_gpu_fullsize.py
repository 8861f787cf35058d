"""Parity at BASELINE.json's full size (configs[1]: 10M observations, 500k reflections, StudentT, MLP 32x20)
through size-independent properties, since the float64 oracle cannot finish 10M rows in seconds:

* subsample oracle: the gradient of v_loc / v_scale of a reflection depends only on that reflection's own
  observations, so the oracle evaluated on ~1000 reflections (with all their rows) must reproduce the
  full-size GPU gradient at those indices, and the NLL is additive over reflections;
* permutation invariance: shuffling the input rows (the library sorts them) changes nothing but the
  floating-point summation order.
"""
import numpy as np
import pytest
import torch

from careless_b200 import synth
from careless_b200.engine import Engine, EngineConfig
from oracle import model as om
from oracle import philox

import _util as U

pytestmark = pytest.mark.gpu

N, R, D, W, L = 10_000_000, 500_000, 5, 32, 20
SEED = 1234


def _engine(p, obs_index=None):
    cfg = EngineConfig(n_refl=R, n_meta=D, mlp_width=W, mlp_layers=L, likelihood="studentt", dof=12.0, seed=SEED)
    eng = Engine(cfg)
    eng.set_observations(p["refl_id"], None, p["metadata"], p["intensities"], p["uncertainties"], obs_index=obs_index)
    eng.set_prior(p["centric"], p["multiplicity"])
    return eng


@pytest.fixture(scope="module")
def problem():
    return synth.make_mono(N, R, d=D, n_images=5000, seed=SEED)


def test_fullsize_subsample_oracle_and_permutation(problem):
    p = problem
    rng = np.random.default_rng(0)
    # a non-trivial scale model: identity init + noise
    ocfg = om.ModelConfig(n_refl=R, n_meta=D, mlp_width=W, mlp_layers=L, likelihood="studentt", dof=12.0)
    eng = _engine(p)
    mlp0 = eng.get_params("mlp").astype(np.float64)
    mlp = (mlp0 + 0.02 * rng.standard_normal(mlp0.shape)).astype(np.float32)
    eng.set_params("mlp", mlp)
    hist = eng.step(1)[0]
    g_loc, g_scale, g_mlp = eng.get_grads("sf_loc_raw"), eng.get_grads("sf_scale_raw"), eng.get_grads("mlp")
    assert np.all(np.isfinite(g_loc)) and np.all(np.isfinite(g_mlp)) and np.isfinite(hist["loss"])
    eng.close()

    # ---- subsample oracle on ~1000 reflections with all of their rows (refl_id is sorted in this problem) ----
    sub = np.sort(rng.choice(R, size=1000, replace=False))
    rows = np.nonzero(np.isin(p["refl_id"], sub))[0]
    remap = np.full(R, -1, dtype=np.int64); remap[sub] = np.arange(len(sub))
    sp = {"refl_id": remap[p["refl_id"][rows]], "image_id": p["image_id"][rows], "metadata": p["metadata"][rows],
          "intensities": p["intensities"][rows], "uncertainties": p["uncertainties"][rows]}
    scfg = om.ModelConfig(n_refl=len(sub), n_meta=D, mlp_width=W, mlp_layers=L, likelihood="studentt", dof=12.0)
    sprior = om.PriorData(p["centric"][sub], p["multiplicity"][sub])
    sparams = om.init_params(scfg, sprior)
    like = om.init_params(scfg, sprior)
    sparams.update(U.mlp_unflat(mlp.astype(np.float64), like, scfg))
    u = philox.refl_uniforms(SEED, 0, 1, sub)
    e = philox.obs_normals(SEED, 0, 1, rows)
    _, g, _ = om.loss_and_grads(sparams, sp, sprior, scfg, u, e)
    for name, got in (("sf_loc_raw", g_loc), ("sf_scale_raw", g_scale)):
        ref = g[name].numpy()
        assert U.rms_err(got[sub], ref) <= 1e-4, name
        assert U.rel_err_q(got[sub], ref, 0.99) <= 5e-4, name

    # ---- permutation invariance at full size ----
    perm = rng.permutation(N)
    q = dict(p)
    for k in ("refl_id", "metadata", "intensities", "uncertainties"):
        q[k] = p[k][perm]
    eng2 = _engine(q, obs_index=perm)          # same global row index -> same draws
    eng2.set_params("mlp", mlp)
    hist2 = eng2.step(1)[0]
    for k in ("loss", "NLL", "F KLDiv", "Grad Norm"):
        assert abs(hist2[k] - hist[k]) <= 1e-6 * abs(hist[k]), (k, hist2[k], hist[k])
    assert np.array_equal(eng2.get_grads("sf_loc_raw"), g_loc) or U.rms_err(eng2.get_grads("sf_loc_raw"), g_loc) < 1e-6
    assert U.rms_err(eng2.get_grads("mlp"), g_mlp) < 1e-5
    eng2.close()


def test_fullsize_training_reduces_loss_and_moves_towards_truth(problem):
    """150 default-Adam steps at full size: finite history, decreasing loss, and the posterior locations
    start to correlate with the synthetic truth (they all start at the prior mean)."""
    p = problem
    cfg = EngineConfig(n_refl=R, n_meta=D, mlp_width=W, mlp_layers=L, likelihood="studentt", dof=12.0, seed=SEED)
    eng = Engine(cfg)
    eng.set_observations(p["refl_id"], None, p["metadata"], p["intensities"], p["uncertainties"])
    eng.set_prior(p["centric"], p["multiplicity"])
    seen = np.bincount(p["refl_id"], minlength=R) > 5
    cc0 = np.corrcoef(np.exp(eng.get_params("sf_loc_raw"))[seen], p["f_true"][seen])[0, 1]
    hist = eng.step(150)
    assert len(hist) == 150 and np.all(np.isfinite([h["loss"] for h in hist]))
    assert hist[-1]["loss"] < 0.9 * hist[0]["loss"]
    cc = np.corrcoef(np.exp(eng.get_params("sf_loc_raw"))[seen], p["f_true"][seen])[0, 1]
    print(f"loss {hist[0]['loss']:.4e} -> {hist[-1]['loss']:.4e};  CC(F_loc, F_true) {cc0:.3f} -> {cc:.3f}")
    assert cc > cc0 + 0.05
    eng.close()
