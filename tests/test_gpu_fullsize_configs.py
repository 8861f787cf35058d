"""Parity at BASELINE.json's full sizes for configs[2] (Laue, 20 M rows), configs[3] (4-dataset DoubleWilson, 40 M observations)
and one GPU's share of configs[4] (stills: 25 M observations, 250 k reflections, 12 500 images, MLP 10x20 + 2 per-image
layers), through the same size-independent property tests/test_gpu_fullsize.py uses for configs[1]:

    the gradient of (v_loc, v_scale) of a reflection depends only on the rows of its own partition group -- the central ray
    for Laue, the root ancestor's family for DoubleWilson, the reflection itself otherwise -- so the float64 oracle evaluated
    on a few hundred groups WITH ALL THEIR ROWS must reproduce the full-size CUDA gradient at those indices.

The sub-problem is cut out with the product's own partitioner (`parallel.shard` with a two-"rank" assignment: selected groups
vs the rest), which also hands back the global row / reflection indices that key the Philox draws."""
import argparse

import numpy as np
import pytest
import torch

import bench
from careless_b200 import parallel, synth
from careless_b200.engine import Engine, EngineConfig
from oracle import model as om
from oracle import philox

import _util as U

pytestmark = pytest.mark.gpu
SEED, D = 1234, 5


def _build(name, obs, refl):
    c = dict(bench.CONFIGS[name])
    c["n_images"] = max(2, int(c["n_images"] * obs / c["obs"]))
    c["obs"], c["refl"], c["scaling"] = obs, refl, "strong"
    args = argparse.Namespace(config=name, obs=obs, refl=refl)
    li, lt, ex = bench.build_problem(args, c, 0, 1)
    return c, li, lt, ex


def _engine(name, c, li, lt, ex):
    use_img = c["image_layers"] > 0
    cfg = EngineConfig(n_refl=len(lt["refl_index"]), n_meta=D, mlp_width=c["width"], mlp_layers=c["layers"], likelihood=c["likelihood"],
                       dof=c["dof"], laue=(name == "laue"), n_images=c["n_images"] if use_img else 0, image_layers=c["image_layers"],
                       prior="double_wilson" if name == "dw" else "wilson", n_asu=4 if name == "dw" else 0, seed=SEED)
    eng = Engine(cfg)
    eng.set_observations(li["refl_id"], li.get("image_id") if use_img else None, li["metadata"], li["intensities"], li["uncertainties"],
                         harmonic_id=li.get("harmonic_id"))
    eng.set_prior(lt["centric"], lt["multiplicity"], None, dw_parent=lt.get("dw_parent"), asu_id=lt.get("asu_id"), r=ex["r"])
    return eng


def _check(name, obs, refl, n_groups):
    c, li, lt, ex = _build(name, obs, refl)
    R = len(lt["refl_index"])
    rng = np.random.default_rng(0)
    eng = _engine(name, c, li, lt, ex)
    mlp = (eng.get_params("mlp").astype(np.float64) + 0.02 * rng.standard_normal(eng.group_size("mlp"))).astype(np.float32)
    eng.set_params("mlp", mlp)
    il = None
    if c["image_layers"]:
        il = (eng.get_params("image_layers").astype(np.float64) + 0.02 * rng.standard_normal(eng.group_size("image_layers"))).astype(np.float32)
        eng.set_params("image_layers", il)
    hist = eng.step(1)[0]
    g_loc, g_scale = eng.get_grads("sf_loc_raw"), eng.get_grads("sf_scale_raw")
    assert np.isfinite(hist["loss"]) and np.all(np.isfinite(g_loc)) and np.all(np.isfinite(g_scale))
    eng.close()

    # ---- the partition groups of this configuration, a random handful of them, and everything that belongs to them ----
    inputs = {"refl_id": li["refl_id"], "image_id": li.get("image_id"), "metadata": li["metadata"], "intensities": li["intensities"],
              "uncertainties": li["uncertainties"], "harmonic_id": li.get("harmonic_id")}
    tables = {"centric": lt["centric"], "multiplicity": lt["multiplicity"], "dw_parent": lt.get("dw_parent"), "asu_id": lt.get("asu_id")}
    groups = parallel.reflection_groups(R, li["refl_id"], li.get("harmonic_id"), lt.get("dw_parent"))
    chosen = rng.choice(np.unique(groups), size=n_groups, replace=False)
    ranks = np.where(np.isin(groups, chosen), 0, 1)
    si, st = parallel.shard(inputs, tables, ranks, 0, laue=(name == "laue"))
    sub = st["refl_index"]
    scfg = om.ModelConfig(n_refl=len(sub), n_meta=D, mlp_width=c["width"], mlp_layers=c["layers"], likelihood=c["likelihood"], dof=c["dof"],
                          laue=(name == "laue"), prior="double_wilson" if name == "dw" else "wilson", n_images=c["n_images"],
                          image_layers=c["image_layers"])
    if name == "dw":
        par = st["dw_parent"].astype(np.int64)
        sprior = om.PriorData(st["centric"], st["multiplicity"], reflids=np.where(par >= 0, par, np.arange(len(sub))), root=(par == -2),
                              asu_ids=st["asu_id"], r=ex["r"])
    else:
        sprior = om.PriorData(st["centric"], st["multiplicity"])
    sparams = om.init_params(scfg, sprior)
    sparams.update(U.mlp_unflat(mlp.astype(np.float64), om.init_params(scfg, sprior), scfg))
    if il is not None:
        w, n_img, off = c["width"], c["n_images"], 0
        for k in range(c["image_layers"]):
            sparams[f"image_layer.{k}.kernel"] = torch.as_tensor(il[off:off + n_img * w * w].astype(np.float64).reshape(n_img, w, w)); off += n_img * w * w
            sparams[f"image_layer.{k}.bias"] = torch.as_tensor(il[off:off + n_img * w].astype(np.float64).reshape(n_img, w)); off += n_img * w
    u = philox.refl_uniforms(SEED, 0, 1, sub)
    e = philox.obs_normals(SEED, 0, 1, si["obs_index"])
    si["n_images"] = c["n_images"]
    _, g, _ = om.loss_and_grads(sparams, si, sprior, scfg, u, e)
    print(f"[{name}] {len(sub)} reflections / {len(si['refl_id'])} rows in the oracle sub-problem; full-size loss {hist['loss']:.6e}")
    for nm, got in (("sf_loc_raw", g_loc), ("sf_scale_raw", g_scale)):
        ref = g[nm].numpy()
        assert U.rms_err(got[sub], ref) <= 1e-4, (name, nm, U.rms_err(got[sub], ref))
        assert U.rel_err_q(got[sub], ref, 0.99) <= 5e-4, (name, nm)


def test_laue_20M_rows_subsample_oracle():
    _check("laue", 20_000_000, 1_000_000, n_groups=300)


def test_double_wilson_40M_observations_subsample_oracle():
    _check("dw", 40_000_000, 2_000_000, n_groups=300)


def test_stills_share_25M_observations_image_layers_subsample_oracle():
    _check("stills", 25_000_000, 250_000, n_groups=600)
