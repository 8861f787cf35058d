#!/usr/bin/env python
"""Regenerates tests/golden/oracle_vectors.npz: float64 oracle outputs (ELBO terms, per-group gradient norms and leading
elements, parameters after 3 Adam steps) for three seeded problems from careless_b200.synth.

    python tests/golden/make_oracle_vectors.py

The inputs are not stored: careless_b200.synth is deterministic for a seed (numpy PCG64), the draws come from
oracle/philox.py.  tests/test_oracle_kat.py checks that today's oracle still reproduces these numbers (a guard against
drift of the checker itself); tests/test_gpu_parity.py compares the CUDA engine with them without running the oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from careless_b200 import synth  # noqa: E402
from oracle import model as om, philox  # noqa: E402

CASES = {
    "mono_t_w32": dict(gen=("mono", 4000, 500, 4, 9, 101), cfg=dict(mlp_width=32, mlp_layers=4, likelihood="studentt", dof=6.0)),
    "laue_n_w10_hybrid": dict(gen=("laue", 3000, 400, 3, 7, 102), cfg=dict(mlp_width=10, mlp_layers=5, laue=True, image_scales=True, mc_samples=2)),
    "dw_w8": dict(gen=("dw", 800, 150, 3, 4, 103), cfg=dict(mlp_width=8, mlp_layers=3, prior="double_wilson", optimize_dw_r=True)),
}
SEED = 20240611


def problem(gen):
    kind, n, r, d, n_img, seed = gen
    if kind == "mono":
        return synth.make_mono(n, r, d=d, n_images=n_img, seed=seed)
    if kind == "laue":
        return synth.make_laue(n, r, d=d, n_images=n_img, seed=seed)
    return synth.make_double_wilson(n, r, n_datasets=3, d=d, n_images=n_img, r=0.9, seed=seed)


def run_case(name):
    c = CASES[name]
    p = problem(c["gen"])
    R, N = len(p["centric"]), len(p["refl_id"])
    cfg = om.ModelConfig(n_refl=R, n_meta=p["metadata"].shape[1], n_images=int(p["n_images"]), **c["cfg"])
    prior = om.PriorData(centric=p["centric"], multiplicity=p["multiplicity"], reflids=p.get("reflids"), root=p.get("root"),
                         asu_ids=p.get("asu_id"), r=p.get("r"))
    params = om.init_params(cfg, prior)
    S = cfg.mc_samples
    draws = [(philox.refl_uniforms(SEED, s, S, np.arange(R)), philox.obs_normals(SEED, s, S, np.arange(N))) for s in range(3)]
    metrics, g, _ = om.loss_and_grads(params, p, prior, cfg, *draws[0])
    out = {f"{name}/metrics": np.array([metrics[k] for k in ("loss", "NLL", "F KLDiv", "Grad Norm")])}
    for k, v in g.items():
        v = v.numpy().reshape(-1)
        out[f"{name}/grad_norm/{k}"] = np.array([np.linalg.norm(v)])
        out[f"{name}/grad_head/{k}"] = v[:8].copy()
    final, hist, _ = om.train(params, p, prior, cfg, om.AdamConfig(lr=1e-2), draws)
    out[f"{name}/loss_history"] = np.array([h["loss"] for h in hist])
    out[f"{name}/sf_loc_raw_final"] = final["sf_loc_raw"].numpy()
    return out


if __name__ == "__main__":
    allv = {}
    for name in CASES:
        allv.update(run_case(name))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_vectors.npz"), **allv)
    print("wrote oracle_vectors.npz with", len(allv), "arrays")
