"""A small tour of the C-ABI for compute-sanitizer runs:
    compute-sanitizer --tool memcheck  python tools/sanitizer_case.py
    compute-sanitizer --tool racecheck python tools/sanitizer_case.py
    compute-sanitizer --tool synccheck python tools/sanitizer_case.py
Covers every observation kernel of the default build (k_obs_tc2 with / without image layers, k_obs_tc16 with / without image
layers, the FP32 fallbacks incl. padded width 64), the device-side row preparation (every case: radix sort, padding, gather), the
fused per-reflection kernels, DoubleWilson, Ev11, eval, results, the deterministic mode and the opt-in kernels k_obs_pp / k_obs_tc3."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from careless_b200 import synth
import _util as U

CASES = [
    ("mono_w32", lambda: (synth.make_mono(3000, 300, d=5, n_images=6, seed=1), dict(mlp_width=32, mlp_layers=3, likelihood="studentt", dof=5.0, image_scales=True))),
    ("laue_il_w32", lambda: (synth.make_laue(2500, 300, d=3, n_images=5, seed=2), dict(mlp_width=32, mlp_layers=2, laue=True, image_layers=1, refine_uncertainties=True))),
    ("mono_w10", lambda: (synth.make_mono(3000, 301, d=3, n_images=6, seed=3), dict(mlp_width=10, mlp_layers=4, mc_samples=2))),
    ("mono_il_w10", lambda: (synth.make_mono(3000, 300, d=3, n_images=6, seed=4), dict(mlp_width=10, mlp_layers=3, image_layers=2))),
    ("dw_w8", lambda: (synth.make_double_wilson(800, 100, n_datasets=3, d=3, n_images=4, seed=5), dict(mlp_width=8, mlp_layers=2, prior="double_wilson", optimize_dw_r=True))),
    ("det_w32", lambda: (synth.make_mono(3000, 300, d=4, n_images=6, seed=6), dict(mlp_width=32, mlp_layers=3, deterministic=True))),
    ("det_w10", lambda: (synth.make_mono(3000, 300, d=4, n_images=6, seed=7), dict(mlp_width=10, mlp_layers=3, deterministic=True))),
    ("mono_w64", lambda: (synth.make_mono(2000, 200, d=40, n_images=6, seed=8), dict(mlp_width=48, mlp_layers=2))),
    ("laue_w32", lambda: (synth.make_laue(3000, 300, d=3, n_images=5, seed=9), dict(mlp_width=32, mlp_layers=2, laue=True, likelihood="studentt", dof=6.0))),
]
only = set(sys.argv[1:])
for name, make in CASES:
    if only and name not in only:
        continue
    p, kw = make()
    if name == "mono_il_w10":
        p["image_id"] = np.sort(p["image_id"])
    for env in ({}, {"CLB_PP": "1"}, {"CLB_TC3": "1"}) if name == "mono_w32" else ({},):
        os.environ.update(env)
        _, _, eng = U.build(p, **kw)
        h = eng.step(2)
        m = eng.eval()
        r = eng.get_results(); sm = eng.get_scale_moments()
        print(name, env, h[-1]["loss"], m["NLL"], float(np.sum(r["N"])))
        eng.close()
        for k in env:
            os.environ.pop(k)
