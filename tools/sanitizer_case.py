"""A small tour of the C-ABI for compute-sanitizer runs:  compute-sanitizer --tool memcheck python tools/sanitizer_case.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from careless_b200 import synth
import _util as U
for kind in ("mono", "laue_il"):
    if kind == "mono":
        p = synth.make_mono(3000, 300, d=5, n_images=6, seed=1)
        _, _, eng = U.build(p, mlp_width=32, mlp_layers=3, likelihood="studentt", dof=5.0, image_scales=True)
    else:
        p = synth.make_laue(2500, 300, d=3, n_images=5, seed=2)
        _, _, eng = U.build(p, mlp_width=32, mlp_layers=2, laue=True, image_layers=1, refine_uncertainties=True)
    h = eng.step(2)
    m = eng.eval()
    r = eng.get_results(); sm = eng.get_scale_moments()
    print(kind, h[-1]["loss"], m["NLL"], float(np.sum(r["N"])))
    eng.close()
