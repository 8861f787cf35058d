"""Find where the first non-finite value appears in a full-size run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from careless_b200 import synth
from careless_b200.engine import Engine, EngineConfig
N, R = int(os.environ.get("N", 10_000_000)), int(os.environ.get("R", 500_000))
p = synth.make_mono(N, R, d=5, n_images=5000, seed=1234)
cfg = EngineConfig(n_refl=R, n_meta=5, mlp_width=32, mlp_layers=20, likelihood="studentt", dof=12.0, seed=1234)
eng = Engine(cfg)
eng.set_observations(p["refl_id"], None, p["metadata"], p["intensities"], p["uncertainties"])
eng.set_prior(p["centric"], p["multiplicity"])
eng.enable_ipred(True)
for step in range(4):
    h = eng.step(1)
    print("step", step, h)
    z = eng.get_samples(); ip = eng.get_ipred()
    for name, arr in (("z", z), ("ipred", ip), ("g_loc", eng.get_grads("sf_loc_raw")), ("g_scale", eng.get_grads("sf_scale_raw")),
                      ("g_mlp", eng.get_grads("mlp")), ("loc_raw", eng.get_params("sf_loc_raw")), ("scale_raw", eng.get_params("sf_scale_raw")),
                      ("mlp", eng.get_params("mlp"))):
        bad = ~np.isfinite(arr)
        print(f"   {name:10s} nonfinite={int(bad.sum()):8d} min={np.nanmin(arr):.4e} max={np.nanmax(arr):.4e}", end="")
        if bad.any():
            idx = np.argwhere(bad)[:3].tolist()
            print("  first bad idx", idx, end="")
        print()
    if not h or not np.isfinite(h[0]["loss"]):
        ib = np.argwhere(~np.isfinite(ip))[:5]
        for s_, i_ in ib:
            r = p["refl_id"][i_]
            print("   bad obs", i_, "refl", r, "z", z[:, r], "I", p["intensities"][i_], "sig", p["uncertainties"][i_], "meta", p["metadata"][i_],
                  "loc_raw", eng.get_params("sf_loc_raw")[r], "scale_raw", eng.get_params("sf_scale_raw")[r])
        break
