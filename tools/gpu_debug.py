"""Stage-by-stage GPU bring-up script (prints with timestamps, unbuffered)."""
import faulthandler, os, sys, time
faulthandler.enable()
faulthandler.dump_traceback_later(40, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.time()
def log(*a):
    print(f"[{time.time()-t0:7.2f}s]", *a, flush=True)
import numpy as np
log("numpy ok")
from careless_b200 import synth
from careless_b200.engine import Engine, EngineConfig
log("package ok")
p = synth.make_mono(3000, 400, d=3, n_images=11, seed=1)
cfg = EngineConfig(n_refl=400, n_meta=3, mlp_width=8, mlp_layers=3)
eng = Engine(cfg); log("engine created")
eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"]); log("obs set")
eng.set_prior(p["centric"], p["multiplicity"]); log("prior set")
eng.step_begin(); eng.synchronize(); log("step_begin done")
eng.step_norms(); eng.synchronize(); log("step_norms done")
m = eng.step_end(True); log("step_end done", m)
h = eng.step(3); log("3 steps", h[-1])
import torch
log("torch imported", torch.__version__, torch.get_num_threads())
from oracle import model as om
ocfg = om.ModelConfig(n_refl=400, n_meta=3, mlp_width=8, mlp_layers=3)
pr = om.PriorData(p["centric"], p["multiplicity"])
params = om.init_params(ocfg, pr)
rng = np.random.default_rng(0)
u = rng.random((1, 400)); e = rng.standard_normal((1, 3000))
met, g, out = om.loss_and_grads(params, p, pr, ocfg, u, e); log("oracle step", met)
eng.close(); log("closed")
