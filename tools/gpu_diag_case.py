"""Print the worst gradient elements of one parity case (GPU vs f64 oracle vs f32 oracle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from careless_b200 import synth
from oracle import model as om
import _util as U

p = synth.make_laue(4000, 500, d=3, n_images=20, seed=8)
kw = dict(mlp_width=8, mlp_layers=3, laue=True, likelihood="studentt", dof=4.0, mc_samples=2)
rng = np.random.default_rng(7)
ocfg, oprior, eng = U.build(p, **kw)
params = U.perturbed_params(ocfg, oprior, rng)
U.push_params(eng, params, ocfg)
S, R, N = 2, 500, 4000
u = rng.random((1, S, R)); u = (np.floor(u * 2 ** 23) + 0.5) / 2 ** 23
e = rng.standard_normal((1, S, N)).astype(np.float32).astype(np.float64)
hist = eng.step(1, u_f=u, eps_s=e)
m, g, out = om.loss_and_grads(params, p, oprior, ocfg, u[0], e[0])
p32 = {k: v.float() for k, v in params.items()}
_, g32, out32 = om.loss_and_grads(p32, p, oprior, ocfg, u[0], e[0])
ge = eng.get_grads("sf_loc_raw").astype(np.float64)
go = g["sf_loc_raw"].numpy(); g3 = g32["sf_loc_raw"].double().numpy()
scale = np.abs(go).max()
err = np.abs(ge - go) / (np.abs(go) + 1e-3 * scale)
err3 = np.abs(g3 - go) / (np.abs(go) + 1e-3 * scale)
cnt = np.bincount(p["refl_id"], minlength=R)
z = eng.get_samples(); zo = out["z_f"].detach().numpy()
print("scale", scale)
for i in np.argsort(-err)[:12]:
    print(i, f"gpu={ge[i]:.6e} f64={go[i]:.6e} f32={g3[i]:.6e} err={err[i]:.2e} err32={err3[i]:.2e} centric={p['centric'][i]} nobs={cnt[i]}",
          f"z_gpu={z[:, i]} z_ref={zo[:, i]} loc={float(torch.exp(params['sf_loc_raw'][i])):.4f} scale={float(torch.exp(params['sf_scale_raw'][i])):.4f}")
