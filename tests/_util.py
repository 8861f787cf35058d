"""Shared helpers of the test-suite: oracle <-> engine parameter plumbing and comparison."""
import numpy as np
import torch

from oracle import model as om
from careless_b200.engine import Engine, EngineConfig


def mlp_names(cfg):
    names = []
    for k in range(cfg.mlp_layers):
        names += [f"mlp.{k}.kernel", f"mlp.{k}.bias"]
    names += ["mlp.out.kernel", "mlp.out.bias"]
    return names


def mlp_flat(params, cfg):
    return np.concatenate([np.asarray(params[n].detach().cpu().numpy(), dtype=np.float64).reshape(-1) for n in mlp_names(cfg)])


def mlp_unflat(flat, like, cfg):
    out, off = {}, 0
    for n in mlp_names(cfg):
        sz = like[n].numel()
        out[n] = torch.as_tensor(np.asarray(flat[off:off + sz], dtype=np.float64).reshape(tuple(like[n].shape)))
        off += sz
    return out


def il_flat(params, cfg):
    return np.concatenate([np.concatenate([params[f"image_layer.{k}.kernel"].detach().numpy().reshape(-1),
                                           params[f"image_layer.{k}.bias"].detach().numpy().reshape(-1)]) for k in range(cfg.image_layers)])


def build(problem, *, mlp_width, mlp_layers, likelihood="normal", dof=None, laue=False, prior="wilson", image_layers=0, refine_uncertainties=False,
          mc_samples=1, kl_weight=None, scale_bijector="exp", scale_shift=None, image_scales=False,
          optimize_dw_r=False, sigma=1.0, opt=None, seed=1234, eps=1e-7, deterministic=False):
    """(oracle cfg, oracle prior, engine) for a synthetic problem dict from careless_b200.synth."""
    R = len(problem["centric"])
    d = problem["metadata"].shape[1]
    n_images = int(problem["n_images"])
    ocfg = om.ModelConfig(n_refl=R, n_meta=d, mlp_width=mlp_width, mlp_layers=mlp_layers, likelihood=likelihood,
                          dof=dof, laue=laue, prior=prior, mc_samples=mc_samples, kl_weight=kl_weight,
                          scale_bijector=scale_bijector, scale_shift=scale_shift, eps=eps,
                          image_scales=image_scales, n_images=n_images, optimize_dw_r=optimize_dw_r, image_layers=image_layers,
                          refine_uncertainties=refine_uncertainties)
    oprior = om.PriorData(centric=problem["centric"], multiplicity=problem["multiplicity"], sigma=sigma,
                          reflids=problem.get("reflids"), root=problem.get("root"), asu_ids=problem.get("asu_id"),
                          r=problem.get("r"))
    opt = opt or om.AdamConfig()
    ecfg = EngineConfig(n_refl=R, n_meta=d, mlp_width=mlp_width, mlp_layers=mlp_layers, n_images=n_images,
                        image_scales=image_scales, image_layers=image_layers, refine_uncertainties=refine_uncertainties, mc_samples=mc_samples, likelihood=likelihood, dof=dof, laue=laue,
                        prior=prior, n_asu=int(problem.get("n_asu", 0)), optimize_dw_r=optimize_dw_r,
                        scale_bijector=scale_bijector, scale_shift=scale_shift, epsilon=eps, kl_weight=kl_weight,
                        learning_rate=opt.lr, beta_1=opt.beta1, beta_2=opt.beta2, adam_epsilon=opt.eps,
                        clipnorm=opt.clipnorm, clipvalue=opt.clipvalue, global_clipnorm=opt.global_clipnorm, seed=seed,
                        deterministic=deterministic)
    eng = Engine(ecfg)
    eng.set_observations(problem["refl_id"], problem["image_id"], problem["metadata"], problem["intensities"],
                         problem["uncertainties"], harmonic_id=problem.get("harmonic_id") if laue else None)
    sig = None if np.isscalar(sigma) and sigma == 1.0 else np.broadcast_to(np.asarray(sigma, dtype=np.float32), (R,))
    eng.set_prior(problem["centric"], problem["multiplicity"], sig, dw_parent=problem.get("dw_parent") if prior == "double_wilson" else None,
                  asu_id=problem.get("asu_id") if prior == "double_wilson" else None,
                  r=problem.get("r") if prior == "double_wilson" else None)
    return ocfg, oprior, eng


def perturbed_params(ocfg, oprior, rng, amount=0.1):
    """Reference initial parameters plus noise, rounded to float32 so both sides start identical."""
    p = om.init_params(ocfg, oprior)
    for k in p:
        noise = amount * rng.standard_normal(tuple(p[k].shape))
        p[k] = torch.as_tensor((p[k].numpy() + noise).astype(np.float32).astype(np.float64))
    return p


def push_params(eng, params, ocfg):
    eng.set_params("sf_loc_raw", params["sf_loc_raw"].numpy())
    eng.set_params("sf_scale_raw", params["sf_scale_raw"].numpy())
    eng.set_params("mlp", mlp_flat(params, ocfg))
    if "image_scales" in params:
        eng.set_params("image_scales", params["image_scales"].numpy())
    if "dw_r_logit" in params:
        eng.set_params("dw_r_logit", params["dw_r_logit"].numpy())
    if ocfg.image_layers > 0:
        eng.set_params("image_layers", il_flat(params, ocfg))
    if "likelihood" in params:
        eng.set_params("likelihood", params["likelihood"].numpy())


def pull_params(eng, like, ocfg):
    out = {"sf_loc_raw": torch.as_tensor(eng.get_params("sf_loc_raw").astype(np.float64)),
           "sf_scale_raw": torch.as_tensor(eng.get_params("sf_scale_raw").astype(np.float64))}
    out.update(mlp_unflat(eng.get_params("mlp"), like, ocfg))
    if "image_scales" in like:
        out["image_scales"] = torch.as_tensor(eng.get_params("image_scales").astype(np.float64))
    if "dw_r_logit" in like:
        out["dw_r_logit"] = torch.as_tensor(eng.get_params("dw_r_logit").astype(np.float64))
    if "likelihood" in like:
        out["likelihood"] = torch.as_tensor(eng.get_params("likelihood").astype(np.float64))
    return out


def engine_grads(eng, ocfg, like):
    g = {"sf_loc_raw": eng.get_grads("sf_loc_raw").astype(np.float64),
         "sf_scale_raw": eng.get_grads("sf_scale_raw").astype(np.float64),
         "mlp": eng.get_grads("mlp").astype(np.float64)}
    if "image_scales" in like:
        g["image_scales"] = eng.get_grads("image_scales").astype(np.float64)
    if "dw_r_logit" in like:
        g["dw_r_logit"] = eng.get_grads("dw_r_logit").astype(np.float64)
    if "likelihood" in like:
        g["likelihood"] = eng.get_grads("likelihood").astype(np.float64)
    if ocfg.image_layers > 0:
        g["image_layers"] = eng.get_grads("image_layers").astype(np.float64)
    return g


def oracle_grads_grouped(g, ocfg):
    out = {"sf_loc_raw": g["sf_loc_raw"].numpy(), "sf_scale_raw": g["sf_scale_raw"].numpy()}
    if all(n in g for n in mlp_names(ocfg)):
        out["mlp"] = mlp_flat(g, ocfg)
    for k in ("image_scales", "dw_r_logit", "likelihood"):
        if k in g:
            out[k] = g[k].numpy()
    if ocfg.image_layers > 0 and "image_layer.0.kernel" in g:
        out["image_layers"] = il_flat(g, ocfg)
    return out


def rel_err_all(a, b):
    """elementwise |a-b| / (|b| + 1e-3*max|b|): relative error with a floor tied to the array scale."""
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    scale = np.max(np.abs(b)) if b.size else 1.0
    return np.abs(a - b) / (np.abs(b) + 1e-3 * scale + 1e-30)


def rel_err(a, b):
    """max over elements of rel_err_all."""
    e = rel_err_all(a, b)
    return float(e.max()) if e.size else 0.0


def rel_err_q(a, b, q=0.995):
    """q-quantile over elements of rel_err_all (robust to the few cancellation-dominated elements)."""
    e = rel_err_all(a, b)
    return float(np.quantile(e, q)) if e.size else 0.0


def rms_err(a, b):
    """||a-b||_2 / ||b||_2"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def load_fixture(name, to_observed=True):
    """tests/golden/fixture_mtz.npz -> careless_b200.io DataSet (same result as read_mtz on the reference's fixture file)."""
    import os
    from careless_b200.io.mtz import DataSet
    from careless_b200.io.symmetry import SpaceGroup, UnitCell
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixture_mtz.npz"))
    keys, types, data = [str(k) for k in z[f"{name}/keys"]], [str(t) for t in z[f"{name}/types"]], z[f"{name}/data"]
    sg = SpaceGroup.from_triplets([str(s) for s in z[f"{name}/symm"]], name=str(z[f"{name}/spacegroup"]))
    cols = {}
    for j, (k, t) in enumerate(zip(keys, types)):
        cols[k] = data[:, j].astype(np.int32) if t in ("H", "B", "Y", "I") else data[:, j].astype(np.float32)
    ds = DataSet(cols, dict(zip(keys, types)), UnitCell(*z[f"{name}/cell"]), sg, merged=False)
    if to_observed:
        ds.set_hkls(sg.hkl_to_observed(ds.get_hkls(), ds["M/ISYM"]))
        del ds.columns["M/ISYM"]
    return ds


def write_stream(path, crystals):
    """A minimal CrystFEL 2.3 stream: `crystals` = list of dicts(cell=(a,b,c nm, al,be,ga), hkl=(n,3) int, I, SigI, fs, ss)."""
    with open(path, "w") as f:
        f.write("CrystFEL stream format 2.3\nGenerated by tests\n----- Begin unit cell -----\n----- End unit cell -----\n")
        for n, c in enumerate(crystals):
            f.write("----- Begin chunk -----\nImage filename: synthetic_%d.h5\nImage serial number: %d\nhit = 1\n" % (n, n + 1))
            f.write("Peaks from peak search\n  fs/px   ss/px (1/d)/nm^-1   Intensity  Panel\n 624.00  259.50       3.55       74.18   p0\nEnd of peak list\n")
            f.write("--- Begin crystal\nCell parameters %.5f %.5f %.5f nm, %.5f %.5f %.5f deg\n" % tuple(c["cell"]))
            f.write("lattice_type = tetragonal\ncentering = P\nnum_reflections = %d\nReflections measured after indexing\n" % len(c["hkl"]))
            f.write("   h    k    l          I   sigma(I)       peak background  fs/px  ss/px panel\n")
            for (h, k, l), i, s, x, y in zip(c["hkl"], c["I"], c["SigI"], c["fs"], c["ss"]):
                f.write("%4d %4d %4d %10.2f %10.2f %10.2f %10.2f %6.1f %6.1f p0\n" % (h, k, l, i, s, 10.0, 1.0, x, y))
            f.write("End of reflections\n--- End crystal\n----- End chunk -----\n")


def synthetic_stream(path, n_crystals=6, n_refl=180, seed=0):
    rng = np.random.default_rng(seed)
    crystals = []
    for _ in range(n_crystals):
        hkl = rng.integers(-12, 13, size=(n_refl, 3))
        hkl = hkl[np.any(hkl != 0, axis=1)]
        i = rng.gamma(2.0, 50.0, size=len(hkl))
        crystals.append(dict(cell=(7.9 + 0.01 * rng.standard_normal(), 7.9, 3.8, 90.0, 90.0, 90.0), hkl=hkl, I=i,
                             SigI=5.0 + 0.1 * np.sqrt(i), fs=rng.uniform(0, 1400, len(hkl)), ss=rng.uniform(0, 1400, len(hkl))))
    write_stream(path, crystals)
    return crystals
