#!/usr/bin/env python
"""make_tf_vectors.py -- golden vectors from the REAL reference (TensorFlow careless) with injected draws.

    python tests/golden/make_tf_vectors.py [--reference /root/reference | baseline/_ref] [--out tests/golden/tf_vectors.npz]

STATUS: this container and the GPU boxes have no tensorflow / tensorflow_probability / tf_keras (and no network), so
the script has NOT been run here; `tests/golden/tf_vectors.npz` does not exist and every ELBO / gradient / Adam-trajectory
comparison in this repository is "parity unpinned" (CUDA vs the repo's own float64 oracle).  Run this script on any
machine that has the reference's pinned stack (tensorflow==2.18.0, tensorflow-probability[tf]==0.25, tf_keras; see
/root/reference/pyproject.toml:17-18) and commit the .npz: tests/test_tf_vectors.py then pins the oracle (CPU) and the
CUDA engine (GPU) to the reference's own numbers at rtol 1e-4.

What it does:
* builds the reference VariationalMergingModel (careless/models/merging/variational.py:11-45) for a few small seeded
  problems of careless_b200.synth -- WilsonPrior, TruncatedNormal surrogate initialised as io/manager.py:432-436 does,
  MLPScaler with the CLI's exp bijector, Normal / StudentT likelihood;
* replaces the two samplers of the hot path by functions that return INJECTED draws:
    - the standardised truncated-normal rejection sampler behind tfd.TruncatedNormal._sample_n (it cannot be driven
      externally) by the inverse-CDF transform e = ndtri(ndtr(a) + u (ndtr(b) - ndtr(a))) of injected uniforms u
      (evaluated in float64, returned in float32) -- TFP's custom gradient of the sampler is kept as is;
    - the N(0,1) sampler behind tfd.Normal._sample_n by the injected eps;
* runs the reference's own train step arithmetic (variational.py:185-224) eagerly: loss / NLL / KL / gradient norm, every
  per-variable gradient of step 1, then n_steps of tf_keras Adam (io/manager.py:494-501 defaults) and the merged
  F = surrogate.mean(), SigF = surrogate.stddev() (io/manager.py:188-197).
Everything needed to replay a case (inputs, prior tables, draws, initial parameters) is stored next to the results.
"""
import argparse
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {
    # name: (problem kwargs, model kwargs)
    "mono_normal_w8l3": (dict(N=3000, R=400, d=3, n_images=11, seed=1), dict(width=8, layers=3, likelihood="normal", dof=None, S=1)),
    "mono_studentt_w10l20": (dict(N=5000, R=700, d=2, n_images=23, seed=2), dict(width=10, layers=20, likelihood="studentt", dof=12.0, S=2)),
    "mono_studentt_w32l20": (dict(N=9000, R=900, d=5, n_images=30, seed=3), dict(width=32, layers=20, likelihood="studentt", dof=12.0, S=1)),
}
N_STEPS = 5


def draws(rng, S, R, N, n_steps):
    u = (np.floor(rng.random((n_steps, S, R)) * 2 ** 23) + 0.5) / 2 ** 23      # on the float32-exact grid, never 0 or 1
    e = rng.standard_normal((n_steps, S, N)).astype(np.float32)
    return u.astype(np.float64), e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=None, help="directory that contains the `careless` package (default: baseline/_ref, then /root/reference)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "tf_vectors.npz"))
    args = ap.parse_args()
    for cand in (args.reference, os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "careless")):
            sys.path.insert(0, cand)
            break
    else:
        raise SystemExit("no `careless` package found: pass --reference")
    import importlib.util
    if importlib.util.find_spec("reciprocalspaceship") is None:        # imported by priors/wilson.py, not used by WilsonPrior
        sys.modules["reciprocalspaceship"] = types.ModuleType("reciprocalspaceship")
    import tensorflow as tf
    import tensorflow_probability as tfp
    import tf_keras as tfk
    from scipy.special import ndtr, ndtri
    from tensorflow_probability import bijectors as tfb
    tf.config.set_visible_devices([], "GPU")          # the CPU path is the FP32-exact one (TF32 matmuls on GPUs)
    tf.config.run_functions_eagerly(True)
    from careless.models.likelihoods.mono import NormalLikelihood, StudentTLikelihood
    from careless.models.merging.surrogate_posteriors import TruncatedNormal
    from careless.models.merging.variational import VariationalMergingModel
    from careless.models.priors.wilson import WilsonPrior
    from careless.models.scaling.nn import MLPScaler
    from careless_b200 import synth

    state = {"u": None, "eps": None}

    def fake_truncated(shape, *a, **kw):
        """Stand-in for (stateless_)parameterized_truncated_normal(shape, [seed], means, stddevs, minvals, maxvals): the
        result has the SAMPLE dimension last ([batch, n])."""
        vals = list(a)
        if "minvals" in kw:
            lo, hi = kw["minvals"], kw["maxvals"]
        else:
            lo, hi = vals[-2], vals[-1]
        lo = np.broadcast_to(np.asarray(lo, dtype=np.float64).reshape(-1, 1), tuple(int(x) for x in shape))
        hi = np.broadcast_to(np.asarray(hi, dtype=np.float64).reshape(-1, 1), tuple(int(x) for x in shape))
        u = state["u"].T                                   # (R, S)
        pa, pb = ndtr(lo), ndtr(hi)
        z = pb - pa
        p = pa + u * z
        q = (1.0 - pb) + (1.0 - u) * z                     # upper tail without cancellation
        e = np.where(p < 0.5, ndtri(p), -ndtri(q))
        return tf.constant(e.astype(np.float32))

    def fake_normal(shape, mean=0.0, stddev=1.0, dtype=tf.float32, seed=None, name=None):
        e = state["eps"]
        shape = tuple(int(x) for x in shape)
        if int(np.prod(shape)) != e.size:
            raise RuntimeError(f"unexpected normal draw of shape {shape}: only the scale sample should draw")
        return tf.constant(e.reshape(shape).astype(np.float32))

    import tensorflow_probability.python.distributions.normal as tfp_normal
    import tensorflow_probability.python.distributions.truncated_normal as tfp_tn
    tfp_normal.samplers.normal = fake_normal
    patched = False
    for mod, name in ((tf.random, "stateless_parameterized_truncated_normal"), (tf.random, "parameterized_truncated_normal")):
        if hasattr(mod, name):
            setattr(mod, name, fake_truncated); patched = True
    for name in ("random_ops", "tf"):
        m = getattr(tfp_tn, name, None)
        for fn in ("parameterized_truncated_normal", "stateless_parameterized_truncated_normal"):
            if m is not None and hasattr(m, fn):
                setattr(m, fn, fake_truncated); patched = True
            if m is not None and hasattr(getattr(m, "random", None), fn):
                setattr(m.random, fn, fake_truncated); patched = True
    if not patched:
        raise SystemExit("could not find TFP's truncated-normal sampler to patch; adapt fake_truncated's hook to this TFP version")

    out = {}
    for name, (pk, mk) in CASES.items():
        p = synth.make_mono(pk["N"], pk["R"], d=pk["d"], n_images=pk["n_images"], seed=pk["seed"])
        N, R, S = pk["N"], pk["R"], mk["S"]
        rng = np.random.default_rng(7)
        u, eps = draws(rng, S, R, N, N_STEPS)
        prior = WilsonPrior(p["centric"], p["multiplicity"])
        loc, scale = prior.mean(), prior.stddev()
        low = (1e-32 * (~p["centric"])).astype("float32")                      # io/manager.py:434
        q = TruncatedNormal.from_loc_and_scale(loc, scale, low)
        lik = StudentTLikelihood(mk["dof"]) if mk["likelihood"] == "studentt" else NormalLikelihood()
        scaler = MLPScaler(mk["layers"], mk["width"], scale_bijector=tfb.Chain([tfb.Shift(1e-7), tfb.Exp()]))   # io/manager.py:457-463
        model = VariationalMergingModel(q, prior, lik, scaler, S)
        opt = tfk.optimizers.Adam(1e-3, 0.9, 0.99)                               # args/optimizer.py:5-27
        model.compile(opt, run_eagerly=True)
        col = lambda a, t: np.asarray(a).reshape(-1, 1).astype(t)
        data = (col(p["refl_id"], "int64"), col(p["image_id"], "int64"), col(p["file_id"], "int64"), p["metadata"].astype("float32"),
                col(p["intensities"], "float32"), col(p["uncertainties"], "float32"))
        data = tuple(tf.convert_to_tensor(x) for x in data)
        hist = {"loss": [], "NLL": [], "F KLDiv": [], "Grad Norm": []}
        for step in range(N_STEPS):
            state["u"], state["eps"] = u[step], eps[step]
            with tf.GradientTape() as tape:                                     # variational.py:185-224
                model(data, training=True)
                losses = list(model.losses)
                loss = tf.add_n(losses)
            grads = tape.gradient(loss, model.trainable_variables)
            gn = tf.linalg.global_norm(grads)
            if step == 0:
                for v, g in zip(model.trainable_variables, grads):
                    out[f"{name}/grad/{v.name}"] = np.asarray(g)
                    out[f"{name}/init/{v.name}"] = np.asarray(v)
                out[f"{name}/var_names"] = np.array([v.name for v in model.trainable_variables])
            # model.losses after call(): [kl (add_loss in add_kl_div), -ll (add_loss)] in that order (variational.py:123-181)
            kl, nll = float(losses[0]), float(losses[1])
            hist["loss"].append(float(loss)); hist["NLL"].append(nll); hist["F KLDiv"].append(kl); hist["Grad Norm"].append(float(gn))
            grads = [tf.where(tf.math.is_finite(g), g, 0.) for g in grads]
            opt.apply_gradients(zip(grads, model.trainable_variables))
        for k, v in hist.items():
            out[f"{name}/hist/{k}"] = np.asarray(v, dtype=np.float64)
        for v in model.trainable_variables:
            out[f"{name}/final/{v.name}"] = np.asarray(v)
        out[f"{name}/F"] = np.asarray(model.surrogate_posterior.mean())
        out[f"{name}/SigF"] = np.asarray(model.surrogate_posterior.stddev())
        out[f"{name}/u"] = u; out[f"{name}/eps"] = eps
        out[f"{name}/problem"] = np.array([pk["N"], pk["R"], pk["d"], pk["n_images"], pk["seed"]])
        out[f"{name}/model"] = np.array([mk["width"], mk["layers"], 1 if mk["likelihood"] == "studentt" else 0, mk["dof"] or 0.0, mk["S"]], dtype=np.float64)
        print(name, {k: v[0] for k, v in hist.items()})
    out["versions"] = np.array([tf.__version__, tfp.__version__, getattr(tfk, "__version__", "?")])
    np.savez_compressed(args.out, **out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
