"""MLP scale model (mirror of careless/models/scaling/nn.py:27-120).  Holds the keras-ordered weights;
forward/backward run fused in the CUDA observation kernel."""
import numpy as np

from ..base import BaseModel


class ScaleDistribution:
    """What `scaling_model(inputs)` returns (nn.py:22-25, image.py:58-63): the per-observation Normal over the scale, with the
    moments computed ON THE GPU by the observation kernel's forward pass (clb_get_scale_moments)."""

    def __init__(self, mean, stddev):
        self._mean, self._std = np.asarray(mean, dtype=np.float32), np.asarray(stddev, dtype=np.float32)

    def mean(self):
        return self._mean

    def stddev(self):
        return self._std

    def sample(self, n=1, seed=None):
        rng = np.random.default_rng(seed)
        return (self._mean + self._std * rng.standard_normal((int(n),) + self._mean.shape)).astype(np.float32)

    def log_prob(self, value):
        from scipy.stats import norm
        return norm.logpdf(np.asarray(value, dtype=np.float64), self._mean, self._std)


class Scaler(BaseModel):
    trainable = True

    def __call__(self, inputs):
        """`scaling_model(inputs) -> dist` of the reference (variational.py:67-69, 156-157).  The network runs on the GPU: a
        forward-only engine is built for `inputs` (any prior: the scale does not depend on the structure factors)."""
        from ..likelihoods.mono import NormalLikelihood
        from ..merging.surrogate_posteriors import TruncatedNormal
        from ..merging.variational import VariationalMergingModel
        from ..priors.wilson import WilsonPrior
        refl_id = np.asarray(self.get_refl_id(inputs)).reshape(-1)
        R = int(refl_id.max()) + 1
        prior = WilsonPrior(np.zeros(R, dtype=bool), np.ones(R, dtype=np.float32), 1.0)
        q = TruncatedNormal.from_loc_and_scale(prior.mean(), prior.stddev(), 1e-32)
        if self.is_laue(inputs):
            from ..likelihoods.laue import NormalLikelihood as LaueNormal
            lik = LaueNormal()
        else:
            lik = NormalLikelihood()
        model = VariationalMergingModel(q, prior, lik, self, 1)
        try:
            eng = model._build_engine(inputs)
            model._push(eng)
            mean, std = model._scale_moments(eng)          # per row, before any Laue convolution
        finally:
            model.close()
        return ScaleDistribution(mean, std)


class MetadataScaler(Scaler):
    def __init__(self, n_layers, width, leakiness=0.01, epsilon=1e-7, scale_bijector=None, scale_multiplier=None):
        if leakiness != 0.01:
            raise NotImplementedError("the CUDA kernel implements LeakyReLU(0.01), the only value the reference CLI uses")
        self.n_layers, self.width = int(n_layers), width
        self.epsilon = float(epsilon)
        # nn.py:14-18: softplus is the default of direct construction; the CLI passes exp (manager.py:457-463)
        self.scale_bijector = "softplus" if scale_bijector is None else str(scale_bijector).lower()
        if self.scale_bijector not in ("exp", "softplus"):
            raise ValueError(f"Unsupported scale bijector type, {scale_bijector}")
        self.scale_multiplier = scale_multiplier     # nn.py:84-87: additive tfb.Shift
        self.weights = None                          # built on first use (needs the metadata width)

    def build(self, n_meta):
        if self.weights is not None:
            if self.weights[0].shape[0] != n_meta:
                raise ValueError(f"the scale model was loaded for {self.weights[0].shape[0]} metadata columns, the data have {n_meta}")
            if self.width is None:
                self.width = int(self.weights[0].shape[1]) if self.n_layers > 0 else int(n_meta)
            return
        width = self.width if self.width is not None else n_meta
        self.width = int(width)
        ws, fan_in = [], n_meta
        for _ in range(self.n_layers):               # nn.py:55-68 identity kernels, zero bias
            ws += [np.eye(fan_in, width, dtype=np.float32), np.zeros(width, dtype=np.float32)]
            fan_in = width
        ws += [np.eye(fan_in, 2, dtype=np.float32), np.zeros(2, dtype=np.float32)]      # nn.py:72-79
        self.weights = ws

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, ws):
        ws = [np.asarray(w, dtype=np.float32) for w in ws]
        if self.weights is None:
            # not built yet (--scale-file on a fresh model, careless.py:48-51): adopt the stored shapes after checking
            # them against the constructor arguments; build(n_meta) later keeps these weights
            ok = len(ws) == 2 * (self.n_layers + 1) and all(w.ndim == (2 if i % 2 == 0 else 1) for i, w in enumerate(ws))
            if ok:
                width = ws[0].shape[1] if self.n_layers > 0 else ws[0].shape[0]
                ok = (self.width is None or int(self.width) == width or self.n_layers == 0) and ws[-2].shape[1] == 2
                fan_in = ws[0].shape[0]
                for k in range(self.n_layers):
                    ok = ok and ws[2 * k].shape == (fan_in, width) and ws[2 * k + 1].shape == (width,)
                    fan_in = width
                ok = ok and ws[-2].shape == (fan_in, 2) and ws[-1].shape == (2,)
            if not ok:
                raise ValueError("weight shapes do not match the model")
            if self.n_layers > 0:
                self.width = int(ws[0].shape[1])
            self.weights = [w.copy() for w in ws]
            return
        if len(ws) != len(self.weights) or any(a.shape != np.shape(b) for a, b in zip(self.weights, ws)):
            raise ValueError("weight shapes do not match the model")
        self.weights = [w.copy() for w in ws]

    def flat(self):
        return np.concatenate([w.reshape(-1) for w in self.weights])

    def from_flat(self, flat):
        off, out = 0, []
        for w in self.weights:
            out.append(np.asarray(flat[off:off + w.size], dtype=np.float32).reshape(w.shape)); off += w.size
        self.weights = out

    def save_weights(self, path):
        np.savez(path if str(path).endswith(".npz") else str(path) + ".npz", *self.weights)

    def load_weights(self, path):
        w = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        self.set_weights([w[k] for k in sorted(w.files, key=lambda s: int(s.split("_")[1]))])


class MLPScaler(MetadataScaler):
    pass
