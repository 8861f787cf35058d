"""Learnable truncated-normal surrogate (mirror of careless/models/merging/surrogate_posteriors.py:45-131).

The trainable state is the pair of raw vectors (log loc, log(scale - eps)); sampling, log-density and the
gradients run in the CUDA kernels (csrc/clb_math.cuh: tn_forward).  Moments are host post-processing."""
import numpy as np


class TruncatedNormal:
    def __init__(self, loc, scale, low, high, scale_shift=1e-7):
        loc = np.asarray(loc, dtype=np.float32)
        scale = np.asarray(scale, dtype=np.float32)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float32), loc.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float32), loc.shape).copy()
        self.scale_shift = float(scale_shift)
        self.loc_raw = np.log(loc)                       # tfb.Exp
        self.scale_raw = np.log(scale - np.float32(scale_shift))   # tfb.Chain([Shift(eps), Exp])
        self.trainable = True

    @classmethod
    def from_loc_and_scale(cls, loc, scale, low=0., high=1e10, scale_shift=1e-7):
        return cls(loc, scale, low, high, scale_shift)

    @property
    def loc(self):
        return np.exp(self.loc_raw)

    @property
    def scale(self):
        return np.exp(self.scale_raw) + np.float32(self.scale_shift)

    @property
    def parameters(self):
        return {"loc": self.loc, "scale": self.scale, "low": self.low, "high": self.high}

    def _ab(self, high=None):
        high = self.high if high is None else high
        loc, scale = self.loc.astype(np.float64), self.scale.astype(np.float64)
        return (self.low - loc) / scale, (high - loc) / scale, loc, scale

    def mean(self):
        from scipy.stats import truncnorm
        a, b, loc, scale = self._ab()
        return truncnorm.mean(a, b, loc, scale).astype(np.float32)

    def stddev(self):
        from scipy.stats import truncnorm
        a, b, loc, scale = self._ab()
        return truncnorm.std(a, b, loc, scale).astype(np.float32)

    def moment_4(self, high=np.inf, method='scipy'):
        if method != 'scipy':
            raise ValueError(f"Unknown method {method} for computing moment_4")
        from scipy.stats import truncnorm
        a, b, loc, scale = self._ab(high)
        return truncnorm.moment(4, a, b, loc, scale)

    # ---- the rest of the reference protocol (surrogate_posteriors.py:11-37, :50-53), host side, float64 scipy ----
    def sample(self, n=1, seed=None):
        """(n, R) draws: inverse-CDF transform of uniforms, then max(low, .) as surrogate_posteriors.py:50-53."""
        from scipy.stats import truncnorm
        a, b, loc, scale = self._ab()
        rng = np.random.default_rng(seed)
        u = rng.random((int(n),) + loc.shape)
        s = truncnorm.ppf(u, a, b, loc, scale)
        return np.maximum(self.low, s).astype(np.float32)

    def log_prob(self, z):
        from scipy.stats import truncnorm
        a, b, loc, scale = self._ab()
        return truncnorm.logpdf(np.asarray(z, dtype=np.float64), a, b, loc, scale)

    def parameter_properties(self):
        """Which unconstraining transform each parameter trains under (the role of tfd's ParameterProperties here)."""
        return {"loc": {"trainable": True, "bijector": "Exp", "raw": "loc_raw"},
                "scale": {"trainable": True, "bijector": f"Chain([Shift({self.scale_shift:g}), Exp])", "raw": "scale_raw"},
                "low": {"trainable": False}, "high": {"trainable": False}}

    def save_weights(self, path):
        np.savez(path if str(path).endswith(".npz") else str(path) + ".npz", loc_raw=self.loc_raw, scale_raw=self.scale_raw)

    def load_weights(self, path):
        w = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        self.loc_raw, self.scale_raw = w["loc_raw"].astype(np.float32), w["scale_raw"].astype(np.float32)
