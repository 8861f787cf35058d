"""Error models for mono data (mirror of careless/models/likelihoods/mono.py:16-37).

During training log_prob and its gradient are evaluated inside the CUDA observation kernel (csrc/clb_math.cuh:
lik_eval / ev11_eval); these objects carry the choice and its hyper-parameters.  They also keep the reference's object
protocol -- `likelihood(inputs)` returns an object with `.log_prob(ipred)`, `.mean()`, `.stddev()` (variational.py:169-171,
likelihoods/mono.py:16-37) -- as a float64 scipy evaluation on the host, for inspection and post-processing only (the
training step never calls it)."""
import numpy as np

from ..base import BaseModel


class LocationScaleDistribution:
    """What `likelihood(inputs)` returns: Normal(loc, scale) or StudentT(dof, loc, scale) over the observed intensities."""

    def __init__(self, kind, loc, scale, dof=None):
        self.kind, self.dof = kind, dof
        self.loc = np.asarray(loc, dtype=np.float64).reshape(-1)
        self.scale = np.asarray(scale, dtype=np.float64).reshape(-1)

    def _frozen(self, scale=None):
        from scipy import stats
        scale = self.scale if scale is None else scale
        return stats.norm(self.loc, scale) if self.kind == "normal" else stats.t(self.dof, self.loc, scale)

    def log_prob(self, value):
        return self._frozen().logpdf(np.asarray(value, dtype=np.float64))

    def mean(self):
        return self.loc.copy()

    def stddev(self):
        return self._frozen().std()


class Ev11Distribution(LocationScaleDistribution):
    """mono.py:46-59: the scale depends on the prediction, sigma' = Sdfac sqrt(sigma^2 + SdB p + Sdadd p^2), p = softplus(x)."""

    def __init__(self, kind, loc, scale, sdfac, sdadd, sdb, dof=None):
        super().__init__(kind, loc, scale, dof)
        self.sdfac, self.sdadd, self.sdb = sdfac, sdadd, sdb

    def log_prob(self, value):
        x = np.asarray(value, dtype=np.float64)
        p = np.logaddexp(0.0, x)
        scale = self.sdfac * np.sqrt(self.scale ** 2 + self.sdb * p + self.sdadd * p * p)
        return self._frozen(scale).logpdf(x)


class Likelihood(BaseModel):
    kind = None
    laue = False
    dof = None
    refine_uncertainties = False
    trainable = True

    def get_loc_and_scale(self, inputs):            # mono.py:11-14
        return (np.asarray(self.get_intensities(inputs), dtype=np.float64).reshape(-1),
                np.asarray(self.get_uncertainties(inputs), dtype=np.float64).reshape(-1))

    def dist(self, inputs):
        loc, scale = self.get_loc_and_scale(inputs)
        if self.refine_uncertainties:
            return Ev11Distribution(self.kind, loc, scale, self.Sdfac, self.Sdadd, self.SdB, self.dof)
        return LocationScaleDistribution(self.kind, loc, scale, self.dof)

    def call(self, inputs):
        return self.dist(inputs)

    def __call__(self, inputs):
        return self.call(inputs)


class NormalLikelihood(Likelihood):
    kind = "normal"


class StudentTLikelihood(Likelihood):
    kind = "studentt"

    def __init__(self, dof):
        self.dof = float(dof)


class Ev11Likelihood(Likelihood):
    """mono.py:39-59 (--refine-uncertainties): sigma' = Sdfac sqrt(sigma^2 + SdB p + Sdadd p^2), p = softplus(Ipred).
    The three softplus-transformed scalars start at 1 and are trained by the engine (group "likelihood")."""
    refine_uncertainties = True

    def __init__(self):
        import numpy as np
        self.raw = np.full(3, np.log(np.e - 1.0), dtype=np.float32)     # softplus^-1(1)

    def _softplus(self, i):
        import numpy as np
        return float(np.log1p(np.exp(self.raw[i])))

    @property
    def Sdfac(self): return self._softplus(0)
    @property
    def Sdadd(self): return self._softplus(1)
    @property
    def SdB(self): return self._softplus(2)


class NormalEv11Likelihood(Ev11Likelihood):
    kind = "normal"


class StudentTEv11Likelihood(Ev11Likelihood):
    kind = "studentt"

    def __init__(self, dof):
        super().__init__()
        self.dof = float(dof)
