"""Error models for mono data (mirror of careless/models/likelihoods/mono.py:16-37).

These objects only carry the choice and its hyper-parameters; log_prob and its gradient are
evaluated inside the CUDA observation kernel (csrc/clb_math.cuh: lik_eval)."""
from ..base import BaseModel


class Likelihood(BaseModel):
    kind = None
    laue = False
    dof = None
    refine_uncertainties = False
    trainable = True


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
