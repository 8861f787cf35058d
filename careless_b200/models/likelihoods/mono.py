"""Error models for mono data (mirror of careless/models/likelihoods/mono.py:16-37).

These objects only carry the choice and its hyper-parameters; log_prob and its gradient are
evaluated inside the CUDA observation kernel (csrc/clb_math.cuh: lik_eval)."""
from ..base import BaseModel


class Likelihood(BaseModel):
    kind = None
    laue = False
    dof = None


class NormalLikelihood(Likelihood):
    kind = "normal"


class StudentTLikelihood(Likelihood):
    kind = "studentt"

    def __init__(self, dof):
        self.dof = float(dof)
