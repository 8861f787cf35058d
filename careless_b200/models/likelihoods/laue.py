"""Laue likelihoods (mirror of careless/models/likelihoods/laue.py:36-100): predicted harmonics are
summed per spot (harmonic_id) before the mono log-density; done inside the CUDA observation kernel."""
import numpy as np

from .mono import Ev11Likelihood, Likelihood


class ConvolvedLikelihood:
    """laue.py:9-34: log-density of the per-spot sums of the predicted harmonics (host-side protocol object)."""

    def __init__(self, distribution, harmonic_id):
        self.distribution = distribution
        self.harmonic_id = np.asarray(harmonic_id).reshape(-1)

    def convolve(self, value):
        return LaueBase.convolve(value, self.harmonic_id)

    def mean(self):
        return self.distribution.mean()

    def stddev(self):
        return self.distribution.stddev()

    def log_prob(self, value):
        return self.distribution.log_prob(self.convolve(value))


class LaueBase(Likelihood):
    laue = True

    def call(self, inputs):                       # laue.py:42-47
        return ConvolvedLikelihood(self.dist(inputs), self.get_harmonic_id(inputs))

    @staticmethod
    def convolve(value, harmonic_id):
        """laue.py:17-25 on the host (post-processing helper): sum rows per spot, zeros elsewhere."""
        value = np.asarray(value)
        hid = np.asarray(harmonic_id).reshape(-1)
        out = np.zeros_like(value)
        np.add.at(out.T, hid, value.T)
        return out


class NormalLikelihood(LaueBase):
    kind = "normal"


class StudentTLikelihood(LaueBase):
    kind = "studentt"

    def __init__(self, dof):
        self.dof = float(dof)


class NormalEv11Likelihood(LaueBase, Ev11Likelihood):
    """laue.py:49-56"""
    kind = "normal"


class StudentTEv11Likelihood(LaueBase, Ev11Likelihood):
    """laue.py:58-65"""
    kind = "studentt"

    def __init__(self, dof):
        Ev11Likelihood.__init__(self)
        self.dof = float(dof)
