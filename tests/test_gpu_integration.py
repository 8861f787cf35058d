"""careless_b200.integration.B200TrainMixin driven through the REFERENCE's object protocol with keras-shaped stand-ins
(variables with .numpy() / .assign(), scaler.trainable_variables in keras order, optimizer attributes, BaseModel accessors):
what a careless maintainer's binding would exercise, on a machine without TensorFlow.  The history must equal the oracle's."""
import numpy as np
import pytest

from careless_b200 import synth
from careless_b200.integration import B200TrainMixin
from oracle import model as om

import _util as U

pytestmark = pytest.mark.gpu


class Var:                                   # tf.Variable look-alike
    def __init__(self, value):
        self.value = np.array(value, dtype=np.float32)

    def numpy(self):
        return self.value

    def assign(self, v):
        v = np.asarray(v, dtype=np.float32)
        assert v.shape == self.value.shape
        self.value = v.copy()


class Holder:
    def __init__(self, variables, **kw):
        self.trainable_variables, self.trainable = variables, True
        self.__dict__.update(kw)


class WilsonPrior:                           # attribute names of priors/wilson.py:43-47
    def __init__(self, centric, epsilon, sigma=1.0):
        self.centric, self.epsilon, self.sigma = np.array(centric, bool), np.array(epsilon, np.float32), np.array(sigma, np.float32)


class FakeKerasModel(B200TrainMixin):
    """Carries exactly what VariationalMergingModel carries (variational.py:15-45) + the BaseModel accessors."""

    def __init__(self, q, prior, likelihood, scaler, mc, optimizer):
        self.surrogate_posterior, self.prior, self.likelihood, self.scaling_model = q, prior, likelihood, scaler
        self.mc_sample_size, self.kl_weight, self.optimizer = mc, None, optimizer

    @staticmethod
    def is_laue(inputs):
        return len(inputs) > 6

    get_refl_id = staticmethod(lambda inputs: inputs[0])
    get_image_id = staticmethod(lambda inputs: inputs[1])
    get_metadata = staticmethod(lambda inputs: inputs[3])
    get_intensities = staticmethod(lambda inputs: inputs[4])
    get_uncertainties = staticmethod(lambda inputs: inputs[5])
    get_harmonic_id = staticmethod(lambda inputs: inputs[7])


def test_mixin_trains_a_keras_shaped_model_like_the_oracle():
    p = synth.make_mono(4000, 500, d=3, n_images=9, seed=21)
    R, S, L, W, steps = 500, 1, 4, 10, 3
    ocfg = om.ModelConfig(n_refl=R, n_meta=3, mlp_width=W, mlp_layers=L, likelihood="studentt", dof=8.0, mc_samples=S)
    oprior = om.PriorData(p["centric"], p["multiplicity"])
    params = U.perturbed_params(ocfg, oprior, np.random.default_rng(3), amount=0.05)
    names = U.mlp_names(ocfg)
    q = Holder([Var(params["sf_loc_raw"].numpy()), Var(params["sf_scale_raw"].numpy())])
    scaler = Holder([Var(params[n].numpy()) for n in names])
    opt = Holder([], learning_rate=1e-3, beta_1=0.9, beta_2=0.99, epsilon=1e-7, clipnorm=None, clipvalue=None, global_clipnorm=None)
    lik = Holder([], dof=8.0)
    model = FakeKerasModel(q, WilsonPrior(p["centric"], p["multiplicity"]), lik, scaler, S, opt)
    col = lambda a, t: np.asarray(a).reshape(-1, 1).astype(t)
    data = (col(p["refl_id"], np.int64), col(p["image_id"], np.int64), col(np.zeros(4000), np.int64), p["metadata"].astype(np.float32),
            col(p["intensities"], np.float32), col(p["uncertainties"], np.float32))
    hist = model.train_model(data, steps, progress=False)
    assert set(hist) == {"loss", "NLL", "F KLDiv", "Grad Norm"} and len(hist["loss"]) == steps
    # the same Philox draws through the oracle
    from oracle import philox
    state, oopt = om.adam_init(params), om.AdamConfig()
    for step in range(steps):
        u = philox.refl_uniforms(1234, step, S, np.arange(R))
        e = philox.obs_normals(1234, step, S, np.arange(4000))
        metrics, g, _ = om.loss_and_grads(params, p, oprior, ocfg, u, e)
        for k in ("loss", "NLL", "F KLDiv", "Grad Norm"):
            assert abs(hist[k][step] - metrics[k]) <= 1e-4 * abs(metrics[k]), (step, k, hist[k][step], metrics[k])
        params = om.adam_apply(params, g, state, oopt)
    assert U.rel_err(q.trainable_variables[0].numpy(), params["sf_loc_raw"].numpy()) <= 1e-4
    for var, n in zip(scaler.trainable_variables, names):
        assert U.rel_err(var.numpy(), params[n].numpy()) <= 1e-4, n


def test_mixin_rejects_what_it_does_not_bind():
    class DoubleWilsonPrior(WilsonPrior):
        pass
    p = synth.make_mono(200, 30, d=2, n_images=3, seed=1)
    q = Holder([Var(np.zeros(30)), Var(np.zeros(30))])
    scaler = Holder([Var(np.eye(2, 2)), Var(np.zeros(2))])
    model = FakeKerasModel(q, DoubleWilsonPrior(p["centric"], p["multiplicity"]), Holder([]), scaler, 1, Holder([]))
    col = lambda a, t: np.asarray(a).reshape(-1, 1).astype(t)
    data = (col(p["refl_id"], np.int64), col(p["image_id"], np.int64), col(np.zeros(200), np.int64), p["metadata"].astype(np.float32),
            col(p["intensities"], np.float32), col(p["uncertainties"], np.float32))
    with pytest.raises(NotImplementedError):
        model.train_model(data, 1)
