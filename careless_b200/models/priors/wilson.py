"""Wilson / DoubleWilson priors (mirror of careless/models/priors/wilson.py:29-175).

The objects hold the per-reflection tables; log_prob and its gradient run on the GPU
(csrc/clb_math.cuh: wilson_logp, dw_child_logp)."""
import numpy as np

from ..base import BaseModel


class Prior(BaseModel):
    pass


class WilsonPrior(Prior):
    kind = "wilson"

    def __init__(self, centric, epsilon, sigma=1.):
        self.epsilon = np.array(epsilon, dtype=np.float32)
        self.centric = np.array(centric, dtype=bool)
        self.sigma = np.broadcast_to(np.array(sigma, dtype=np.float32), self.epsilon.shape).copy()

    def _scale(self):
        return np.sqrt(self.epsilon.astype(np.float64) * self.sigma)

    def mean(self):
        s = self._scale()
        return np.where(self.centric, s * np.sqrt(2. / np.pi), s * np.sqrt(np.pi) / 2.).astype(np.float32)

    def stddev(self):
        s = self._scale()
        return np.where(self.centric, s * np.sqrt(1. - 2. / np.pi), s * np.sqrt(1. - np.pi / 4.)).astype(np.float32)


class DoubleWilsonPrior(Prior):
    """wilson.py:82-175.  The reference builds ``reflids`` from a ReciprocalASUCollection
    (:112-137); here the already-mapped tables are passed in (the ASU algebra is host prep)."""
    kind = "double_wilson"

    def __init__(self, centric, epsilon, asu_ids, reflids, root, r_values, sigma=1., optimize_r=False):
        self.wilson_prior = WilsonPrior(centric, epsilon, sigma)
        self.centric, self.epsilon, self.sigma = self.wilson_prior.centric, self.wilson_prior.epsilon, self.wilson_prior.sigma
        self.asu_ids = np.asarray(asu_ids, dtype=np.int32)
        self.reflids = np.asarray(reflids, dtype=np.int64)
        self.root = np.asarray(root, dtype=bool)
        self.r = np.asarray(r_values, dtype=np.float32)
        for r in self.r:
            if (r >= 1.) or (r <= -1.):     # manager.py:415-419
                raise ValueError(f"Supplied --double-wilson-r value {r} outside of allowed range (-1, 1)")
        self.optimize_r = bool(optimize_r)

    @classmethod
    def from_asu_collection(cls, asu_collection, parents, r_values, reindexing_ops=None, sigma=1., optimize_r=False):
        """The reference constructor (wilson.py:83-137): parents[i] = j makes ASU j the parent of ASU i; every child
        reflection (optionally reindexed) is mapped into the parent's ASU and looked up in the collection (-1 = parent
        absent).  Root ASUs keep their LOCAL ids, as the reference does (never used: roots have no parent term)."""
        reflids, root = [], []
        for child, parent in enumerate(parents):
            child_asu = asu_collection.reciprocal_asus[child]
            n = len(child_asu.Hall)
            if parent is None:
                reflids.append(np.arange(n, dtype=np.int64))
                root.append(np.ones(n, dtype=bool))
            else:
                root.append(np.zeros(n, dtype=bool))
                parent_asu = asu_collection.reciprocal_asus[parent]
                h = child_asu.Hall
                if reindexing_ops is not None:
                    h = reindexing_ops[child].apply_to_hkl(h)
                h, _ = parent_asu.spacegroup.hkl_to_asu(h)
                reflids.append(asu_collection.to_refl_id(np.full((len(h), 1), parent), h, allow_missing=True))
        out = cls(asu_collection.centric, asu_collection.multiplicity, asu_collection.asu_ids, np.concatenate(reflids),
                  np.concatenate(root), r_values, sigma=sigma, optimize_r=optimize_r)
        out.parents = parents
        return out

    @property
    def dw_parent(self):
        """Device encoding: -2 root entry, -1 parent absent, >=0 surrogate index of the parent."""
        out = np.where(self.reflids >= 0, self.reflids, -1).astype(np.int32)
        out[self.root] = -2
        return out

    def mean(self):
        return self.wilson_prior.mean()

    def stddev(self):
        return self.wilson_prior.stddev()
