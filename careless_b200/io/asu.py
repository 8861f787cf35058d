"""Reciprocal ASUs and their collection: Miller index <-> `refl_id`.

Mirrors careless/io/asu.py:5-178 (`ReciprocalASU`, `ReciprocalASUCollection`, same attribute and method names) on
numpy arrays; the pandas MultiIndex joins of asu.py:150-172 become one packed-integer key + binary search, which is
what stays cheap at 2 M reflections / 200 M observations.
"""
from __future__ import annotations

import numpy as np

from .symmetry import generate_reciprocal_asu

_OFF = 1 << 15      # Miller indices are packed as 16-bit fields, asu_id above them
_SPAN = 1 << 16


def _pack(asu_id, hkl):
    hkl = np.asarray(hkl, dtype=np.int64).reshape(-1, 3)
    if hkl.size and (np.abs(hkl).max() >= _OFF):
        raise ValueError("Miller index out of the +-32767 range")
    a = np.asarray(asu_id, dtype=np.int64).reshape(-1)
    return ((a * _SPAN + hkl[:, 0] + _OFF) * _SPAN + hkl[:, 1] + _OFF) * _SPAN + hkl[:, 2] + _OFF


class ReciprocalASU:
    def __init__(self, cell, spacegroup, dmin, anomalous):
        self.cell, self.spacegroup, self.dmin, self.anomalous = cell, spacegroup, dmin, anomalous
        self.Hall = generate_reciprocal_asu(cell, spacegroup, dmin, anomalous)
        self._centric = spacegroup.is_centric(self.Hall)
        self._epsilon = spacegroup.epsilon(self.Hall).astype(np.float32)
        self._dHKL = cell.calculate_d_array(self.Hall).astype(np.float32)
        keys = _pack(np.zeros(len(self.Hall)), self.Hall)
        self._order = np.argsort(keys, kind="stable")
        self._keys = keys[self._order]

    def __len__(self):
        return len(self.Hall)

    @property
    def centric(self):
        return self._centric

    @property
    def multiplicity(self):
        return self._epsilon

    @property
    def dHKL(self):
        return self._dHKL

    def to_refl_id(self, H):
        k = _pack(np.zeros(len(H)), H)
        pos = np.searchsorted(self._keys, k)
        pos = np.minimum(pos, len(self._keys) - 1)
        if not np.all(self._keys[pos] == k):
            raise KeyError("Miller indices outside this reciprocal ASU")
        return self._order[pos].astype(np.int64)

    def to_miller_index(self, refl_id):
        return self.Hall[np.asarray(refl_id).reshape(-1)]


class ReciprocalASUCollection:
    def __init__(self, reciprocal_asus):
        self.reciprocal_asus = list(reciprocal_asus)
        self._hkls = np.concatenate([a.Hall for a in self.reciprocal_asus]) if self.reciprocal_asus else np.zeros((0, 3), np.int32)
        self._asu_ids = np.concatenate([np.full(len(a), i, dtype=np.int64) for i, a in enumerate(self.reciprocal_asus)])
        self._centric = np.concatenate([a.centric for a in self.reciprocal_asus])
        self._epsilon = np.concatenate([a.multiplicity for a in self.reciprocal_asus])
        self._dHKL = np.concatenate([a.dHKL for a in self.reciprocal_asus])
        keys = _pack(self._asu_ids, self._hkls)
        self._order = np.argsort(keys, kind="stable")
        self._keys = keys[self._order]

    @property
    def centric(self):
        return self._centric

    @property
    def multiplicity(self):
        return self._epsilon

    @property
    def dHKL(self):
        return self._dHKL

    @property
    def hkls(self):
        return self._hkls

    @property
    def asu_ids(self):
        return self._asu_ids

    def to_asu_id_and_miller_index(self, refl_id):
        r = np.asarray(refl_id).reshape(-1)
        return self._asu_ids[r][:, None], self._hkls[r]

    def to_refl_id(self, asu_id, H, allow_missing=False):
        k = _pack(np.asarray(asu_id).reshape(-1), H)
        pos = np.minimum(np.searchsorted(self._keys, k), max(len(self._keys) - 1, 0))
        hit = self._keys[pos] == k if len(self._keys) else np.zeros(len(k), dtype=bool)
        if not allow_missing and not np.all(hit):
            raise KeyError(f"{int((~hit).sum())} (asu_id, H, K, L) entries are not in the collection")
        return np.where(hit, self._order[pos], -1).astype(np.int64)

    def __getitem__(self, i):
        return self.reciprocal_asus[i]

    def __len__(self):
        return len(self.reciprocal_asus)

    def __iter__(self):
        return iter(self.reciprocal_asus)
