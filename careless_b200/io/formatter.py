"""Reflection files -> model inputs: ASU mapping, `refl_id`, `image_id`, metadata, Laue harmonics, `harmonic_id`.

Host-side mirror of careless/io/formatter.py (`DataFormatter` :60-186, `MonoFormatter` :188-400, `LaueFormatter`
:402-664), careless/utils/laue.py:5-81 (`expand_harmonics`) and careless/utils/positional_encoding.py:3-18, with the same
class names, constructor arguments and return values (a tuple of 2-D numpy arrays in `BaseModel.input_index` order and
a `ReciprocalASUCollection`).  The pandas group-bys become packed-key sorts; the integer outputs (`refl_id`,
`image_id`, `harmonic_id`, `file_id`) are exact by construction and are what SURVEY.md 8(f) rank 1 asks to be bit-exact.
"""
from __future__ import annotations

import warnings

import numpy as np

from ..models.base import BaseModel
from .asu import ReciprocalASU, ReciprocalASUCollection
from .mtz import DataSet, concat, read_mtz
from .symmetry import SpaceGroup


def ngroup(*keys):
    """pandas `groupby(keys).ngroup()`: dense rank of each row's key tuple in lexicographically sorted order."""
    n = len(keys[0])
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.lexsort(tuple(np.asarray(k) for k in reversed(keys)))
    new = np.zeros(n, dtype=bool)
    for k in keys:
        ks = np.asarray(k)[order]
        new[1:] |= ks[1:] != ks[:-1]
    out = np.empty(n, dtype=np.int64)
    out[order] = np.cumsum(new)
    return out


def positional_encoding(X, L):
    """NeRF encoding of columns normalised to [-1, 1] (careless/utils/positional_encoding.py:3-18)."""
    p = 2.0 * (X - X.min(-2)) / (X.max(-2) - X.min(-2)) - 1.0
    freq = np.pi * 2 ** np.arange(L, dtype=X.dtype)
    fp = (freq[..., None, :] * p[..., :, None]).reshape(p.shape[:-1] + (-1,))
    return np.concatenate((np.cos(fp), np.sin(fp)), axis=-1)


def standardize_metadata(metadata, metadata_keys=None):
    """Zero-mean / unit-variance columns; zero-variance columns are left alone (formatter.py:41-57)."""
    std = metadata.std(0)
    zeros = std == 0.0
    for k, v in enumerate(std):
        if v == 0.0:
            name = metadata_keys[k] if metadata_keys is not None else k
            warnings.warn(f'Metadata column "{name}" with zero standard deviation will not be standardized.')
    metadata[:, ~zeros] = (metadata[:, ~zeros] - metadata[:, ~zeros].mean(0)) / metadata[:, ~zeros].std(0)
    return metadata


def calculate_harmonic(H):
    return np.gcd.reduce(H, axis=-1)


def expand_harmonics(ds, dmin=None, wavelength_key="Wavelength"):
    """Every reflection on each observed central ray out to dmin (careless/utils/laue.py:9-81).
    Adds H_0, K_0, L_0 (innermost reflection of the ray); H,K,L and the wavelength become those of harmonic n."""
    if ds.merged:
        raise ValueError("Expected unmerged data, but ds.merged is True")
    ds = ds.copy()
    if "dHKL" not in ds:
        ds.compute_dHKL()
    if dmin is None:
        dmin = ds["dHKL"].min() - 1e-12
    Hobs = ds.get_hkls()
    nobs = calculate_harmonic(Hobs).astype(np.int64)
    H_0 = (Hobs / nobs[:, None]).astype(np.int32)
    d_0 = ds["dHKL"] * nobs
    wavelength_0 = ds[wavelength_key] * nobs
    n_max = np.floor_divide(d_0, dmin).astype(int)
    n = np.arange(n_max.max()) + 1
    idx, n = np.where(n[None, :] <= n_max[:, None])
    n = n + 1
    ds = ds.take(idx)
    ds["H_0"], ds["K_0"], ds["L_0"] = H_0[idx].T
    ds[wavelength_key] = wavelength_0[idx] / n
    ds.set_hkls(n[:, None] * H_0[idx])
    ds.compute_dHKL()
    return ds


def _check_key(key, dtype, flag, ds):
    if key in ds:
        return
    if key is None:
        msg = f"Unable to determine the {dtype} column key. Please use {flag} to specify the {dtype} key name."
    else:
        msg = f"User supplied {dtype} column key {key}, but {key} is not available in the input data. "
    raise ValueError(msg + " Available keys are: \n" + ",".join(ds.keys()))


class DataFormatter:
    spacegroups = None

    def pack_inputs(self, inputs_dict):
        inputs = ()
        for i in range(len(BaseModel.input_index)):
            k = BaseModel.get_name_by_index(i)
            if k not in inputs_dict:
                break
            inputs += (inputs_dict[k],)
        return inputs

    def _annotate(self, ds):
        image_key = self.image_key or ds.first_key_of_dtype("B")
        _check_key(image_key, "Batch", "--image-key", ds)
        intensity_key = self.intensity_key or ds.first_key_of_dtype("J")
        _check_key(intensity_key, "Intensity", "--intensity-key", ds)
        uncertainty_key = self.uncertainty_key
        if uncertainty_key is None:
            for prefix in ("Sig", "SIG"):
                if prefix + intensity_key in ds:
                    uncertainty_key = prefix + intensity_key
        if uncertainty_key is None:
            uncertainty_key = ds.first_key_of_dtype("Q")
        _check_key(uncertainty_key, "Stddev", "--uncertainty-key", ds)
        ds["intensity"] = ds[intensity_key]
        ds["uncertainty"] = ds[uncertainty_key]
        ds["image_id"] = ds[image_key]
        if self.isigi_cutoff is not None:
            ds = ds.take(~(ds["intensity"] / ds["uncertainty"] < self.isigi_cutoff))
        return ds

    def get_data_and_asu_collection(self, datasets):
        parts, cells, spacegroups = [], [], []
        for file_id, ds in enumerate(datasets):
            if self.spacegroups is not None:
                sg = self.spacegroups[file_id]
            elif ds.spacegroup is not None:
                sg = ds.spacegroup
            else:
                raise ValueError("Could not determine spacegroups. Please supply the --spacegroups flag")
            ds = self.prep_dataset(ds, sg)
            ds["file_id"] = np.full(len(ds), file_id, dtype=np.int64)
            ds["asu_id"] = np.full(len(ds), file_id if self.separate_outputs else 0, dtype=np.int64)
            parts.append(ds)
            cells.append(ds.cell)
            spacegroups.append(sg)
            if not ds.cell.is_compatible_with_spacegroup(sg):
                raise ValueError(f"Spacegroup {sg} found to be incompatible with unit cell constants {ds.cell} cannot proceed.")
        data = concat(parts)
        dmin = data["dHKL"].min()
        if self.separate_outputs:
            asus = [ReciprocalASU(c, s, dmin, self.anomalous) for c, s in zip(cells, spacegroups)]
        else:
            asus = [ReciprocalASU(data.cell, data.spacegroup, dmin, self.anomalous)]
        rac = ReciprocalASUCollection(asus)
        data["image_id"] = ngroup(data["file_id"], data["image_id"])
        return data, rac

    def __call__(self, datasets):
        data, rac = self.get_data_and_asu_collection(datasets)
        return self.finalize(data, rac)

    def format_files(self, files):
        def load(filename):
            if filename.endswith(".mtz"):
                return read_mtz(filename)
            if filename.endswith(".stream"):
                from .crystfel import read_crystfel
                return read_crystfel(filename)
            raise ValueError(f"{filename}: expected a .mtz or .stream file")
        return self(load(f) for f in files)

    def _metadata(self, data):
        missing = [k for k in self.metadata_keys if k not in data]
        if missing:
            raise ValueError("".join(f'Metadata key "{k}" not found in input data. \n' for k in missing)
                             + "Available keys are: \n" + ",".join(data.keys()))
        metadata = np.stack([np.asarray(data[k]).astype(np.float32) for k in self.metadata_keys], axis=1)
        if self.standardize:
            metadata = standardize_metadata(metadata, self.metadata_keys)
        if self.positional_encoding_keys is not None:
            to_encode = np.stack([np.asarray(data[k]).astype(np.float32) for k in self.positional_encoding_keys], axis=1)
            metadata = np.concatenate((metadata, positional_encoding(to_encode, self.ecoding_bit_depth)), axis=1)
        return metadata


class MonoFormatter(DataFormatter):
    def __init__(self, intensity_key, uncertainty_key, image_key, metadata_keys, separate_outputs, anomalous, dmin=0.0,
                 isigi_cutoff=None, positional_encoding_keys=None, encoding_bit_depth=5, spacegroups=None, standardize=True):
        self.intensity_key, self.uncertainty_key, self.image_key = intensity_key, uncertainty_key, image_key
        self.metadata_keys = list(metadata_keys)
        self.separate_outputs, self.anomalous = separate_outputs, anomalous
        self.dmin = 0.0 if dmin is None else dmin
        self.isigi_cutoff = isigi_cutoff
        self.positional_encoding_keys = positional_encoding_keys
        self.ecoding_bit_depth = encoding_bit_depth
        self.spacegroups = _spacegroups(spacegroups)
        self.standardize = standardize

    def prep_dataset(self, ds, spacegroup=None, inplace=True):
        """dmin cut, absences removed, observed indices kept as {H,K,L}obs, H,K,L mapped to the ASU (formatter.py:274-352)."""
        ds = ds if inplace else ds.copy()
        if spacegroup is not None:
            ds.spacegroup = spacegroup
        ds.compute_dHKL()
        ds = ds.take(~(ds["dHKL"] < self.dmin))
        ds = ds.remove_absences()
        hkl = ds.get_hkls()
        ds["Hobs"], ds["Kobs"], ds["Lobs"] = hkl.T
        ds.hkl_to_asu(anomalous=self.anomalous)
        return self._annotate(ds)

    def finalize(self, data, rac):
        data["dHKL"] = data["dHKL"] ** -2.0
        metadata = self._metadata(data)
        refl_id = rac.to_refl_id(data["asu_id"], data.get_hkls())
        inputs = {
            "refl_id": refl_id[:, None],
            "image_id": np.asarray(data["image_id"], dtype=np.int64)[:, None],
            "metadata": metadata,
            "intensities": np.asarray(data["intensity"], dtype=np.float32)[:, None],
            "uncertainties": np.asarray(data["uncertainty"], dtype=np.float32)[:, None],
            "file_id": np.asarray(data["file_id"], dtype=np.int64)[:, None],
        }
        return self.pack_inputs(inputs), rac


class LaueFormatter(DataFormatter):
    def __init__(self, wavelength_key, intensity_key, uncertainty_key, image_key, metadata_keys, separate_outputs, anomalous,
                 lam_min=None, lam_max=None, dmin=0.0, isigi_cutoff=None, positional_encoding_keys=None, encoding_bit_depth=5,
                 spacegroups=None, standardize=True):
        self.wavelength_key, self.lam_min, self.lam_max = wavelength_key, lam_min, lam_max
        self.intensity_key, self.uncertainty_key, self.image_key = intensity_key, uncertainty_key, image_key
        self.metadata_keys = list(metadata_keys)
        self.separate_outputs, self.anomalous = separate_outputs, anomalous
        self.dmin, self.isigi_cutoff = dmin, isigi_cutoff
        self.positional_encoding_keys = positional_encoding_keys
        self.ecoding_bit_depth = encoding_bit_depth
        self.spacegroups = _spacegroups(spacegroups)
        self.standardize = standardize

    def prep_dataset(self, ds, spacegroup=None):
        """Harmonic expansion to dmin, wavelength window, absences, ASU mapping (formatter.py:505-597)."""
        if spacegroup is not None:
            ds.spacegroup = spacegroup
        ds.compute_dHKL()
        dmin = self.dmin if self.dmin is not None else ds["dHKL"].min()
        wl = ds[self.wavelength_key]
        lam_min = self.lam_min if self.lam_min is not None else wl.min()
        lam_max = self.lam_max if self.lam_max is not None else wl.max()
        ds = expand_harmonics(ds, dmin, self.wavelength_key)
        hkl = ds.get_hkls()
        ds["Hobs"], ds["Kobs"], ds["Lobs"] = hkl.T
        wl = ds[self.wavelength_key]
        ds = ds.take(~((wl < lam_min) | (wl > lam_max)))
        ds = ds.remove_absences()
        ds.hkl_to_asu(anomalous=self.anomalous)
        return self._annotate(ds)

    def finalize(self, data, rac):
        harmonic_id = ngroup(data["image_id"], data["H_0"], data["K_0"], data["L_0"])
        data["harmonic_id"] = harmonic_id
        data["dHKL"] = data["dHKL"] ** -2.0
        metadata = self._metadata(data)
        refl_id = rac.to_refl_id(data["asu_id"], data.get_hkls())
        _, first = np.unique(harmonic_id, return_index=True)
        pad = len(refl_id) - len(first)
        iobs = np.pad(np.asarray(data["intensity"], dtype=np.float32)[first][:, None], [[0, pad], [0, 0]], constant_values=1.0)
        sigma = np.pad(np.asarray(data["uncertainty"], dtype=np.float32)[first][:, None], [[0, pad], [0, 0]], constant_values=1.0)
        inputs = {
            "refl_id": refl_id[:, None],
            "image_id": np.asarray(data["image_id"], dtype=np.int64)[:, None],
            "metadata": metadata,
            "intensities": iobs,
            "uncertainties": sigma,
            "file_id": np.asarray(data["file_id"], dtype=np.int64)[:, None],
            "wavelength": np.asarray(data[self.wavelength_key], dtype=np.float32)[:, None],
            "harmonic_id": harmonic_id[:, None],
        }
        return self.pack_inputs(inputs), rac

    def format_files(self, files):
        for file in files:
            if file.endswith(".stream"):
                raise ValueError("careless poly does not support .stream files. Use careless mono instead.")
        return super().format_files(files)


def _spacegroups(spacegroups):
    if spacegroups is None:
        return None
    return [s if isinstance(s, SpaceGroup) else SpaceGroup.from_name(s) for s in spacegroups]
