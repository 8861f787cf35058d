"""Synthetic careless inputs of the BASELINE.json shapes.

Everything is generated with ``numpy.random.Generator(PCG64(seed))`` (reference default
``--seed 1234``, ``careless/args/tf_options.py:50-54``) in the reference's input-tuple layout
(``careless/models/base.py:22-31``): a dict with ``refl_id, image_id, file_id, metadata,
intensities, uncertainties`` (+ ``wavelength, harmonic_id`` for Laue) plus the per-reflection
prior tables (``centric, multiplicity``).
"""
from __future__ import annotations

import numpy as np


def _wilson_sample(rng, centric, eps):
    """F_true ~ Wilson(eps, Sigma=1): half-normal (centric) / Rayleigh-like Weibull(2) (acentric)."""
    s = np.sqrt(eps)
    fc = np.abs(rng.standard_normal(eps.shape)) * s
    fa = s * np.sqrt(-np.log(1.0 - rng.random(eps.shape)))
    return np.where(centric, fc, fa)


def reflection_tables(rng, R):
    centric = rng.random(R) < 0.10
    mult = rng.choice(np.array([1.0, 2.0, 3.0, 4.0, 6.0], dtype=np.float32), size=R, p=[0.9, 0.05, 0.02, 0.02, 0.01])
    f_true = _wilson_sample(rng, centric, mult.astype(np.float64))
    f_true = np.maximum(f_true, 1e-3)
    return centric, mult.astype(np.float32), f_true


def _scale_and_noise(rng, meta, f2):
    k_true = np.exp(0.3 * meta[:, 0] - 0.2 * meta[:, 1] ** 2) if meta.shape[1] >= 2 else np.exp(0.3 * meta[:, 0])
    mean = k_true * f2
    return mean


def make_mono(N, R, d=5, n_images=5000, seed=1234):
    """configs[1]: synthetic mono (10M obs, 500k reflections at full size)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    centric, mult, f_true = reflection_tables(rng, R)
    refl_id = np.sort(rng.integers(0, R, size=N)).astype(np.int64)
    image_id = rng.integers(0, n_images, size=N).astype(np.int64)
    meta = rng.standard_normal((N, d)).astype(np.float32)
    mean = _scale_and_noise(rng, meta.astype(np.float64), f_true[refl_id] ** 2)
    sig = np.sqrt((0.03 * mean) ** 2 + 0.05 ** 2)
    iobs = mean + sig * rng.standard_normal(N)
    return {
        "refl_id": refl_id, "image_id": image_id, "file_id": np.zeros(N, dtype=np.int64), "metadata": meta,
        "intensities": iobs.astype(np.float32), "uncertainties": sig.astype(np.float32),
        "centric": centric, "multiplicity": mult, "n_images": n_images, "f_true": f_true,
    }


def make_laue(n_rows, R, d=5, n_images=2000, seed=1234):
    """configs[2]: expanded harmonic rows; spot multiplicity in {1..5} w.p. {.84,.10,.04,.015,.005}."""
    rng = np.random.Generator(np.random.PCG64(seed))
    centric, mult, f_true = reflection_tables(rng, R)
    p = np.array([0.84, 0.10, 0.04, 0.015, 0.005])
    mean_len = float((p * np.arange(1, 6)).sum())
    n_spots_guess = int(n_rows / mean_len) + 8
    lens = rng.choice(np.arange(1, 6), size=n_spots_guess, p=p)
    csum = np.cumsum(lens)
    n_spots = int(np.searchsorted(csum, n_rows, side="right"))
    lens = lens[:n_spots]
    rest = n_rows - int(lens.sum())
    if rest > 0:
        lens = np.concatenate([lens, np.ones(rest, dtype=lens.dtype)])
        n_spots += rest
    spot_of_row = np.repeat(np.arange(n_spots), lens)
    spot_image = np.sort(rng.integers(0, n_images, size=n_spots))       # ngroup order is image-major
    refl_id = rng.integers(0, R, size=n_rows).astype(np.int64)
    meta = rng.standard_normal((n_rows, d)).astype(np.float32)
    wavelength = (1.0 + 0.1 * rng.random(n_rows)).astype(np.float32)
    mean_row = _scale_and_noise(rng, meta.astype(np.float64), f_true[refl_id] ** 2)
    mean_spot = np.bincount(spot_of_row, weights=mean_row, minlength=n_spots)
    sig_spot = np.sqrt((0.03 * mean_spot) ** 2 + 0.05 ** 2)
    i_spot = mean_spot + sig_spot * rng.standard_normal(n_spots)
    # shuffle the rows: files are not ordered by spot
    perm = rng.permutation(n_rows)
    iobs = np.ones(n_rows, dtype=np.float32)
    sig = np.ones(n_rows, dtype=np.float32)
    iobs[:n_spots] = i_spot                       # formatter.py:637-640
    sig[:n_spots] = sig_spot
    return {
        "refl_id": refl_id[perm], "image_id": spot_image[spot_of_row][perm].astype(np.int64),
        "file_id": np.zeros(n_rows, dtype=np.int64), "metadata": meta[perm],
        "intensities": iobs, "uncertainties": sig, "wavelength": wavelength[perm],
        "harmonic_id": spot_of_row[perm].astype(np.int64), "n_spots": n_spots,
        "centric": centric, "multiplicity": mult, "n_images": n_images, "f_true": f_true,
    }


def make_double_wilson(n_per_dataset, r_per_dataset, n_datasets=4, d=5, n_images=2000, r=0.99, seed=1234):
    """configs[3]: time-resolved merge, separate ASUs; parents = None,0,0,..; reflids[i] = i mod R0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R0 = r_per_dataset
    centric0, mult0, f0 = reflection_tables(rng, R0)
    centric = np.tile(centric0, n_datasets)
    mult = np.tile(mult0, n_datasets)
    f_true = [f0]
    for k in range(1, n_datasets):       # child correlated with the parent at r (doc/double_wilson.md)
        noise = _wilson_sample(rng, centric0, mult0.astype(np.float64))
        f_true.append(np.maximum(np.abs(r * f0 + np.sqrt(1 - r * r) * noise * rng.choice([-1.0, 1.0], R0)), 1e-3))
    f_true = np.concatenate(f_true)
    R = R0 * n_datasets
    N = n_per_dataset * n_datasets
    asu_of_row = np.repeat(np.arange(n_datasets), n_per_dataset)
    refl_local = rng.integers(0, R0, size=N)
    refl_id = (asu_of_row * R0 + refl_local).astype(np.int64)
    image_id = (asu_of_row * n_images + rng.integers(0, n_images, size=N)).astype(np.int64)
    meta = rng.standard_normal((N, d)).astype(np.float32)
    mean = _scale_and_noise(rng, meta.astype(np.float64), f_true[refl_id] ** 2)
    sig = np.sqrt((0.03 * mean) ** 2 + 0.05 ** 2)
    iobs = mean + sig * rng.standard_normal(N)
    asu_id = np.repeat(np.arange(n_datasets), R0).astype(np.int32)
    dw_parent = np.concatenate([np.full(R0, -2, dtype=np.int32)] +
                               [np.arange(R0, dtype=np.int32) for _ in range(1, n_datasets)])
    # reference bookkeeping (priors/wilson.py:112-137): reflids (parent id, local ids for roots), root flags
    reflids = np.concatenate([np.arange(R0)] + [np.arange(R0) for _ in range(1, n_datasets)]).astype(np.int64)
    root = asu_id == 0
    r_values = np.array([0.0] + [r] * (n_datasets - 1), dtype=np.float32)
    return {
        "refl_id": refl_id, "image_id": image_id, "file_id": asu_of_row.astype(np.int64), "metadata": meta,
        "intensities": iobs.astype(np.float32), "uncertainties": sig.astype(np.float32),
        "centric": centric, "multiplicity": mult, "n_images": n_images * n_datasets, "f_true": f_true,
        "asu_id": asu_id, "dw_parent": dw_parent, "reflids": reflids, "root": root, "r": r_values,
        "n_asu": n_datasets,
    }


# ----------------------------------------------------------------------------------------------------------------
# Two-phase generators for the multi-GPU benchmark: phase 1 draws only the INTEGER structure of the global problem
# (which observation belongs to which reflection / image / spot) -- every rank draws the same arrays from the same
# seed and runs the real partitioner on them (careless_b200.parallel) -- phase 2 attaches metadata and intensities to
# the rows a rank keeps.  Same distributions as make_mono / make_laue / make_double_wilson above.
# ----------------------------------------------------------------------------------------------------------------
def ids_mono(N, R, seed=1234):
    """configs[1]: refl_id = sorted uniform draw over R (multinomial counts, some reflections unobserved)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    counts = rng.multinomial(N, np.full(R, 1.0 / R))
    return {"refl_id": np.repeat(np.arange(R, dtype=np.int64), counts)}


def ids_stills(N, R, n_images, seed=1234):
    """configs[4]: image-major rows (as they come out of stills processing), N / n_images observations per image, and
    a NON-uniform popularity of the reflections (density ~ x^-1/3 over the reflection index: low-resolution reflections
    are observed several times more often than the weakest shell), so that a partition by reflection count would be
    unbalanced and the partitioner has to balance by observation count."""
    rng = np.random.Generator(np.random.PCG64(seed))
    refl_id = np.minimum((rng.random(N) ** 1.5 * R).astype(np.int64), R - 1)
    per = -(-N // n_images)
    image_id = np.repeat(np.arange(n_images, dtype=np.int64), per)[:N]
    return {"refl_id": refl_id, "image_id": image_id}


def ids_laue(n_rows, R, n_images, seed=1234):
    """configs[2]: spots of 1..5 harmonics (p = .84,.10,.04,.015,.005); the harmonics of a spot are the first m orders of
    one central ray (R / 5 rays x 5 orders), so the spot <-> reflection graph decomposes into rays (utils/laue.py:5-7)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.array([0.84, 0.10, 0.04, 0.015, 0.005])
    mean_len = float((p * np.arange(1, 6)).sum())
    lens = rng.choice(np.arange(1, 6), size=int(n_rows / mean_len) + 8, p=p)
    csum = np.cumsum(lens)
    n_spots = int(np.searchsorted(csum, n_rows, side="right"))
    lens = lens[:n_spots]
    rest = n_rows - int(lens.sum())
    if rest > 0:
        lens = np.concatenate([lens, np.ones(rest, dtype=lens.dtype)]); n_spots += rest
    spot_of_row = np.repeat(np.arange(n_spots, dtype=np.int64), lens)
    first = np.cumsum(lens) - lens
    order = np.arange(n_rows, dtype=np.int64) - np.repeat(first, lens)
    n_rays = R // 5
    ray = rng.integers(0, n_rays, size=n_spots)
    refl_id = ray[spot_of_row] * 5 + order
    spot_image = np.sort(rng.integers(0, n_images, size=n_spots))        # harmonic ids are image-major (formatter.py:617)
    return {"refl_id": refl_id, "harmonic_id": spot_of_row, "image_id": spot_image[spot_of_row].astype(np.int64), "n_spots": n_spots}


def ids_double_wilson(n_per_dataset, r_per_dataset, n_datasets=4, n_images=2500, seed=1234):
    """configs[3]: separate ASUs, parents = None,0,0,0; reflids[i] = i mod R0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R0, N = r_per_dataset, n_per_dataset * n_datasets
    asu_of_row = np.repeat(np.arange(n_datasets, dtype=np.int64), n_per_dataset)
    refl_id = asu_of_row * R0 + rng.integers(0, R0, size=N)
    image_id = asu_of_row * n_images + rng.integers(0, n_images, size=N)
    asu_id = np.repeat(np.arange(n_datasets), R0).astype(np.int32)
    dw_parent = np.concatenate([np.full(R0, -2, dtype=np.int32)] + [np.arange(R0, dtype=np.int32) for _ in range(1, n_datasets)])
    return {"refl_id": refl_id, "image_id": image_id, "asu_id": asu_id, "dw_parent": dw_parent,
            "r": np.array([0.0] + [0.99] * (n_datasets - 1), dtype=np.float32)}


def global_tables(R, seed=1234, n_datasets=1):
    """Per-reflection prior tables + true amplitudes of the GLOBAL problem (cheap: R <= a few million)."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    if n_datasets == 1:
        centric, mult, f_true = reflection_tables(rng, R)
        return {"centric": centric, "multiplicity": mult, "f_true": f_true}
    R0 = R // n_datasets
    centric0, mult0, f0 = reflection_tables(rng, R0)
    f = [f0]
    for _ in range(1, n_datasets):
        noise = _wilson_sample(rng, centric0, mult0.astype(np.float64))
        f.append(np.maximum(np.abs(0.99 * f0 + np.sqrt(1 - 0.99 ** 2) * noise * rng.choice([-1.0, 1.0], R0)), 1e-3))
    return {"centric": np.tile(centric0, n_datasets), "multiplicity": np.tile(mult0, n_datasets), "f_true": np.concatenate(f)}


def attach_payload(local, f_true_rows, d=5, seed=1234, laue=False):
    """Phase 2: metadata (N_local, d) ~ N(0,1), I = K_true F^2 + sigma N(0,1), sigma = sqrt((0.03 K F^2)^2 + 0.05^2) for the rows
    of one rank (`local` = the dict parallel.shard returned for id-only inputs; f_true_rows = F_true of each local row).
    Laue: per-spot sums over the local harmonic ids, stored in the reference's padded layout (formatter.py:637-640)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = len(local["refl_id"])
    meta = rng.standard_normal((n, d), dtype=np.float32)
    mean = np.exp(0.3 * meta[:, 0].astype(np.float64) - 0.2 * meta[:, 1].astype(np.float64) ** 2) * f_true_rows ** 2
    if laue:
        hid = local["harmonic_id"]
        n_spots = int(hid.max()) + 1
        mean_spot = np.bincount(hid, weights=mean, minlength=n_spots)
        sig_spot = np.sqrt((0.03 * mean_spot) ** 2 + 0.05 ** 2)
        iobs = np.ones(n, dtype=np.float32); sig = np.ones(n, dtype=np.float32)
        iobs[:n_spots] = mean_spot + sig_spot * rng.standard_normal(n_spots)
        sig[:n_spots] = sig_spot
    else:
        sig = np.sqrt((0.03 * mean) ** 2 + 0.05 ** 2)
        iobs = (mean + sig * rng.standard_normal(n)).astype(np.float32)
        sig = sig.astype(np.float32)
    local["metadata"], local["intensities"], local["uncertainties"] = meta, iobs, sig
    return local
