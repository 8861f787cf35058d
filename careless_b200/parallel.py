"""Reflection-partitioned data parallelism; one process per GPU.

The reference is single-device (``careless/parser.py:25-40``); this is the B200-native addition.
Surrogate entries (unique reflections) are partitioned over ranks and every observation goes to the
rank that owns its reflection, so sampling, the refl_id gather, the segmented reduction of dL/dz_f,
the prior KL and the surrogate's Adam update are rank-local.  Only the replicated parameters' gradients
(scale MLP, image scales, DoubleWilson r) and a handful of float64 scalars are all-reduced (NCCL) per
step:   step_begin -> all_reduce(grads f32) -> step_norms -> all_reduce(scalars f64) -> step_end.

Partition keys that keep every segment whole:
* mono          : the reflection itself;
* Laue          : connected components of the spot<->reflection graph (= the central ray H/gcd(H),
                  ``careless/utils/laue.py:5-7``), so all harmonics of a spot share a rank;
* DoubleWilson  : the root ancestor of a reflection (``priors/wilson.py:112-128``), so the parent
                  gather / gradient scatter stays local.
Draws are keyed by GLOBAL observation / reflection indices (Philox counters), so results do not depend
on the number of ranks.
"""
from __future__ import annotations

import numpy as np


def reflection_groups(n_refl, refl_id=None, harmonic_id=None, dw_parent=None):
    """Group id per reflection such that reflections in one group must share a rank."""
    group = np.arange(n_refl, dtype=np.int64)
    if dw_parent is not None:
        parent = np.asarray(dw_parent, dtype=np.int64)
        cur = np.where(parent >= 0, parent, np.arange(n_refl))
        for _ in range(64):                      # follow parents to the root ancestor (trees are shallow)
            nxt = np.where(parent[cur] >= 0, parent[cur], cur)
            if np.array_equal(nxt, cur):
                break
            cur = nxt
        group = group[cur]
    if harmonic_id is not None:
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        refl_id = np.asarray(refl_id, dtype=np.int64).reshape(-1)
        hid = np.asarray(harmonic_id, dtype=np.int64).reshape(-1)
        n_spots = int(hid.max()) + 1
        g = group[refl_id]                       # compose with the DW grouping if both are present
        a = coo_matrix((np.ones(len(hid), dtype=np.int8), (g, n_refl + hid)), shape=(n_refl + n_spots,) * 2)
        _, labels = connected_components(a, directed=False)
        group = labels[group].astype(np.int64)
    return group


def assign_ranks(group, weight, world_size):
    """Rank per reflection: groups are cut into ``world_size`` contiguous runs of ~equal observation count.

    ``weight`` = observations per reflection.  Group ids are processed in order of first appearance so that
    the assignment is deterministic; a group is never split."""
    group = np.asarray(group, dtype=np.int64)
    uniq, inv = np.unique(group, return_inverse=True)
    gw = np.bincount(inv, weights=np.asarray(weight, dtype=np.float64), minlength=len(uniq))
    csum = np.cumsum(gw)
    total = csum[-1] if len(csum) else 0.0
    mid = csum - 0.5 * gw
    grank = np.minimum((mid / max(total, 1e-300) * world_size).astype(np.int64), world_size - 1)
    return grank[inv].astype(np.int32)


def shard(inputs, tables, rank_of_refl, rank, laue=False):
    """Rows and tables of one rank.

    inputs: dict with refl_id, image_id, metadata, intensities, uncertainties (+ harmonic_id)
    tables: dict with centric, multiplicity (+ sigma, dw_parent, asu_id); per-reflection arrays
    Returns (local_inputs, local_tables) with LOCAL refl / parent / spot indices and the global indices
    ``obs_index`` / ``refl_index`` that key the RNG."""
    rank_of_refl = np.asarray(rank_of_refl)
    refl_id = np.asarray(inputs["refl_id"], dtype=np.int64).reshape(-1)
    mine_r = np.nonzero(rank_of_refl == rank)[0]
    g2l = np.full(len(rank_of_refl), -1, dtype=np.int64)
    g2l[mine_r] = np.arange(len(mine_r))
    rows = np.nonzero(rank_of_refl[refl_id] == rank)[0]
    out = {"refl_id": g2l[refl_id[rows]], "obs_index": rows.astype(np.int64), "n_rows_total": len(refl_id)}
    for k in ("image_id", "metadata", "wavelength", "file_id"):
        if inputs.get(k) is not None:
            out[k] = np.asarray(inputs[k])[rows]
    iobs = np.asarray(inputs["intensities"]).reshape(-1)
    sig = np.asarray(inputs["uncertainties"]).reshape(-1)
    if laue:
        hid = np.asarray(inputs["harmonic_id"], dtype=np.int64).reshape(-1)[rows]
        spots, local_hid = np.unique(hid, return_inverse=True)
        li = np.ones(len(rows), dtype=np.float32); ls = np.ones(len(rows), dtype=np.float32)
        li[:len(spots)] = iobs[spots]; ls[:len(spots)] = sig[spots]      # formatter.py:637-640 layout, per rank
        out["harmonic_id"], out["intensities"], out["uncertainties"] = local_hid.astype(np.int64), li, ls
    else:
        out["intensities"], out["uncertainties"] = iobs[rows], sig[rows]
    t = {"refl_index": mine_r.astype(np.int64)}
    for k in ("centric", "multiplicity", "sigma", "asu_id"):
        if tables.get(k) is not None and np.ndim(tables[k]) > 0:
            t[k] = np.asarray(tables[k])[mine_r]
        elif tables.get(k) is not None:
            t[k] = tables[k]
    if tables.get("dw_parent") is not None:
        p = np.asarray(tables["dw_parent"], dtype=np.int64)[mine_r]
        lp = np.where(p >= 0, g2l[np.maximum(p, 0)], p)
        if np.any((p >= 0) & (lp < 0)):
            raise ValueError("partition splits a DoubleWilson parent from its child")
        t["dw_parent"] = lp.astype(np.int32)
    return out, t


def allreduce_step(engine, dist, grad_tensor, scalar_tensor, want_metrics=True, u_f=None, eps_s=None):
    """One data-parallel step.  ``engine`` exposes step_begin/step_norms/step_end; ``grad_tensor`` and
    ``scalar_tensor`` are torch views of its reduce buffers (device memory for NCCL)."""
    engine.step_begin(u_f, eps_s)
    dist.all_reduce(grad_tensor)
    engine.step_norms()
    dist.all_reduce(scalar_tensor)
    return engine.step_end(want_metrics)


class DeviceView:
    """``__cuda_array_interface__`` shim so torch can wrap a device pointer owned by the C library."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def reduce_tensors(engine, device):
    import torch
    pf, nf, pd, nd = engine.reduce_buffers()
    g = torch.as_tensor(DeviceView(pf, nf, "<f4"), device=device)
    s = torch.as_tensor(DeviceView(pd, nd, "<f8"), device=device)
    return g, s
