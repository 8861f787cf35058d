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

    inputs: dict with refl_id (+ any of image_id, metadata, intensities, uncertainties, wavelength, file_id, harmonic_id);
            per-row arrays that are absent are simply not produced (a caller may attach them to the local rows later)
    tables: dict with centric, multiplicity (+ sigma, dw_parent, asu_id); per-reflection arrays
    Returns (local_inputs, local_tables) with LOCAL refl / parent / spot indices and the global indices
    ``obs_index`` / ``refl_index`` that key the RNG (Laue: also ``spots``, the global harmonic_id of every local spot)."""
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
    iobs = None if inputs.get("intensities") is None else np.asarray(inputs["intensities"]).reshape(-1)
    sig = None if inputs.get("uncertainties") is None else np.asarray(inputs["uncertainties"]).reshape(-1)
    if laue:
        hid = np.asarray(inputs["harmonic_id"], dtype=np.int64).reshape(-1)[rows]
        spots, local_hid = np.unique(hid, return_inverse=True)
        out["harmonic_id"], out["spots"] = local_hid.astype(np.int64), spots
        if iobs is not None:
            li = np.ones(len(rows), dtype=np.float32); ls = np.ones(len(rows), dtype=np.float32)
            li[:len(spots)] = iobs[spots]; ls[:len(spots)] = sig[spots]      # formatter.py:637-640 layout, per rank
            out["intensities"], out["uncertainties"] = li, ls
    elif iobs is not None:
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
    """One data-parallel step driven from Python with CALLER-side all-reduces (``torch.distributed``): the form for a host
    application that owns its own process group.  ``grad_tensor`` / ``scalar_tensor`` are torch views of the engine's
    reduce buffers.  The engine's kernels run on ``engine.cfg.stream`` while ``dist.all_reduce`` runs on torch's current
    stream, so the two MUST be the same stream (create the engine with ``stream=torch.cuda.current_stream().cuda_stream``);
    anything else would let NCCL read the gradient buffer while the kernels are still writing it.
    The product path does not use this: ``Engine.init_comm`` moves the exchange into the library (one grouped NCCL
    all-reduce per step on the engine's own stream) and ``engine.step(n)`` then runs n steps without Python in the loop."""
    if getattr(grad_tensor, "is_cuda", False):
        import torch
        cur = torch.cuda.current_stream(grad_tensor.device).cuda_stream
        if int(engine.cfg.stream or 0) != int(cur):
            raise RuntimeError("allreduce_step: the engine must run on torch's current CUDA stream (EngineConfig.stream = "
                               "torch.cuda.current_stream().cuda_stream); otherwise the all-reduce is not ordered against the step's kernels")
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


# ----------------------------------------------------------------------------------------------------------------
# Process-level context of the product path: `torchrun --nproc-per-node N -m careless_b200.careless mono ...`
# ----------------------------------------------------------------------------------------------------------------
class DistContext:
    """Rank / world size of this process plus the HOST-side plumbing the product path needs around the library:
    handing rank 0's NCCL id to the other ranks and summing host arrays (gathering the partitioned surrogate and
    the per-observation predictions) over a gloo group.  The per-step exchange itself is inside the library."""

    def __init__(self, rank=0, world=1, local_rank=0, group=None):
        self.rank, self.world, self.local_rank, self.group = int(rank), int(world), int(local_rank), group

    @property
    def active(self):
        return self.world > 1

    def broadcast_bytes(self, payload, src=0):
        if not self.active:
            return payload
        import torch
        import torch.distributed as dist
        t = torch.zeros(len(payload), dtype=torch.uint8)
        if self.rank == src:
            t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
        dist.broadcast(t, src=src, group=self.group)
        return bytes(t.numpy().tobytes())

    def allsum(self, arr):
        """Element-wise sum of a host array over all ranks (float64 on the wire for float inputs)."""
        if not self.active:
            return arr
        import torch
        import torch.distributed as dist
        a = np.ascontiguousarray(arr)
        wire = a.astype(np.float64) if a.dtype.kind == "f" else a.astype(np.int64)
        t = torch.from_numpy(wire)
        dist.all_reduce(t, group=self.group)
        return t.numpy().astype(a.dtype)

    def barrier(self):
        if self.active:
            import torch.distributed as dist
            dist.barrier(group=self.group)


_CTX = None


def context():
    """The process's DistContext: from RANK / WORLD_SIZE / LOCAL_RANK (torchrun) when WORLD_SIZE > 1, else single-process.
    A gloo group is created for the host-side exchanges (an existing NCCL-only process group gets a gloo sub-group)."""
    global _CTX
    if _CTX is not None:
        return _CTX
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        _CTX = DistContext()
        return _CTX
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
    group = None
    if not dist.is_initialized():
        dist.init_process_group("gloo", rank=rank, world_size=world)
    elif "gloo" not in str(dist.get_backend()):
        group = dist.new_group(backend="gloo")
    _CTX = DistContext(rank, world, local, group)
    return _CTX


def set_context(ctx):
    """Tests and embedding applications: install an explicit context (None resets to the environment's)."""
    global _CTX
    _CTX = ctx


def init_engine_comm(engine, ctx):
    """Create the library-side NCCL communicator of ``engine`` (one per engine; rank 0's id is broadcast on the host)."""
    if not ctx.active:
        return
    from . import _lib as L
    payload = L.comm_unique_id() if ctx.rank == 0 else bytes(128)
    engine.init_comm(ctx.broadcast_bytes(payload, src=0))


def partition(n_refl, refl_id, world, harmonic_id=None, dw_parent=None, extra=()):
    """Rank of every reflection for this job.  ``extra`` = further (refl_id, harmonic_id) pairs whose rows must obey the
    same partition (the held-out rows of --test-fraction: their spots link reflections exactly like training spots)."""
    refl_id = np.asarray(refl_id, dtype=np.int64).reshape(-1)
    rid, hid = [refl_id], [None if harmonic_id is None else np.asarray(harmonic_id, dtype=np.int64).reshape(-1)]
    for r2, h2 in extra:
        rid.append(np.asarray(r2, dtype=np.int64).reshape(-1))
        hid.append(None if h2 is None else np.asarray(h2, dtype=np.int64).reshape(-1))
    if harmonic_id is not None:
        off, parts = 0, []
        for h in hid:
            parts.append(h + off)
            off += int(h.max()) + 1 if len(h) else 0
        groups = reflection_groups(n_refl, np.concatenate(rid), np.concatenate(parts), dw_parent)
    else:
        groups = reflection_groups(n_refl, None, None, dw_parent)
    weight = np.bincount(np.concatenate(rid), minlength=n_refl)
    return assign_ranks(groups, weight, world)
