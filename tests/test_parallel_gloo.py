"""Host-side logic of the reflection-partitioned data parallelism, on CPU with gloo (world_size 2).

The CUDA engine cannot run here, so the per-rank engine is replaced by an oracle-backed stand-in with
the same step_begin / step_norms / step_end protocol and reduce buffers; what is under test is the
partitioner, the sharding of inputs and tables, and the all-reduce protocol of careless_b200.parallel:
the 2-rank result must equal the single-process oracle.
"""
import os
import socket

import numpy as np
import pytest
import torch

from careless_b200 import parallel, synth
from oracle import model as om

import _util as U


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class OracleShardEngine:
    """Stand-in for careless_b200.Engine on one shard (test infrastructure)."""

    def __init__(self, params, inputs, tables, cfg, rank):
        self.cfg, self.rank = cfg, rank
        self.inputs = inputs
        self.prior = om.PriorData(tables["centric"], tables["multiplicity"], tables.get("sigma", 1.0),
                                  reflids=tables.get("reflids"), root=tables.get("root"), asu_ids=tables.get("asu_id"),
                                  r=tables.get("r"))
        self.params = params
        self.rep_names = [k for k in params if not k.startswith("sf_")]
        n = sum(params[k].numel() for k in self.rep_names)
        self.grad_buf = torch.zeros(n, dtype=torch.float32)
        self.scalar_buf = torch.zeros(4, dtype=torch.float64)

    def step_begin(self, u_f, eps_s):
        out = om.forward({k: v.clone().requires_grad_(True) for k, v in self.params.items()}, self.inputs, self.prior, self.cfg, u_f, eps_s)
        self._leaves = None
        leaves = {k: v.clone().requires_grad_(True) for k, v in self.params.items()}
        out = om.forward(leaves, self.inputs, self.prior, self.cfg, u_f, eps_s)
        g = torch.autograd.grad(out["loss"], list(leaves.values()), allow_unused=True)
        self.g = {k: (torch.zeros_like(leaves[k]) if gi is None else gi) for k, gi in zip(leaves, g)}
        self.grad_buf[:] = torch.cat([self.g[k].reshape(-1) for k in self.rep_names]).float()
        self.kl, self.nll = float(out["kl"].detach()), float(out["nll"].detach())

    def step_norms(self):
        rep = float((self.grad_buf.double() ** 2).sum()) if self.rank == 0 else 0.0     # replicated: count once
        loc = sum(float((self.g[k] ** 2).sum()) for k in ("sf_loc_raw", "sf_scale_raw"))
        self.scalar_buf[:] = torch.tensor([self.kl, self.nll, rep + loc, 0.0], dtype=torch.float64)

    def step_end(self, want_metrics=True):
        kl, nll, ss, _ = [float(x) for x in self.scalar_buf]
        return {"loss": kl + nll, "NLL": nll, "F KLDiv": kl, "Grad Norm": ss ** 0.5}


def _worker(rank, world, port, kind, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, cfg, tables, u, e = _problem(kind)
        R = cfg.n_refl
        groups = parallel.reflection_groups(R, p["refl_id"], p.get("harmonic_id") if cfg.laue else None, tables.get("dw_parent"))
        ranks = parallel.assign_ranks(groups, np.bincount(p["refl_id"], minlength=R), world)
        li, lt = parallel.shard(p, tables, ranks, rank, laue=cfg.laue)
        if cfg.prior == "double_wilson":
            lt["reflids"] = np.where(lt["dw_parent"] >= 0, lt["dw_parent"], np.where(lt["dw_parent"] == -2, 0, -1))
            lt["root"] = lt["dw_parent"] == -2
            lt["r"] = tables["r"]
        params = _params(cfg, tables)
        mine = lt["refl_index"]
        lparams = {k: (v[mine] if k.startswith("sf_") else v) for k, v in params.items()}
        lcfg = om.ModelConfig(**{**cfg.__dict__, "n_refl": len(mine)})
        eng = OracleShardEngine(lparams, li, lt, lcfg, rank)
        m = parallel.allreduce_step(eng, dist, eng.grad_buf, eng.scalar_buf, u_f=u[:, mine], eps_s=e[:, li["obs_index"]])
        ret[rank] = {"metrics": m, "rep": eng.grad_buf.numpy().copy(), "mine": mine,
                     "g_loc": eng.g["sf_loc_raw"].numpy().copy(), "n_rows": len(li["refl_id"])}
    finally:
        dist.destroy_process_group()


def _problem(kind):
    rng = np.random.default_rng(11)
    if kind == "mono":
        p = synth.make_mono(1200, 150, d=3, n_images=6, seed=21)
        cfg = om.ModelConfig(n_refl=150, n_meta=3, mlp_width=5, mlp_layers=2, likelihood="studentt", dof=6.0,
                             image_scales=True, n_images=6, mc_samples=2)
        tables = {"centric": p["centric"], "multiplicity": p["multiplicity"]}
    elif kind == "laue":
        p = synth.make_laue(1500, 200, d=3, n_images=8, seed=22)
        # make harmonics of a spot lie on a common "ray" so that components stay small: refl = ray*4 + order
        ray = np.random.default_rng(1).integers(0, 50, size=p["n_spots"])
        order = np.zeros(1500, dtype=np.int64)
        for k in np.unique(p["harmonic_id"]):
            rows = np.nonzero(p["harmonic_id"] == k)[0]
            order[rows] = np.arange(len(rows)) % 4
        p["refl_id"] = ray[p["harmonic_id"]] * 4 + order
        cfg = om.ModelConfig(n_refl=200, n_meta=3, mlp_width=5, mlp_layers=2, laue=True)
        tables = {"centric": p["centric"], "multiplicity": p["multiplicity"]}
    else:
        p = synth.make_double_wilson(500, 60, n_datasets=3, d=3, n_images=4, r=0.9, seed=23)
        cfg = om.ModelConfig(n_refl=180, n_meta=3, mlp_width=5, mlp_layers=2, prior="double_wilson", optimize_dw_r=True)
        tables = {"centric": p["centric"], "multiplicity": p["multiplicity"], "dw_parent": p["dw_parent"],
                  "asu_id": p["asu_id"], "r": p["r"], "reflids": p["reflids"], "root": p["root"]}
    S, R, N = cfg.mc_samples, cfg.n_refl, len(p["refl_id"])
    u = rng.random((S, R)); e = rng.standard_normal((S, N))
    return p, cfg, tables, u, e


def _params(cfg, tables):
    prior = om.PriorData(tables["centric"], tables["multiplicity"], 1.0, r=tables.get("r"))
    rng = np.random.default_rng(5)
    p = om.init_params(cfg, prior)
    return {k: v + 0.05 * torch.as_tensor(rng.standard_normal(tuple(v.shape))) for k, v in p.items()}


@pytest.mark.parametrize("kind", ["mono", "laue", "dw"])
def test_two_rank_step_equals_single_process(kind):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, kind, ret), nprocs=world, join=True)
        ret = dict(ret)
    p, cfg, tables, u, e = _problem(kind)
    prior = om.PriorData(tables["centric"], tables["multiplicity"], 1.0, reflids=tables.get("reflids"),
                         root=tables.get("root"), asu_ids=tables.get("asu_id"), r=tables.get("r"))
    params = _params(cfg, tables)
    metrics, g, _ = om.loss_and_grads(params, p, prior, cfg, u, e)
    assert ret[0]["n_rows"] + ret[1]["n_rows"] == len(p["refl_id"])
    assert min(ret[0]["n_rows"], ret[1]["n_rows"]) > 0.25 * len(p["refl_id"])          # balanced enough
    for r in range(world):
        for k in ("loss", "NLL", "F KLDiv", "Grad Norm"):
            assert abs(ret[r]["metrics"][k] - metrics[k]) <= 1e-6 * abs(metrics[k]) + 1e-9, (r, k)
        rep_ref = torch.cat([g[k].reshape(-1) for k in params if not k.startswith("sf_")]).numpy()
        assert np.allclose(ret[r]["rep"], rep_ref, rtol=1e-5, atol=1e-5 * np.abs(rep_ref).max())
        assert np.allclose(ret[r]["g_loc"], g["sf_loc_raw"].numpy()[ret[r]["mine"]], rtol=1e-9, atol=1e-12)
    assert np.array_equal(np.sort(np.concatenate([ret[0]["mine"], ret[1]["mine"]])), np.arange(cfg.n_refl))


def test_partition_keeps_segments_whole():
    p = synth.make_double_wilson(400, 50, n_datasets=4, d=2, n_images=3, r=0.9, seed=3)
    groups = parallel.reflection_groups(200, dw_parent=p["dw_parent"])
    assert np.array_equal(groups, np.tile(np.arange(50), 4))          # root ancestor = i mod R0
    ranks = parallel.assign_ranks(groups, np.bincount(p["refl_id"], minlength=200), 4)
    assert np.array_equal(ranks[:50], ranks[50:100]) and np.array_equal(ranks[:50], ranks[150:])
    assert set(ranks) == {0, 1, 2, 3}
    with pytest.raises(ValueError):
        bad = ranks.copy(); bad[60] = (bad[10] + 1) % 4
        parallel.shard(p, {"centric": p["centric"], "multiplicity": p["multiplicity"], "dw_parent": p["dw_parent"]}, bad, int(bad[60]))
