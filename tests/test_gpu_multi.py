"""Multi-GPU parity (needs >= 2 B200s on the box; skipped on a single-GPU box): a reflection-partitioned
2-rank run over NCCL must reproduce the single-GPU run (same Philox draws by construction: counters are
global indices)."""
import os
import socket

import numpy as np
import pytest

from careless_b200 import parallel, synth
from careless_b200.engine import Engine, EngineConfig

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cfg(kind, R, n_images, **kw):
    base = dict(n_refl=R, n_meta=3, mlp_width=10, mlp_layers=4, mc_samples=2, seed=99, learning_rate=1e-2)
    if kind == "mono":
        base.update(likelihood="studentt", dof=8.0, image_scales=True, n_images=n_images)
    elif kind == "laue":
        base.update(laue=True, image_scales=True, n_images=n_images)
    else:
        base.update(prior="double_wilson", n_asu=3, optimize_dw_r=True)
    base.update(kw)
    return base


def _problem(kind):
    if kind == "mono":
        p = synth.make_mono(20000, 1500, d=3, n_images=16, seed=31)
    elif kind == "laue":
        p = synth.make_laue(20000, 2000, d=3, n_images=16, seed=32)
        ray = np.random.default_rng(1).integers(0, 500, size=p["n_spots"])
        order = np.zeros(20000, dtype=np.int64)
        srt = np.argsort(p["harmonic_id"], kind="stable")
        hs = p["harmonic_id"][srt]
        first = np.r_[0, np.nonzero(np.diff(hs))[0] + 1]
        pos = np.arange(20000) - np.repeat(first, np.diff(np.r_[first, 20000]))
        order[srt] = pos % 4
        p["refl_id"] = ray[p["harmonic_id"]] * 4 + order
    else:
        p = synth.make_double_wilson(7000, 500, n_datasets=3, d=3, n_images=6, r=0.9, seed=33)
    return p


def _tables(p, kind):
    t = {"centric": p["centric"], "multiplicity": p["multiplicity"]}
    if kind == "dw":
        t.update(dw_parent=p["dw_parent"], asu_id=p["asu_id"])
    return t


def _worker(rank, world, port, kind, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        stream = torch.cuda.Stream(device=rank)
        torch.cuda.set_stream(stream)
        p = _problem(kind)
        R = len(p["centric"]); N = len(p["refl_id"])
        tables = _tables(p, kind)
        groups = parallel.reflection_groups(R, p["refl_id"], p.get("harmonic_id") if kind == "laue" else None, tables.get("dw_parent"))
        ranks = parallel.assign_ranks(groups, np.bincount(p["refl_id"], minlength=R), world)
        li, lt = parallel.shard(p, tables, ranks, rank, laue=(kind == "laue"))
        cfg = EngineConfig(**_cfg(kind, len(lt["refl_index"]), int(p["n_images"]), n_refl_total=R, device=rank,
                                  stream=stream.cuda_stream, rank=rank, world_size=world))
        eng = Engine(cfg)
        eng.set_observations(li["refl_id"], li.get("image_id"), li["metadata"], li["intensities"], li["uncertainties"],
                             harmonic_id=li.get("harmonic_id"), obs_index=li["obs_index"], n_rows_total=N)
        eng.set_prior(lt["centric"], lt["multiplicity"], None, dw_parent=lt.get("dw_parent"), asu_id=lt.get("asu_id"),
                      r=p.get("r") if kind == "dw" else None, refl_index=lt["refl_index"])
        g, s = parallel.reduce_tensors(eng, f"cuda:{rank}")
        hist = [parallel.allreduce_step(eng, dist, g, s) for _ in range(4)]
        ret[rank] = {"hist": hist, "mine": lt["refl_index"], "loc": eng.get_params("sf_loc_raw"), "mlp": eng.get_params("mlp"),
                     "img": eng.get_params("image_scales") if cfg.image_scales else None,
                     "r": eng.get_params("dw_r_logit") if kind == "dw" else None}
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind", ["mono", "laue", "dw"])
def test_two_gpus_match_one(kind):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, kind, ret), nprocs=world, join=True)
        ret = dict(ret)
    p = _problem(kind)
    R = len(p["centric"])
    eng = Engine(EngineConfig(**_cfg(kind, R, int(p["n_images"]))))
    eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"],
                         harmonic_id=p.get("harmonic_id") if kind == "laue" else None)
    eng.set_prior(p["centric"], p["multiplicity"], None, dw_parent=p.get("dw_parent") if kind == "dw" else None,
                  asu_id=p.get("asu_id") if kind == "dw" else None, r=p.get("r") if kind == "dw" else None)
    hist = eng.step(4)
    for r in range(world):
        for i in range(4):
            for k in ("loss", "NLL", "F KLDiv", "Grad Norm"):
                a, b = ret[r]["hist"][i][k], hist[i][k]
                assert abs(a - b) <= 2e-5 * abs(b) + 1e-7, (r, i, k, a, b)
        assert np.allclose(ret[r]["mlp"], eng.get_params("mlp"), rtol=2e-4, atol=2e-5)
        assert np.allclose(ret[r]["loc"], eng.get_params("sf_loc_raw")[ret[r]["mine"]], rtol=2e-4, atol=2e-5)
        if ret[r]["img"] is not None:
            assert np.allclose(ret[r]["img"], eng.get_params("image_scales"), rtol=2e-4, atol=2e-5)
        if ret[r]["r"] is not None:
            assert np.allclose(ret[r]["r"], eng.get_params("dw_r_logit"), rtol=2e-4, atol=2e-5)
    assert np.array_equal(np.sort(np.concatenate([ret[0]["mine"], ret[1]["mine"]])), np.arange(R))
    eng.close()
