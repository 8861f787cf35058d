"""GPU tests of the device-side row preparation (csrc/clb_prep.cuh): the rows the GPU sorts / pads / gathers inside
clb_set_observations are BIT-IDENTICAL to the host preparation (clb_prepare_rows: stable counting sort, the reference's
grouping of io/formatter.py:145, :617), for every row order, with ragged / empty / maximum-length groups, and the error
messages name the same first offending row."""
import numpy as np
import pytest

from careless_b200 import _lib as L
from careless_b200 import synth
from careless_b200.engine import Engine, EngineConfig
from test_abi_and_prep import _prepare

pytestmark = pytest.mark.gpu

ARRAYS = ("refl", "image", "spot", "oidx", "meta", "iobs", "sig")


def _engine(p, *, laue=False, image_layers=0, image_scales=False, width=8, likelihood="normal", dof=None, **kw):
    cfg = EngineConfig(n_refl=len(p["centric"]), n_meta=p["metadata"].shape[1], mlp_width=width, mlp_layers=2, n_images=int(p["n_images"]),
                       image_scales=image_scales, image_layers=image_layers, laue=laue, likelihood=likelihood, dof=dof, **kw)
    return Engine(cfg)


def _compare(p, *, laue=False, image_layers=0, image_scales=False, width=8, likelihood="normal", dof=None, obs_index=None, n_total=0):
    eng = _engine(p, laue=laue, image_layers=image_layers, image_scales=image_scales, width=width, likelihood=likelihood, dof=dof)
    try:
        eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"],
                             harmonic_id=p.get("harmonic_id") if laue else None, obs_index=obs_index, n_rows_total=n_total)
        dev = eng.download_rows()
    finally:
        eng.close()
    host = _prepare(p, len(p["centric"]), laue=laue, likelihood=1 if likelihood == "studentt" else 0, dof=dof or 0.0, obs_index=obs_index,
                    image_tile=128 if image_layers else 0)
    assert dev["refl"].shape == host["refl"].shape
    for k in ARRAYS:
        if k == "spot" and not laue:
            continue
        assert np.array_equal(dev[k], host[k]), f"{k} differs between the device and the host preparation"
    assert dev["ll_const"] == host["ll_const"]
    return dev


def _shuffled(p, seed, keys=("refl_id", "image_id", "metadata", "intensities", "uncertainties")):
    perm = np.random.default_rng(seed).permutation(len(p["refl_id"]))
    for k in keys:
        p[k] = p[k][perm]
    return p


@pytest.mark.parametrize("n,R", [(1, 1), (31, 4), (1000, 64), (4096, 1), (4097, 300), (70_001, 5000), (300_000, 70_000)])
def test_mono_rows_match_the_host_stable_sort(n, R):
    # several blocks of the radix sort (4096 rows each), 1..3 passes (R up to 2^17), a single key, ragged tails
    p = _shuffled(synth.make_mono(n, R, d=3, n_images=7, seed=n), seed=n + 1)
    dev = _compare(p)
    order = np.argsort(p["refl_id"], kind="stable")
    assert np.array_equal(dev["oidx"][:n], order)


def test_mono_with_obs_index_and_image_scales():
    p = _shuffled(synth.make_mono(5000, 333, d=5, n_images=11, seed=2), seed=3)
    oi = np.random.default_rng(4).permutation(5000)        # (the host-only entry point takes n_rows_total = n_rows)
    _compare(p, image_scales=True, obs_index=oi, n_total=5000)


@pytest.mark.parametrize("n,R,lik", [(3000, 200, "normal"), (50_000, 3000, "studentt"), (33, 5, "normal")])
def test_laue_rows_match(n, R, lik):
    p = synth.make_laue(n, R, d=2, n_images=9, seed=5)
    _compare(p, laue=True, likelihood=lik, dof=4.0 if lik == "studentt" else None)


def test_laue_spot_of_exactly_32_harmonics_and_gaps():
    p = synth.make_laue(2000, 150, d=2, n_images=4, seed=8)
    hid = p["harmonic_id"].copy()
    k = hid[100]
    hid[hid == k] = k + 1                         # move the spot's own rows away, then give it exactly 32 (the maximum)
    hid[100:132] = k
    p["harmonic_id"] = hid
    _compare(p, laue=True)


@pytest.mark.parametrize("laue", [False, True])
def test_image_layer_rows_match(laue):
    p = synth.make_laue(6000, 300, d=2, n_images=13, seed=6) if laue else _shuffled(synth.make_mono(6000, 300, d=2, n_images=13, seed=6), seed=7)
    dev = _compare(p, laue=laue, image_layers=1, width=10)
    live = dev["refl"] >= 0
    tiles = np.arange(len(live)) // 128
    for t in np.unique(tiles[live]):
        assert len(np.unique(dev["image"][live & (tiles == t)])) == 1


def test_images_without_rows_and_single_row_images():
    p = synth.make_mono(2000, 100, d=2, n_images=40, seed=9)
    img = p["image_id"].copy()
    img[img == 5] = 6; img[img == 17] = 0          # images 5 and 17 have no rows
    idx39 = np.nonzero(img == 39)[0]; img[idx39[1:]] = 38        # image 39 keeps a single row
    p["image_id"] = img
    _compare(_shuffled(p, seed=10), image_layers=2, width=10)


def test_rows_kept_as_given_with_order_none():
    """CLB_ORDER_NONE: the caller's row order is kept (no sort on either path)."""
    p = _shuffled(synth.make_mono(2000, 100, d=2, n_images=5, seed=21), seed=22)
    eng = _engine(p)
    try:
        eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"], order=L.ORDER_NONE)
        dev = eng.download_rows()
    finally:
        eng.close()
    assert np.array_equal(dev["refl"][:2000], p["refl_id"]) and np.array_equal(dev["oidx"][:2000], np.arange(2000))
    assert np.array_equal(dev["iobs"][:2000], p["intensities"])


def test_device_prep_reports_the_same_first_bad_row():
    p = synth.make_mono(10_000, 100, d=2, n_images=3, seed=1)
    rid = p["refl_id"].copy(); rid[7777] = 100; rid[4321] = -1
    p["refl_id"] = rid
    eng = _engine(p)
    try:
        with pytest.raises(L.ClbError, match=r"refl_id\[4321\]=-1 outside \[0,100\)"):
            eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"])
    finally:
        eng.close()
    q = synth.make_laue(400, 20, d=2, n_images=3, seed=1)
    q["harmonic_id"] = np.zeros(400, dtype=np.int64)       # one spot with 400 harmonics
    eng = _engine(q, laue=True)
    try:
        with pytest.raises(L.ClbError, match="at most 32"):
            eng.set_observations(q["refl_id"], q["image_id"], q["metadata"], q["intensities"], q["uncertainties"], harmonic_id=q["harmonic_id"])
    finally:
        eng.close()


def test_host_prep_switch_and_reupload(monkeypatch):
    """CLB_DEVICE_PREP=0 takes the host path; after a device-side preparation the pinned mirror for re-uploads is filled on demand
    and a step after upload / prefetch sees the same rows."""
    p = _shuffled(synth.make_mono(3000, 200, d=3, n_images=5, seed=11), seed=12)
    monkeypatch.setenv("CLB_DEVICE_PREP", "0")
    eng = _engine(p)
    eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"])
    host_rows = eng.download_rows(); eng.close()
    monkeypatch.delenv("CLB_DEVICE_PREP")
    eng = _engine(p)
    try:
        eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"])
        eng.set_prior(p["centric"], p["multiplicity"])
        eng.upload_observations()
        dev_rows = eng.download_rows()
        for k in ARRAYS:
            assert np.array_equal(dev_rows[k], host_rows[k]), k
        eng.prefetch_observations()
        assert np.isfinite(eng.eval()["NLL"])          # the evaluation switches to the prefetched buffer
        dev_rows = eng.download_rows()
        for k in ARRAYS:
            assert np.array_equal(dev_rows[k], host_rows[k]), k
    finally:
        eng.close()


def test_full_size_device_prep_is_a_stable_sort():
    """10 M rows / 500 k reflections (BASELINE configs[1]): the device rows are a stable sort by refl_id -- checked through
    size-independent properties (sortedness, stability inside every reflection, permutation checksum, payload follows the row)."""
    n, R = 10_000_000, 500_000
    rng = np.random.default_rng(0)
    refl = rng.integers(0, R, n, dtype=np.int64)
    meta = rng.standard_normal((n, 5)).astype(np.float32)
    iobs = rng.standard_normal(n).astype(np.float32); sig = np.abs(iobs) + 1.0
    cfg = EngineConfig(n_refl=R, n_meta=5, mlp_width=32, mlp_layers=2)
    eng = Engine(cfg)
    try:
        eng.set_observations(refl, None, meta, iobs, sig)
        dev = eng.download_rows()
    finally:
        eng.close()
    r, o = dev["refl"][:n].astype(np.int64), dev["oidx"][:n].astype(np.int64)
    assert np.all(np.diff(r) >= 0)
    same = np.diff(r) == 0
    assert np.all(np.diff(o)[same] > 0)                                  # stable: input order inside a reflection
    assert o.sum() == n * (n - 1) // 2 and len(np.unique(o[::997])) == len(o[::997])
    assert np.array_equal(r, refl[o]) and np.array_equal(dev["iobs"][:n], iobs[o]) and np.array_equal(dev["meta"][3, :n], meta[o, 3])
    print(f"device row preparation of {n} rows: {dev['prep_ms']:.1f} ms")
