"""CPU tests: the C-ABI library loads and exports every declared symbol; the host prep
(sort / pad into the device layout) is bit-exact; the product fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from careless_b200 import _lib as L
from careless_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "careless_b200.h")).read()
    declared = set(re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", header))
    declared -= {"clb_status"}
    lib = L.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in L.SYMBOLS, f"{name} has no ctypes prototype"
    assert set(L.SYMBOLS) <= declared | {"clb_abi_version"}
    assert lib.clb_abi_version() == L.ABI_VERSION


def test_config_struct_matches_header_field_order():
    header = open(os.path.join(ROOT, "include", "careless_b200.h")).read()
    body = header[header.index("typedef struct {", header.index("flattened")):header.index("} clb_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).replace("typedef struct {", "")
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt or stmt.startswith("typedef"):
            continue
        parts = stmt.replace("*", " ").split()
        for nm in " ".join(parts[1:]).split(","):
            names.append(nm.strip())
    assert names == [f[0] for f in L.clb_config._fields_]


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a CPU-only box")
def test_no_cpu_fallback():
    from careless_b200 import ClbError, Engine, EngineConfig
    with pytest.raises(ClbError) as ei:
        Engine(EngineConfig(n_refl=10, n_meta=2, mlp_width=4, mlp_layers=1))
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def _prepare(p, n_refl, laue, order=L.ORDER_AUTO, likelihood=0, dof=0.0, obs_index=None, image_tile=0):
    lib = L.load()
    n = len(p["refl_id"]); d = p["metadata"].shape[1]
    refl = np.ascontiguousarray(p["refl_id"], dtype=np.int64)
    img = np.ascontiguousarray(p["image_id"], dtype=np.int64)
    meta = np.ascontiguousarray(p["metadata"], dtype=np.float32)
    iobs = np.ascontiguousarray(p["intensities"], dtype=np.float32)
    sig = np.ascontiguousarray(p["uncertainties"], dtype=np.float32)
    hid = np.ascontiguousarray(p["harmonic_id"], dtype=np.int64) if laue else None
    oi = None if obs_index is None else np.ascontiguousarray(obs_index, dtype=np.int64)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    npad = C.c_int64(); llc = C.c_double()
    args = [n, n_refl, d, int(p["n_images"]), int(laue), likelihood, dof, ptr(refl), ptr(img), ptr(meta), ptr(iobs), ptr(sig),
            ptr(hid), ptr(oi), order, image_tile]
    rc = lib.clb_prepare_rows(*args, 0, C.byref(npad), None, None, None, None, None, None, None, C.byref(llc))
    L.check(rc)
    m = npad.value
    out = dict(refl=np.empty(m, np.int32), image=np.empty(m, np.int32), spot=np.empty(m, np.int32),
               oidx=np.empty(m, np.uint32), meta=np.empty((d, m), np.float32), iobs=np.empty(m, np.float32), sig=np.empty(m, np.float32))
    rc = lib.clb_prepare_rows(*args, m, C.byref(npad), ptr(out["refl"]), ptr(out["image"]), ptr(out["spot"]), ptr(out["oidx"]),
                              ptr(out["meta"]), ptr(out["iobs"]), ptr(out["sig"]), C.byref(llc))
    L.check(rc)
    out["ll_const"] = llc.value
    return out


def test_mono_prep_is_a_stable_sort_by_refl_id():
    rng = np.random.default_rng(0)
    p = synth.make_mono(1000, 64, d=3, n_images=7, seed=3)
    perm = rng.permutation(1000)
    for k in ("refl_id", "image_id", "metadata", "intensities", "uncertainties"):
        p[k] = p[k][perm]
    out = _prepare(p, 64, laue=False)
    order = np.argsort(p["refl_id"], kind="stable")
    n = 1000
    assert len(out["refl"]) == 1024
    assert np.array_equal(out["refl"][:n], p["refl_id"][order])
    assert np.array_equal(out["oidx"][:n], order)
    assert np.array_equal(out["image"][:n], p["image_id"][order])
    assert np.array_equal(out["meta"][:, :n], p["metadata"][order].T)
    assert np.array_equal(out["iobs"][:n], p["intensities"][order])
    assert np.all(out["refl"][n:] == -1) and np.all(out["sig"][n:] == 1.0) and np.all(out["meta"][:, n:] == 0)
    assert out["ll_const"] == 0.0


def test_laue_prep_groups_harmonics_inside_warp_chunks():
    import scipy.stats as st
    p = synth.make_laue(3000, 200, d=2, n_images=9, seed=5)
    out = _prepare(p, 200, laue=True)
    refl, spot = out["refl"], out["spot"]
    live = refl >= 0
    assert live.sum() == 3000
    # every original row appears exactly once, with its own ids
    assert np.array_equal(np.sort(out["oidx"][live]), np.arange(3000))
    assert np.array_equal(refl[live], p["refl_id"][out["oidx"][live]])
    assert np.array_equal(spot[live], p["harmonic_id"][out["oidx"][live]])
    assert np.array_equal(out["iobs"][live], p["intensities"][spot[live]])        # formatter.py:637-640
    # spots are contiguous, in ascending order, and never straddle a 32-row chunk
    s = spot[live]
    assert np.all(np.diff(s) >= 0)
    rows = np.nonzero(live)[0]
    first = {}; last = {}
    for r, k in zip(rows, s):
        first.setdefault(k, r); last[k] = r
    for k in first:
        assert first[k] // 32 == last[k] // 32
        assert last[k] - first[k] + 1 == np.sum(s == k)
    # constant log-density of the empty (padded) slots: logpdf(0; 1, 1) each (laue.py:23-25)
    n_empty = 3000 - p["n_spots"]
    assert np.isclose(out["ll_const"], n_empty * st.norm.logpdf(0.0, 1.0, 1.0))
    out_t = _prepare(p, 200, laue=True, likelihood=1, dof=4.0)
    assert np.isclose(out_t["ll_const"], n_empty * st.t.logpdf(0.0, 4.0, 1.0, 1.0))


def test_prep_rejects_bad_input():
    p = synth.make_mono(100, 10, d=2, n_images=3, seed=1)
    p["refl_id"] = p["refl_id"].copy(); p["refl_id"][3] = 10
    with pytest.raises(L.ClbError):
        _prepare(p, 10, laue=False)
    q = synth.make_laue(200, 20, d=2, n_images=3, seed=1)
    q["harmonic_id"] = np.zeros(200, dtype=np.int64)       # one spot with 200 harmonics
    with pytest.raises(L.ClbError):
        _prepare(q, 20, laue=True)


@pytest.mark.parametrize("laue", [False, True])
def test_image_layer_prep_keeps_one_image_per_tile(laue):
    """--image-layers: rows become image-major and every 128-row tile holds rows of a single image."""
    p = synth.make_laue(3000, 200, d=2, n_images=9, seed=5) if laue else synth.make_mono(3000, 200, d=2, n_images=9, seed=5)
    out = _prepare(p, 200, laue=laue, image_tile=128)
    live = out["refl"] >= 0
    assert live.sum() == 3000 and np.array_equal(np.sort(out["oidx"][live]), np.arange(3000))
    assert np.array_equal(out["image"][live], p["image_id"][out["oidx"][live]])
    assert len(out["refl"]) % 32 == 0
    for t in range(0, len(out["refl"]), 128):
        imgs = np.unique(out["image"][t:t + 128][live[t:t + 128]])
        assert len(imgs) <= 1
        if live[t:t + 128].any():
            assert live[t]                      # the tile's first row is real: the kernel reads its image id
    assert np.all(np.diff(out["image"][live]) >= 0)
    if laue:
        s = out["spot"][live]; rows = np.nonzero(live)[0]
        for k in np.unique(s):
            r = rows[s == k]
            assert r[0] // 32 == r[-1] // 32 and r[-1] - r[0] + 1 == len(r)
