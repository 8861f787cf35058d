#!/usr/bin/env python
"""Regenerates tests/golden/fixture_mtz.npz from the reference's own test fixtures (run in the build container only;
/root/reference does not exist on the GPU box).

    python tests/golden/make_fixture_vectors.py

Stored per file (pyp_off, pyp_2ms: P6_3; pyp_2ms_P3: P3): every column exactly as it sits in the MTZ -- H,K,L there
are ASU-mapped and M/ISYM records which operation did it, both written by reciprocalspaceship/gemmi -- plus the cell,
the space-group name and the SYMM triplets.  tests/test_io_formatter.py un-maps H,K,L with M/ISYM and requires our
`hkl_to_asu` to reproduce both columns bit for bit: the only stored evidence of the reference's ASU convention.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from careless_b200.io.mtz import read_mtz  # noqa: E402

SRC = "/root/reference/tests/data"
out = {}
for name in ("pyp_off", "pyp_2ms", "pyp_2ms_P3"):
    ds = read_mtz(os.path.join(SRC, name + ".mtz"), to_observed=False)
    keys = ds.keys()
    out[f"{name}/keys"] = np.array(keys)
    out[f"{name}/types"] = np.array([ds.dtypes[k] for k in keys])
    out[f"{name}/data"] = np.stack([np.asarray(ds[k], dtype=np.float32) for k in keys], axis=1)
    out[f"{name}/cell"] = np.array(ds.cell.parameters)
    out[f"{name}/spacegroup"] = np.array(ds.spacegroup.name)
    out[f"{name}/symm"] = np.array([o.triplet() for o in ds.spacegroup.all_ops()])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fixture_mtz.npz"), **out)
print("wrote fixture_mtz.npz:", {k: v.shape for k, v in out.items() if k.endswith("/data")})
