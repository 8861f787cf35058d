"""CrystFEL .stream reader (serial crystallography), replacing `rs.read_crystfel` as the reference uses it
(careless/io/formatter.py:178-186 `format_files`, careless/io/manager.py:39-41; exercised by tests/test_cli.py:112-119).

A stream is a sequence of chunks (one detector frame each) holding zero or more indexed crystals; every crystal has
`Cell parameters a b c nm, al be ga deg` and a table `h k l I sigma(I) peak background fs/px ss/px panel` between
`Reflections measured after indexing` and `End of reflections`.  The result is an unmerged DataSet with the observed
Miller indices, I / SigI, the per-crystal BATCH number (careless's image key), peak / background and the detector
coordinates; the cell is the mean over the crystals (Angstrom); a stream carries no space group (use --spacegroups).
"""
from __future__ import annotations

import numpy as np

from .mtz import DataSet
from .symmetry import UnitCell


def read_crystfel(path, spacegroup=None):
    cells, cols = [], {k: [] for k in ("H", "K", "L", "I", "SigI", "peak", "background", "XDET", "YDET", "BATCH")}
    batch, in_refl = -1, False
    with open(path, "r", errors="replace") as f:
        for line in f:
            if in_refl:
                if line.startswith("End of reflections"):
                    in_refl = False
                    continue
                p = line.split()
                if len(p) < 9 or p[0] == "h":
                    continue
                cols["H"].append(int(p[0])); cols["K"].append(int(p[1])); cols["L"].append(int(p[2]))
                cols["I"].append(float(p[3])); cols["SigI"].append(float(p[4]))
                cols["peak"].append(float(p[5])); cols["background"].append(float(p[6]))
                cols["XDET"].append(float(p[7])); cols["YDET"].append(float(p[8]))
                cols["BATCH"].append(batch)
            elif line.startswith("Cell parameters"):
                p = line.split()
                cells.append([10.0 * float(p[2]), 10.0 * float(p[3]), 10.0 * float(p[4]), float(p[6]), float(p[7]), float(p[8])])
            elif line.startswith("--- Begin crystal"):
                batch += 1
            elif line.startswith("Reflections measured after indexing"):
                in_refl = True
    if not cols["H"]:
        raise ValueError(f"{path}: no indexed reflections found in the stream")
    out = {k: np.asarray(cols[k], dtype=np.int32) for k in ("H", "K", "L", "BATCH")}
    out.update({k: np.asarray(cols[k], dtype=np.float32) for k in ("I", "SigI", "peak", "background", "XDET", "YDET")})
    types = {"H": "H", "K": "H", "L": "H", "I": "J", "SigI": "Q", "BATCH": "B", "peak": "R", "background": "R", "XDET": "R", "YDET": "R"}
    ordered = {k: out[k] for k in ("H", "K", "L", "I", "SigI", "BATCH", "peak", "background", "XDET", "YDET")}
    cell = UnitCell(*np.mean(np.asarray(cells), axis=0)) if cells else None
    return DataSet(ordered, types, cell, spacegroup, merged=False)
