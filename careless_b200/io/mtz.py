"""Minimal MTZ reader / writer and the reflection table the formatter works on.

Replaces `rs.read_mtz` / `rs.DataSet` / `DataSet.write_mtz` as the reference uses them
(careless/io/formatter.py:178-186 `format_files`; careless/io/manager.py:164-250 results) -- reciprocalspaceship and
gemmi are absent from this image.  MTZ (CCP4) layout: bytes 0-3 'MTZ ', word 2 = 1-based word offset of the
header, word 3 = machine stamp; reflection records are `ncol` float32 words each starting at word 21; the
header is a list of 80-character records ending with 'END'.  Unmerged files carry ASU-mapped H,K,L plus an
M/ISYM column; like rs.read_mtz this reader returns the OBSERVED indices and drops M/ISYM.
"""
from __future__ import annotations

import struct

import numpy as np

from .symmetry import SpaceGroup, UnitCell


class DataSet:
    """Column store: name -> 1-D numpy array (+ MTZ column type letters, cell, space group)."""

    def __init__(self, columns=None, dtypes=None, cell=None, spacegroup=None, merged=False):
        self.columns = dict(columns or {})
        self.dtypes = dict(dtypes or {})
        self.cell, self.spacegroup, self.merged = cell, spacegroup, merged

    # -- mapping interface -----------------------------------------------------------
    def __contains__(self, k): return k in self.columns
    def __iter__(self): return iter(self.columns)
    def __len__(self): return len(next(iter(self.columns.values()))) if self.columns else 0
    def keys(self): return list(self.columns.keys())
    def __getitem__(self, k): return self.columns[k]

    def __setitem__(self, k, v):
        v = np.asarray(v)
        if v.ndim == 0:
            v = np.full(len(self), v)
        self.columns[k] = v
        self.dtypes.setdefault(k, "I" if np.issubdtype(v.dtype, np.integer) else "R")

    def copy(self):
        return DataSet({k: v.copy() for k, v in self.columns.items()}, self.dtypes, self.cell, self.spacegroup, self.merged)

    def take(self, idx):
        """Rows `idx` (integer index or boolean mask) of every column."""
        return DataSet({k: v[idx] for k, v in self.columns.items()}, self.dtypes, self.cell, self.spacegroup, self.merged)

    def first_key_of_dtype(self, letter):
        for k in self.columns:
            if self.dtypes.get(k) == letter:
                return k
        return None

    # -- crystallography -----------------------------------------------------------
    def get_hkls(self):
        return np.stack([self.columns[k] for k in ("H", "K", "L")], axis=1).astype(np.int32)

    def set_hkls(self, hkl):
        for i, k in enumerate(("H", "K", "L")):
            self.columns[k] = np.asarray(hkl[:, i], dtype=np.int32)

    def compute_dHKL(self):
        self.columns["dHKL"] = self.cell.calculate_d_array(self.get_hkls()).astype(np.float32)
        self.dtypes["dHKL"] = "R"
        return self

    def remove_absences(self):
        return self.take(~self.spacegroup.is_absent(self.get_hkls()))

    def hkl_to_asu(self, anomalous=False):
        """H,K,L -> ASU; adds M/ISYM.  With anomalous=True acentric Friedel-minus reflections keep the sign."""
        asu, isym = self.spacegroup.hkl_to_asu(self.get_hkls())
        if anomalous:
            minus = (isym % 2 == 0) & ~self.spacegroup.is_centric(asu)
            asu[minus] *= -1
        self.set_hkls(asu)
        self.columns["M/ISYM"] = isym.astype(np.int32)
        self.dtypes["M/ISYM"] = "Y"
        return self


def concat(datasets):
    first = datasets[0]
    keys = [k for k in first.columns if all(k in d for d in datasets)]
    cols = {k: np.concatenate([d[k] for d in datasets]) for k in keys}
    return DataSet(cols, first.dtypes, first.cell, first.spacegroup, first.merged)


# ----------------------------------------------------------------------------------------
def read_mtz(path, to_observed=True):
    raw = open(path, "rb").read()
    if raw[:4] != b"MTZ ":
        raise ValueError(f"{path}: not an MTZ file")
    stamp = raw[8]
    endian = "<" if (stamp >> 4) == 4 else ">"          # 0x44 'DA' = little-endian IEEE, 0x11 = big-endian
    (hdr,) = struct.unpack(endian + "i", raw[4:8])
    if hdr == -1:
        (hdr,) = struct.unpack(endian + "q", raw[12:20])
    text = raw[(hdr - 1) * 4:]
    names, types, symm = [], [], []
    ncol = nref = None
    cell = None
    sgname, sgnum = None, None
    for i in range(0, len(text), 80):
        rec = text[i:i + 80].decode("ascii", "replace")
        key = rec[:8].split()[0].upper() if rec.strip() else ""
        if key == "END":
            break
        if key == "NCOL":
            f = rec.split()
            ncol, nref = int(f[1]), int(f[2])
        elif key == "CELL":
            cell = UnitCell(*[float(x) for x in rec.split()[1:7]])
        elif key == "SYMINF":
            q = rec.split("'")
            if len(q) >= 2:
                sgname = q[1]
            f = rec.split()
            try:
                sgnum = int(f[4])
            except (IndexError, ValueError):
                sgnum = None
        elif key == "SYMM":
            symm.append(rec[4:].strip())
        elif key == "COLUMN":
            # name may not contain blanks; type letter follows
            f = rec.split()
            names.append(f[1])
            types.append(f[2])
    if ncol is None or len(names) != ncol:
        raise ValueError(f"{path}: header lists {len(names)} COLUMN records for NCOL {ncol}")
    data = np.frombuffer(raw, dtype=endian + "f4", count=ncol * nref, offset=80).reshape(nref, ncol)
    sg = SpaceGroup.from_triplets(symm, name=sgname, number=sgnum) if symm else None
    cols, dts = {}, {}
    for j, (nm, ty) in enumerate(zip(names, types)):
        col = data[:, j]
        if ty in ("H", "B", "Y", "I"):
            col = np.where(np.isnan(col), 0, col).astype(np.int32)
        else:
            col = col.astype(np.float32)
        cols[nm], dts[nm] = col, ty
    ds = DataSet(cols, dts, cell, sg, merged=True)
    misym = [k for k, t in dts.items() if t == "Y"]
    ds.merged = not misym and "BATCH" not in cols
    if misym and to_observed:
        if len(misym) != 1:
            raise ValueError(f"{path}: expected one M/ISYM column, found {misym}")
        ds.set_hkls(sg.hkl_to_observed(ds.get_hkls(), ds[misym[0]]))
        del ds.columns[misym[0]]
    return ds


def write_mtz(path, ds, title="careless_b200"):
    """Little-endian MTZ with one dataset; integer columns are stored as floats as the format requires."""
    names = list(ds.columns.keys())
    n = len(ds)
    data = np.empty((n, len(names)), dtype="<f4")
    for j, k in enumerate(names):
        data[:, j] = ds.columns[k].astype(np.float32)
    recs = ["VERS MTZ:V1.1", f"TITLE {title}", f"NCOL {len(names):8d} {n:12d} {0:8d}"]
    cellstr = "".join(f"{x:10.4f}" for x in ds.cell.parameters)
    recs.append(f"CELL  {cellstr}")
    recs.append("SORT    0   0   0   0   0")
    sg = ds.spacegroup
    ops = sg.all_ops()
    lat = (sg.name or "P")[0]
    recs.append(f"SYMINF {len(ops):3d} {len(sg.sym_ops):2d} {lat} {sg.number or 0:5d} {(chr(39) + (sg.name or '') + chr(39)):>22s} PG{sg.laue}")
    for o in ops:
        recs.append("SYMM " + o.triplet().upper())
    d = ds.cell.calculate_d_array(ds.get_hkls()) if n else np.array([1.0])
    recs.append(f"RESO {np.min(1 / d ** 2):.12f}       {np.max(1 / d ** 2):.12f}")
    recs.append("VALM NAN")
    for j, k in enumerate(names):
        col = data[:, j]
        fin = col[np.isfinite(col)]
        lo, hi = (float(fin.min()), float(fin.max())) if fin.size else (0.0, 0.0)
        recs.append(f"COLUMN {k:<30s} {ds.dtypes.get(k, 'R')} {lo:17.9f} {hi:17.9f} {0 if k in ('H', 'K', 'L') else 1:4d}")
    recs += ["NDIF        2", "PROJECT       0 HKL_base", "CRYSTAL       0 HKL_base", "DATASET       0 HKL_base",
             f"DCELL         0 {cellstr}", "DWAVEL        0    0.00000",
             "PROJECT       1 careless_b200", "CRYSTAL       1 careless_b200", "DATASET       1 careless_b200",
             f"DCELL         1 {cellstr}", "DWAVEL        1    0.00000", "END", "MTZENDOFHEADERS"]
    body = data.tobytes()
    head = b"MTZ " + struct.pack("<i", 21 + data.size) + bytes([0x44, 0x41, 0x00, 0x00]) + b"\0" * 68
    with open(path, "wb") as f:
        f.write(head)
        f.write(body)
        for r in recs:
            f.write(r.ljust(80)[:80].encode("ascii"))
