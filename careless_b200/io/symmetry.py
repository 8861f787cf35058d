"""Symmetry algebra and unit cells for the input formatter -- a from-scratch replacement for the gemmi /
reciprocalspaceship calls the reference makes on its way to `refl_id`.

What the reference uses and where:
  * `ds.hkl_to_asu(anomalous=...)`, `ds.remove_absences()`, `ds.compute_dHKL()`   careless/io/formatter.py:296-306
  * `rs.utils.generate_reciprocal_asu(cell, sg, dmin, anomalous)`                careless/io/asu.py:23-28
  * `.compute_multiplicity().label_centrics().compute_dHKL()`                    careless/io/asu.py:30-39

gemmi and reciprocalspaceship are third-party dependencies that are absent from this image, so their published
conventions are restated here:
  * Miller indices transform as row vectors, h' = h R;  M/ISYM = 2*k+1 when h_asu = h R_k and 2*k+2 when
    h_asu = -h R_k, ops tried in file order, "+" before "-" (CCP4 convention).
  * reciprocal-space ASU wedges per Laue class: the CCP4 `pgdefine` choices (table `_ASU_CONDITIONS`).
  * epsilon counts the operations (centering translations included) that fix h; a reflection is centric when
    some operation maps h to -h; it is systematically absent when an operation fixes h with a non-integral h.t.
Pinned against stored evidence: the fixtures' own ASU-mapped H,K,L + M/ISYM columns (tests/golden/).
Integer work, exact.
"""
from __future__ import annotations

import re
from fractions import Fraction

import numpy as np

DEN = 24  # translations are stored in 24ths, like gemmi's Op::DEN


# ----------------------------------------------------------------------------------------
# operations
# ----------------------------------------------------------------------------------------
class Op:
    """x' = R x + t/24.  `rot` is a 3x3 int array, `tran` three ints in 24ths."""
    __slots__ = ("rot", "tran")

    def __init__(self, rot, tran=(0, 0, 0)):
        self.rot = np.asarray(rot, dtype=np.int64).reshape(3, 3)
        self.tran = np.asarray(tran, dtype=np.int64).reshape(3) % DEN

    def __mul__(self, other):
        return Op(self.rot @ other.rot, self.rot @ other.tran + self.tran)

    def key(self):
        return tuple(self.rot.reshape(-1).tolist()) + tuple(self.tran.tolist())

    def inverse(self):
        ri = np.rint(np.linalg.inv(self.rot)).astype(np.int64)
        return Op(ri, -(ri @ self.tran))

    def apply_to_hkl(self, hkl):
        """Row-vector convention: h' = h R."""
        return np.asarray(hkl, dtype=np.int64) @ self.rot

    def triplet(self):
        out = []
        for i in range(3):
            s = ""
            for j, ax in enumerate("xyz"):
                c = int(self.rot[i, j])
                if c:
                    s += ("+" if c > 0 else "-") + ("" if abs(c) == 1 else str(abs(c))) + ax
            t = Fraction(int(self.tran[i]), DEN)
            if t:
                s += f"+{t.numerator}/{t.denominator}"
            out.append(s.lstrip("+") or "0")
        return ",".join(out)

    def __repr__(self):
        return f"Op({self.triplet()})"


def parse_triplet(text):
    """'X-Y,X,Z+1/2' -> Op (the form MTZ SYMM records and the International Tables use)."""
    parts = text.replace(" ", "").lower().split(",")
    if len(parts) != 3:
        raise ValueError(f"not a symmetry triplet: {text!r}")
    rot = np.zeros((3, 3), dtype=np.int64)
    tran = np.zeros(3, dtype=np.int64)
    for i, p in enumerate(parts):
        for sign, num, den, ax in re.findall(r"([+-]?)(\d*)(?:/(\d+))?([xyz]?)", p):
            if not (num or ax):
                continue
            sgn = -1 if sign == "-" else 1
            if ax:
                rot[i, "xyz".index(ax)] += sgn * (int(num) if num else 1)
            else:
                f = Fraction(int(num), int(den) if den else 1) * DEN
                if f.denominator != 1:
                    raise ValueError(f"translation not a multiple of 1/{DEN} in {text!r}")
                tran[i] += sgn * int(f)
    return Op(rot, tran)


# ----------------------------------------------------------------------------------------
# space groups
# ----------------------------------------------------------------------------------------
_LATTICE = {
    "P": [(0, 0, 0)],
    "A": [(0, 0, 0), (0, 12, 12)],
    "B": [(0, 0, 0), (12, 0, 12)],
    "C": [(0, 0, 0), (12, 12, 0)],
    "I": [(0, 0, 0), (12, 12, 12)],
    "R": [(0, 0, 0), (16, 8, 8), (8, 16, 16)],
    "F": [(0, 0, 0), (0, 12, 12), (12, 0, 12), (12, 12, 0)],
}

# proper rotations about z; x and y follow by cyclic permutation of the axes
_ROT_Z = {
    1: [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
    2: [[-1, 0, 0], [0, -1, 0], [0, 0, 1]],
    3: [[0, -1, 0], [1, -1, 0], [0, 0, 1]],
    4: [[0, -1, 0], [1, 0, 0], [0, 0, 1]],
    6: [[1, -1, 0], [1, 0, 0], [0, 0, 1]],
}
_TWO_DIAG = {"'": [[0, -1, 0], [-1, 0, 0], [0, 0, -1]], '"': [[0, 1, 0], [1, 0, 0], [0, 0, -1]]}   # 2-folds along a-b, a+b
_THREE_STAR = [[0, 0, 1], [1, 0, 0], [0, 1, 0]]
_HALL_TRANS = {"a": (12, 0, 0), "b": (0, 12, 0), "c": (0, 0, 12), "n": (12, 12, 12),
               "u": (6, 0, 0), "v": (0, 6, 0), "w": (0, 0, 6), "d": (6, 6, 6)}


def _axis_rotation(n, axis):
    rz = np.array(_ROT_Z[n], dtype=np.int64)
    if axis == "z":
        return rz
    perm = {"x": [1, 2, 0], "y": [2, 0, 1]}[axis]   # new index of old (x, y, z)
    p = np.zeros((3, 3), dtype=np.int64)
    for old, new in enumerate(perm):
        p[new, old] = 1
    return p @ rz @ p.T


def ops_from_hall(symbol):
    """Generators of a Hall symbol (Hall 1981) closed into a group; origin-shift vectors are ignored
    (they change neither the ASU mapping, epsilon, centricity nor the systematic absences)."""
    sym = re.sub(r"\(.*?\)", "", symbol).split()
    lat = sym[0]
    centro = lat.startswith("-")
    lat = lat.lstrip("-")
    gens = []
    prev_n, prev_axis = None, None
    for pos, tok in enumerate(sym[1:]):
        m = re.match(r"(-?)([12346])([1-5]?)([xyz'\"*]?)([abcnuvwd]*)$", tok)
        if not m:
            raise ValueError(f"cannot parse Hall symbol element {tok!r} in {symbol!r}")
        improper, n, screw, axis, trans = m.group(1) == "-", int(m.group(2)), m.group(3), m.group(4), m.group(5)
        if not axis:
            if pos == 0:
                axis = "z"
            elif pos == 1:
                axis = ("x" if prev_n in (2, 4) else "'") if n == 2 else "z"
            else:
                axis = "*" if n == 3 else "z"
        if axis in ("'", '"'):
            rot = np.array(_TWO_DIAG[axis], dtype=np.int64)
            direction = None
        elif axis == "*":
            rot = np.array(_THREE_STAR, dtype=np.int64)
            direction = None
        else:
            rot = _axis_rotation(n, axis)
            direction = "xyz".index(axis)
        t = np.zeros(3, dtype=np.int64)
        for ch in trans:
            t += np.array(_HALL_TRANS[ch])
        if screw:
            if direction is None:
                raise ValueError("screw component on a diagonal axis")
            t[direction] += DEN * int(screw) // n
        if improper:
            rot = -rot
        gens.append(Op(rot, t))
        prev_n, prev_axis = n, axis
    if centro:
        gens.append(Op(-np.eye(3, dtype=np.int64)))
    ops = _close([Op(np.eye(3, dtype=np.int64))] + gens)
    # screw generators of centred groups close onto lattice translations: keep one operation per rotation part
    cen = {tuple(c) for c in _LATTICE[lat]}
    reps = {}
    for o in ops:
        reps.setdefault(tuple(o.rot.reshape(-1).tolist()), o)
    ident = tuple(np.eye(3, dtype=np.int64).reshape(-1).tolist())
    for o in ops:
        if tuple(o.rot.reshape(-1).tolist()) == ident and tuple(o.tran.tolist()) not in cen:
            raise ValueError(f"Hall symbol {symbol!r} generates a translation outside its lattice")
    return list(reps.values()), sorted(cen)


def _close(gens):
    ops = {}
    order = []
    for g in gens:
        if g.key() not in ops:
            ops[g.key()] = g
            order.append(g)
    grew = True
    while grew:
        grew = False
        for a in list(order):
            for b in list(order):
                c = a * b
                if c.key() not in ops:
                    ops[c.key()] = c
                    order.append(c)
                    grew = True
        if len(order) > 192:
            raise ValueError("group closure exceeded 192 operations: inconsistent generators")
    return order


# Hermann-Mauguin name -> (number, Hall symbol) for the 65 Sohncke groups (the ones a protein crystal can have)
_SOHNCKE = {
    "P 1": (1, "P 1"), "P 1 2 1": (3, "P 2y"), "P 1 21 1": (4, "P 2yb"), "C 1 2 1": (5, "C 2y"),
    "P 2 2 2": (16, "P 2 2"), "P 2 2 21": (17, "P 2c 2"), "P 21 21 2": (18, "P 2 2ab"), "P 21 21 21": (19, "P 2ac 2ab"),
    "C 2 2 21": (20, "C 2c 2"), "C 2 2 2": (21, "C 2 2"), "F 2 2 2": (22, "F 2 2"), "I 2 2 2": (23, "I 2 2"),
    "I 21 21 21": (24, "I 2b 2c"),
    "P 4": (75, "P 4"), "P 41": (76, "P 4w"), "P 42": (77, "P 4c"), "P 43": (78, "P 4cw"), "I 4": (79, "I 4"), "I 41": (80, "I 4bw"),
    "P 4 2 2": (89, "P 4 2"), "P 4 21 2": (90, "P 4ab 2ab"), "P 41 2 2": (91, "P 4w 2c"), "P 41 21 2": (92, "P 4abw 2nw"),
    "P 42 2 2": (93, "P 4c 2"), "P 42 21 2": (94, "P 4n 2n"), "P 43 2 2": (95, "P 4cw 2c"), "P 43 21 2": (96, "P 4nw 2abw"),
    "I 4 2 2": (97, "I 4 2"), "I 41 2 2": (98, "I 4bw 2bw"),
    "P 3": (143, "P 3"), "P 31": (144, "P 31"), "P 32": (145, "P 32"), "R 3": (146, "R 3"),
    "P 3 1 2": (149, "P 3 2"), "P 3 2 1": (150, 'P 3 2"'), "P 31 1 2": (151, "P 31 2c"), "P 31 2 1": (152, 'P 31 2"'),
    "P 32 1 2": (153, "P 32 2c"), "P 32 2 1": (154, 'P 32 2"'), "R 3 2": (155, 'R 3 2"'),
    "P 6": (168, "P 6"), "P 61": (169, "P 61"), "P 65": (170, "P 65"), "P 62": (171, "P 62"), "P 64": (172, "P 64"), "P 63": (173, "P 6c"),
    "P 6 2 2": (177, "P 6 2"), "P 61 2 2": (178, "P 61 2"), "P 65 2 2": (179, "P 65 2"), "P 62 2 2": (180, "P 62 2c"),
    "P 64 2 2": (181, "P 64 2c"), "P 63 2 2": (182, "P 6c 2c"),
    "P 2 3": (195, "P 2 2 3"), "F 2 3": (196, "F 2 2 3"), "I 2 3": (197, "I 2 2 3"), "P 21 3": (198, "P 2ac 2ab 3"), "I 21 3": (199, "I 2b 2c 3"),
    "P 4 3 2": (207, "P 4 2 3"), "P 42 3 2": (208, "P 4n 2 3"), "F 4 3 2": (209, "F 4 2 3"), "F 41 3 2": (210, "F 4d 2 3"),
    "I 4 3 2": (211, "I 4 2 3"), "P 43 3 2": (212, "P 4acd 2ab 3"), "P 41 3 2": (213, "P 4bd 2ab 3"), "I 41 3 2": (214, "I 4bd 2c 3"),
}
_SHORT = {k.replace(" ", ""): k for k in _SOHNCKE}
_SHORT.update({"P2": "P 1 2 1", "P21": "P 1 21 1", "C2": "C 1 2 1", "P 2": "P 1 2 1", "P 21": "P 1 21 1", "C 2": "C 1 2 1",
               "P121": "P 1 2 1", "P1211": "P 1 21 1", "C121": "C 1 2 1", "R3:H": "R 3", "R32:H": "R 3 2", "H3": "R 3", "H32": "R 3 2"})

# Laue-class wedge of reciprocal space, CCP4 convention (h, k, l integer arrays -> bool array)
_ASU_CONDITIONS = {
    "-1": lambda h, k, l: (l > 0) | ((l == 0) & ((h > 0) | ((h == 0) & (k >= 0)))),
    "2/m": lambda h, k, l: (k >= 0) & ((l > 0) | ((l == 0) & (h >= 0))),
    "2/m(c)": lambda h, k, l: (l >= 0) & ((h > 0) | ((h == 0) & (k >= 0))),
    "mmm": lambda h, k, l: (h >= 0) & (k >= 0) & (l >= 0),
    "4/m": lambda h, k, l: (l >= 0) & (((h >= 0) & (k > 0)) | ((h == 0) & (k == 0))),
    "4/mmm": lambda h, k, l: (h >= k) & (k >= 0) & (l >= 0),
    "-3": lambda h, k, l: ((h >= 0) & (k > 0)) | ((h == 0) & (k == 0) & (l >= 0)),
    "-31m": lambda h, k, l: (h >= k) & (k >= 0) & ((k > 0) | (l >= 0)),
    "-3m1": lambda h, k, l: (h >= k) & (k >= 0) & ((h > k) | (l >= 0)),
    "6/m": lambda h, k, l: (l >= 0) & (((h >= 0) & (k > 0)) | ((h == 0) & (k == 0))),
    "6/mmm": lambda h, k, l: (h >= k) & (k >= 0) & (l >= 0),
    "m-3": lambda h, k, l: (h >= 0) & (((l >= h) & (k > h)) | ((l == h) & (k == h))),
    "m-3m": lambda h, k, l: (k >= l) & (l >= h) & (h >= 0),
}


class SpaceGroup:
    """A space group in its reference setting: `sym_ops` (coset representatives, file order) x `cen_ops`."""

    def __init__(self, sym_ops, cen_ops=((0, 0, 0),), name=None, number=None):
        self.sym_ops = list(sym_ops)
        self.cen_ops = [tuple(int(x) % DEN for x in c) for c in cen_ops]
        self.name, self.number = name, number
        self._rots = np.stack([o.rot for o in self.sym_ops])          # (n, 3, 3)
        self._trans = np.stack([o.tran for o in self.sym_ops])        # (n, 3)
        self.laue = self._laue_class()
        self._in_asu = _ASU_CONDITIONS[self.laue]

    # -- constructors ----------------------------------------------------------------
    @classmethod
    def from_name(cls, name):
        key = " ".join(str(name).split())
        if key.isdigit():      # ITA number (gemmi.SpaceGroup("19")), reference setting
            key = next((k for k, (num, _) in _SOHNCKE.items() if num == int(key)), None)
        else:
            key = key if key in _SOHNCKE else _SHORT.get(key.replace(" ", ""), _SHORT.get(key))
        if key is None:
            raise ValueError(f"unknown space group {name!r} (the 65 Sohncke groups are tabulated, reference settings)")
        number, hall = _SOHNCKE[key]
        ops, cen = ops_from_hall(hall)
        return cls(ops, cen, name=key, number=number)

    @classmethod
    def from_triplets(cls, triplets, name=None, number=None):
        """All operations of the group as the file lists them (MTZ SYMM records); the centering translations are
        split off so that `sym_ops` keeps the file's order of first appearance (what M/ISYM indexes)."""
        ops = [parse_triplet(t) for t in triplets]
        first = {}
        for o in ops:
            first.setdefault(tuple(o.rot.reshape(-1).tolist()), o)
        ident = tuple(np.eye(3, dtype=np.int64).reshape(-1).tolist())
        cen = sorted({tuple(o.tran.tolist()) for o in ops if tuple(o.rot.reshape(-1).tolist()) == ident})
        if len(first) * len(cen) != len(ops):
            raise ValueError("symmetry operations do not factor into rotations x centering translations")
        return cls(list(first.values()), cen, name=name, number=number)

    # -- classification ----------------------------------------------------------------
    def _laue_class(self):
        keys = set()
        for r in self._rots:
            keys.add(tuple(r.reshape(-1).tolist()))
            keys.add(tuple((-r).reshape(-1).tolist()))
        n = len(keys)
        proper = [np.array(k).reshape(3, 3) for k in keys if round(np.linalg.det(np.array(k).reshape(3, 3))) == 1]
        traces = {int(np.trace(p)) for p in proper}
        has6, has4, has3 = 2 in traces, 1 in traces, 0 in traces

        def contains(m):
            return tuple(np.array(m).reshape(-1).tolist()) in keys

        def need_c_axis(order):
            if not contains(_ROT_Z[order]):
                raise ValueError(f"{self.name}: principal axis is not c; only reference settings are supported")

        if n == 2:
            return "-1"
        if n == 4:
            if contains(_axis_rotation(2, "y")):
                return "2/m"
            if contains(_ROT_Z[2]):
                return "2/m(c)"
            raise ValueError("monoclinic unique axis a is not supported")
        if n == 8:
            if has4:
                need_c_axis(4)
                return "4/m"
            return "mmm"
        if n == 16:
            need_c_axis(4)
            return "4/mmm"
        if n == 6:
            need_c_axis(3)
            return "-3"
        if n == 12:
            if has6:
                need_c_axis(6)
                return "6/m"
            need_c_axis(3)
            return "-3m1" if contains(_TWO_DIAG['"']) else "-31m"
        if n == 24:
            if has6:
                need_c_axis(6)
                return "6/mmm"
            return "m-3"
        if n == 48:
            return "m-3m"
        raise ValueError(f"cannot classify a point group of Laue order {n}")

    @property
    def order(self):
        return len(self.sym_ops) * len(self.cen_ops)

    def all_ops(self):
        return [Op(o.rot, o.tran + np.array(c)) for c in self.cen_ops for o in self.sym_ops]

    def xhm(self):
        return self.name

    def __repr__(self):
        return f"SpaceGroup({self.name!r}, {len(self.sym_ops)}x{len(self.cen_ops)} ops, Laue {self.laue})"

    # -- per-reflection properties (vectorised over an (n, 3) integer array) ---------------
    def _images(self, hkl):
        hkl = np.asarray(hkl, dtype=np.int64).reshape(-1, 3)
        return np.einsum("ni,kij->knj", hkl, self._rots)               # (ops, n, 3), h R_k

    def in_asu(self, hkl):
        hkl = np.asarray(hkl, dtype=np.int64).reshape(-1, 3)
        return self._in_asu(hkl[:, 0], hkl[:, 1], hkl[:, 2])

    def hkl_to_asu(self, hkl):
        """(asu_hkl, isym): first operation (file order, '+' before '-') that lands in the ASU wedge."""
        hkl = np.asarray(hkl, dtype=np.int64).reshape(-1, 3)
        out = np.zeros_like(hkl)
        isym = np.zeros(len(hkl), dtype=np.int32)
        todo = np.ones(len(hkl), dtype=bool)
        for k, r in enumerate(self._rots):
            if not todo.any():
                break
            for sign, code in ((1, 2 * k + 1), (-1, 2 * k + 2)):
                idx = np.nonzero(todo)[0]
                if idx.size == 0:
                    break
                cand = sign * (hkl[idx] @ r)
                ok = self._in_asu(cand[:, 0], cand[:, 1], cand[:, 2])
                sel = idx[ok]
                out[sel] = cand[ok]
                isym[sel] = code
                todo[sel] = False
        if todo.any():
            raise ValueError(f"{int(todo.sum())} reflections have no image in the ASU wedge (inconsistent operations?)")
        return out, isym

    def hkl_to_observed(self, asu_hkl, isym):
        """Inverse of hkl_to_asu for a stored M/ISYM column (the part below 256)."""
        asu_hkl = np.asarray(asu_hkl, dtype=np.int64).reshape(-1, 3)
        isym = np.asarray(isym, dtype=np.int64) % 256
        out = np.zeros_like(asu_hkl)
        for code in np.unique(isym):
            k = (int(code) - 1) // 2
            if code < 1 or k >= len(self.sym_ops):
                raise ValueError(f"M/ISYM value {int(code)} does not index one of the {len(self.sym_ops)} operations")
            rinv = np.rint(np.linalg.inv(self._rots[k])).astype(np.int64)
            sel = isym == code
            obs = asu_hkl[sel] @ rinv
            out[sel] = obs if code % 2 == 1 else -obs
        return out

    def is_centric(self, hkl):
        hkl = np.asarray(hkl, dtype=np.int64).reshape(-1, 3)
        return np.any(np.all(self._images(hkl) == -hkl[None], axis=-1), axis=0)

    def epsilon(self, hkl, include_centering=True):
        hkl = np.asarray(hkl, dtype=np.int64).reshape(-1, 3)
        eps = np.sum(np.all(self._images(hkl) == hkl[None], axis=-1), axis=0)
        return eps * (len(self.cen_ops) if include_centering else 1)

    def is_absent(self, hkl):
        hkl = np.asarray(hkl, dtype=np.int64).reshape(-1, 3)
        absent = np.zeros(len(hkl), dtype=bool)
        for c in self.cen_ops:
            absent |= (hkl @ np.array(c, dtype=np.int64)) % DEN != 0
        fixed = np.all(self._images(hkl) == hkl[None], axis=-1)        # (ops, n)
        shift = np.einsum("ni,ki->kn", hkl, self._trans) % DEN
        absent |= np.any(fixed & (shift != 0), axis=0)
        return absent


# ----------------------------------------------------------------------------------------
# unit cell
# ----------------------------------------------------------------------------------------
class UnitCell:
    def __init__(self, a, b, c, alpha, beta, gamma):
        self.parameters = tuple(float(x) for x in (a, b, c, alpha, beta, gamma))
        a, b, c, al, be, ga = self.parameters
        ca, cb, cg = (np.cos(np.deg2rad(x)) for x in (al, be, ga))
        sa, sb, sg = (np.sin(np.deg2rad(x)) for x in (al, be, ga))
        for ang, name in ((al, "ca"), (be, "cb"), (ga, "cg")):
            if ang == 90.0:
                if name == "ca": ca, sa = 0.0, 1.0
                if name == "cb": cb, sb = 0.0, 1.0
                if name == "cg": cg, sg = 0.0, 1.0
        self.volume = a * b * c * np.sqrt(max(0.0, 1 - ca * ca - cb * cb - cg * cg + 2 * ca * cb * cg))
        self.ar, self.br, self.cr = b * c * sa / self.volume, a * c * sb / self.volume, a * b * sg / self.volume
        self.cos_alphar = (cb * cg - ca) / (sb * sg)
        self.cos_betar = (ca * cg - cb) / (sa * sg)
        self.cos_gammar = (ca * cb - cg) / (sa * sb)

    @property
    def a(self): return self.parameters[0]
    @property
    def b(self): return self.parameters[1]
    @property
    def c(self): return self.parameters[2]

    def calculate_1_d2_array(self, hkl):
        hkl = np.asarray(hkl, dtype=np.float64).reshape(-1, 3)
        h, k, l = hkl[:, 0] * self.ar, hkl[:, 1] * self.br, hkl[:, 2] * self.cr
        return h * h + k * k + l * l + 2 * (h * k * self.cos_gammar + h * l * self.cos_betar + k * l * self.cos_alphar)

    def calculate_d_array(self, hkl):
        with np.errstate(divide="ignore"):
            return 1.0 / np.sqrt(self.calculate_1_d2_array(hkl))

    def get_hkl_limits(self, dmin):
        return tuple(int(x / dmin) for x in self.parameters[:3])

    def is_compatible_with_spacegroup(self, sg, eps=1e-3):
        """Metric tensor invariant under every rotation of the group (within a relative eps)."""
        a, b, c, al, be, ga = self.parameters
        ca, cb, cg = (np.cos(np.deg2rad(x)) for x in (al, be, ga))
        g = np.array([[a * a, a * b * cg, a * c * cb], [a * b * cg, b * b, b * c * ca], [a * c * cb, b * c * ca, c * c]])
        for o in sg.sym_ops:
            r = o.rot.astype(np.float64)
            if np.max(np.abs(r.T @ g @ r - g)) > eps * np.max(np.abs(g)):
                return False
        return True

    def __repr__(self):
        return "UnitCell(%g, %g, %g, %g, %g, %g)" % self.parameters


# ----------------------------------------------------------------------------------------
# reciprocal-space enumeration (rs.utils.generate_reciprocal_asu as used by careless/io/asu.py:23-28)
# ----------------------------------------------------------------------------------------
def generate_reciprocal_cell(cell, dmin):
    """Every hkl != 0 with d >= dmin.  Order: k slowest, then h, then l (numpy's default 'xy' meshgrid of
    (h, k, l) ranges flattened in C order -- the order rs.utils.generate_reciprocal_cell produces;
    the numbering itself is unpinned by stored vectors, see oracle/__init__.py)."""
    hmax, kmax, lmax = cell.get_hkl_limits(dmin)
    hs = np.arange(-hmax, hmax + 2, dtype=np.int32)
    ks = np.arange(-kmax, kmax + 2, dtype=np.int32)
    ls = np.arange(-lmax, lmax + 2, dtype=np.int32)
    hh, kk, ll = np.meshgrid(hs, ks, ls)
    hkl = np.stack([hh, kk, ll]).reshape(3, -1).T
    hkl = hkl[np.any(hkl != 0, axis=1)]
    d = cell.calculate_d_array(hkl).astype(np.float32)
    return hkl[d >= np.float32(dmin)]


def generate_reciprocal_asu(cell, spacegroup, dmin, anomalous=False):
    hkl = generate_reciprocal_cell(cell, dmin)
    hkl = hkl[~spacegroup.is_absent(hkl)]
    hasu = hkl[spacegroup.in_asu(hkl)]
    if anomalous:
        minus = -hasu[~spacegroup.is_centric(hasu)]
        hasu = np.unique(np.concatenate([hasu, minus]), axis=0)
    return hasu
