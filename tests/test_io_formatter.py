"""CPU tests of the input formatter: symmetry algebra, ASU numbering, MTZ i/o, mono and Laue
formatters.  Models: tests/io/test_asu.py:8-81, tests/io/test_data_formatter.py:10-120, tests/utils/test_laue.py of the
reference.  The ASU mapping is pinned against the reference fixtures' own H,K,L + M/ISYM columns (tests/golden/)."""
import itertools
import os

import numpy as np
import pytest

import _util as U
from careless_b200.io import symmetry as S
from careless_b200.io.asu import ReciprocalASU, ReciprocalASUCollection
from careless_b200.io.formatter import LaueFormatter, MonoFormatter, expand_harmonics, ngroup, positional_encoding
from careless_b200.io.mtz import read_mtz, write_mtz
from careless_b200.models.base import BaseModel

FIXTURES = ("pyp_off", "pyp_2ms", "pyp_2ms_P3")
CELLS = [((10., 20., 30., 90., 80., 75.), "P 1"), ((30., 50., 80., 90., 100., 90.), "P 1 21 1"),
         ((10., 20., 30., 90., 90., 90.), "P 21 21 21"), ((89., 89., 105., 90., 90., 120.), "P 31 2 1"),
         ((30., 30., 30., 90., 90., 120.), "R 32")]          # tests/conftest.py:31-40 of the reference


@pytest.mark.parametrize("name", FIXTURES)
def test_hkl_to_asu_reproduces_fixture_columns(name):
    """Un-map the stored ASU indices with the stored M/ISYM, map them back: H,K,L and M/ISYM must match bit for bit."""
    raw = U.load_fixture(name, to_observed=False)
    obs = U.load_fixture(name)
    assert not np.array_equal(obs.get_hkls(), raw.get_hkls())          # the un-mapping did something
    asu, isym = obs.spacegroup.hkl_to_asu(obs.get_hkls())
    assert np.array_equal(asu, raw.get_hkls())
    assert np.array_equal(isym, raw["M/ISYM"] % 256)
    assert obs.spacegroup.in_asu(raw.get_hkls()).all()


def test_sohncke_tables_close_and_have_one_asu_image_per_orbit():
    point_order = lambda num: (1 if num < 3 else 2 if num < 16 else 4 if num < 89 else 8 if num < 143 else 3 if num < 149
                               else 6 if num < 177 else 12 if num < 207 else 24)
    special = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 1, 1], [2, 1, 0], [0, 0, -1], [1, 0, 1], [0, 1, 1],
               [-1, -1, 0], [1, 2, 0], [2, -1, 0], [1, 1, -1], [-2, 1, 1]]
    for name, (num, _) in S._SOHNCKE.items():
        sg = S.SpaceGroup.from_name(name)
        assert len(sg.sym_ops) == point_order(num), name
        rng = np.random.default_rng(num)
        hkl = np.concatenate([rng.integers(-6, 7, size=(200, 3)), special])
        hkl = hkl[np.any(hkl != 0, 1)]
        imgs = sg._images(hkl)
        imgs = np.concatenate([imgs, -imgs])
        for j in range(len(hkl)):
            orbit = np.unique(imgs[:, j, :], axis=0)
            assert sg.in_asu(orbit).sum() == 1, (name, hkl[j])
        asu, isym = sg.hkl_to_asu(hkl)
        assert np.array_equal(sg.hkl_to_observed(asu, isym), hkl)
        # epsilon, centricity and absences are class functions of the orbit
        assert np.array_equal(sg.epsilon(asu), sg.epsilon(hkl))
        assert np.array_equal(sg.is_centric(asu), sg.is_centric(hkl))
        assert np.array_equal(sg.is_absent(asu), sg.is_absent(hkl))


def test_known_absences_centrics_epsilon():
    p212121 = S.SpaceGroup.from_name("P 21 21 21")
    h = np.array([[1, 0, 0], [2, 0, 0], [0, 3, 0], [0, 0, 4], [1, 2, 0], [1, 2, 3]])
    assert p212121.is_absent(h).tolist() == [True, False, True, False, False, False]
    assert p212121.is_centric(h).tolist() == [True, True, True, True, True, False]
    assert p212121.epsilon(h).tolist() == [2, 2, 2, 2, 1, 1]
    p63 = S.SpaceGroup.from_name("P 63")
    h = np.array([[0, 0, 1], [0, 0, 2], [1, 0, 0], [1, 2, 3]])
    assert p63.is_absent(h).tolist() == [True, False, False, False]
    assert p63.epsilon(h).tolist() == [6, 6, 1, 1]
    assert p63.is_centric(h).tolist() == [False, False, True, False]
    c2 = S.SpaceGroup.from_name("C 1 2 1")
    h = np.array([[1, 0, 0], [1, 1, 0], [0, 2, 0], [2, 0, 1]])
    assert c2.is_absent(h).tolist() == [True, False, False, False]
    assert c2.epsilon(h).tolist() == [2, 2, 4, 2]                   # centering included, like rs.compute_multiplicity
    assert c2.is_centric(h).tolist() == [True, False, False, True]
    p41212 = S.SpaceGroup.from_name("P 41 21 2")
    h = np.array([[0, 0, 1], [0, 0, 4], [1, 0, 0], [2, 0, 0], [1, 1, 0]])
    assert p41212.is_absent(h).tolist() == [True, False, True, False, False]


def test_triplet_round_trip():
    for t in ("x,y,z", "x-y,x,z+1/2", "-y,x-y,z+1/3", "-x+1/2,-y,z+1/2", "y+1/4,x+3/4,-z+3/4"):
        assert S.parse_triplet(t).triplet() == t
    assert S.parse_triplet("X-Y, X, Z+1/2").triplet() == "x-y,x,z+1/2"


@pytest.mark.parametrize("anomalous", [True, False])
@pytest.mark.parametrize("dmin", [10., 5.])
def test_reciprocal_asu(dmin, anomalous):
    for cellp, name in CELLS:
        cell, sg = S.UnitCell(*cellp), S.SpaceGroup.from_name(name)
        rasu = ReciprocalASU(cell, sg, dmin, anomalous)
        Hall = S.generate_reciprocal_asu(cell, sg, dmin, anomalous)
        assert len(Hall) > 0
        assert np.all(rasu.centric == sg.is_centric(Hall))
        assert np.all(rasu.multiplicity == sg.epsilon(Hall))
        assert np.all(cell.calculate_d_array(Hall).astype(np.float32) >= dmin)
        assert not sg.is_absent(Hall).any()
        assert np.all(rasu.to_refl_id(Hall) == np.arange(len(Hall)))
        assert np.all(rasu.to_miller_index(np.arange(len(Hall))) == Hall)
        assert np.all(np.isfinite(rasu.dHKL))
        assert len(np.unique(Hall, axis=0)) == len(Hall)
        # completeness: every non-absent reflection of the sphere maps onto a member
        sphere = S.generate_reciprocal_cell(cell, dmin)
        sphere = sphere[~sg.is_absent(sphere)]
        asu, isym = sg.hkl_to_asu(sphere)
        if anomalous:
            minus = (isym % 2 == 0) & ~sg.is_centric(asu)
            asu[minus] *= -1
        assert np.all(np.sort(np.unique(rasu.to_refl_id(asu))) == np.arange(len(Hall)))


@pytest.mark.parametrize("anomalous", [[True, True], [True, False], [False, False]])
@pytest.mark.parametrize("dmin", [[10., 10.], [5., 10.], [5., 5.]])
def test_double_reciprocal_asu_collection(dmin, anomalous):
    for cellp, name in CELLS:
        cell, sg = S.UnitCell(*cellp), S.SpaceGroup.from_name(name)
        rasus = [ReciprocalASU(cell, sg, d, a) for d, a in zip(dmin, anomalous)]
        rac = ReciprocalASUCollection(rasus)
        per_asu = [S.generate_reciprocal_asu(cell, sg, d, a) for d, a in zip(dmin, anomalous)]
        refl_ids = []
        for asu_id, h in enumerate(per_asu):
            refl_id = rac.to_refl_id(asu_id * np.ones((len(h), 1)), h)
            a_test, h_test = rac.to_asu_id_and_miller_index(refl_id)
            assert np.all(a_test == asu_id) and np.all(h_test == h)
            refl_ids.append(refl_id)
        refl_ids = np.concatenate(refl_ids)
        assert np.all(refl_ids == np.arange(len(refl_ids)))                  # no gaps, no duplicates
        assert len(rac.hkls) == sum(len(h) for h in per_asu)
        assert np.all(rac.centric == np.concatenate([sg.is_centric(h) for h in per_asu]))
        assert np.all(rac.multiplicity == np.concatenate([sg.epsilon(h) for h in per_asu]))
        assert rasus[0] is rac[0] and rasus[1] is rac[1]
        missing = rac.to_refl_id(np.array([[0]]), np.array([[99, 99, 99]]), allow_missing=True)
        assert missing.tolist() == [-1]
        with pytest.raises(KeyError):
            rac.to_refl_id(np.array([[0]]), np.array([[99, 99, 99]]))


def test_ngroup_matches_sorted_rank():
    rng = np.random.default_rng(0)
    a, b, c = rng.integers(0, 4, 200), rng.integers(-3, 3, 200), rng.integers(0, 2, 200)
    got = ngroup(a, b, c)
    keys = sorted(set(zip(a.tolist(), b.tolist(), c.tolist())))
    want = np.array([keys.index(t) for t in zip(a.tolist(), b.tolist(), c.tolist())])
    assert np.array_equal(got, want)


METADATA_KEYS = ["dHKL", "Hobs", "image_id"]
GRID = list(itertools.product(["I", None], ["SigI", None], ["BATCH", None], [True, False], [True, False]))


@pytest.mark.parametrize("intensity_key,sigma_key,image_key,separate,anomalous", GRID)
@pytest.mark.parametrize("dmin,isigi,pe", [(0., None, None), (7., 3., ["X", "Y"])])
def test_mono_formatter(intensity_key, sigma_key, image_key, separate, anomalous, dmin, isigi, pe):
    ds = [U.load_fixture("pyp_off"), U.load_fixture("pyp_2ms")]
    f = MonoFormatter(intensity_key, sigma_key, image_key, METADATA_KEYS, separate, anomalous, dmin, isigi, pe, 3)
    inputs, rac = f(ds)
    n = inputs[0].shape[0]
    assert len(inputs) == 6 and n > 0
    for v in inputs:
        assert v.ndim == 2 and v.dtype in (np.float32, np.int64) and v.shape[0] == n
    refl_id = BaseModel.get_refl_id(inputs).reshape(-1)
    assert refl_id.min() >= 0 and refl_id.max() < len(rac.hkls)
    assert BaseModel.get_metadata(inputs).shape[1] == 3 + (0 if pe is None else 2 * 2 * 3)
    file_id = inputs[2].reshape(-1)
    assert set(np.unique(file_id)) == {0, 1}
    assert np.all(rac.asu_ids[refl_id] == (file_id if separate else 0))
    image_id = inputs[1].reshape(-1)
    assert np.array_equal(np.unique(image_id), np.arange(image_id.max() + 1))
    if isigi is not None:
        assert np.all(inputs[4] / inputs[5] >= isigi)


def test_mono_formatter_ids_are_the_asu_rows_of_the_file():
    """refl_id must point at exactly the H,K,L the reference's own writer stored in the file."""
    raw = U.load_fixture("pyp_off", to_observed=False)
    f = MonoFormatter(None, None, None, ["dHKL"], False, False)
    inputs, rac = f([U.load_fixture("pyp_off")])
    assert np.array_equal(rac.hkls[inputs[0][:, 0]], raw.get_hkls())
    d = raw.cell.calculate_d_array(raw.get_hkls())
    z = d ** -2.0
    assert np.allclose(inputs[3][:, 0], (z - z.mean()) / z.std(), atol=2e-4)
    assert np.array_equal(inputs[4][:, 0], raw["I"]) and np.array_equal(inputs[5][:, 0], raw["SigI"])
    assert np.array_equal(inputs[1][:, 0], ngroup(raw["BATCH"]))


@pytest.mark.parametrize("lam_min,lam_max", [(None, None), (1.05, 1.15)])
@pytest.mark.parametrize("separate,anomalous", [(True, True), (False, False)])
@pytest.mark.parametrize("dmin,isigi,pe", [(None, None, None), (7., 3., ["X", "Y"])])
def test_laue_formatter(lam_min, lam_max, separate, anomalous, dmin, isigi, pe):
    ds = [U.load_fixture("pyp_off"), U.load_fixture("pyp_2ms")]
    f = LaueFormatter("Wavelength", None, None, None, METADATA_KEYS, separate, anomalous, lam_min, lam_max, dmin, isigi, pe, 3)
    inputs, rac = f(ds)
    n = inputs[0].shape[0]
    assert len(inputs) == 8
    for v in inputs:
        assert v.ndim == 2 and v.dtype in (np.float32, np.int64) and v.shape[0] == n
    hid = BaseModel.get_harmonic_id(inputs).reshape(-1)
    n_spots = hid.max() + 1
    assert np.array_equal(np.unique(hid), np.arange(n_spots))
    assert np.all(inputs[4][n_spots:] == 1.0) and np.all(inputs[5][n_spots:] == 1.0)      # padding
    wl = BaseModel.get_wavelength(inputs).reshape(-1)
    if lam_min is not None:
        assert wl.min() >= lam_min and wl.max() <= lam_max
    # all rows of a spot sit on one image
    image_id = inputs[1].reshape(-1)
    assert np.all(np.bincount(hid, weights=image_id) == np.bincount(hid) * image_id[np.unique(hid, return_index=True)[1]])


def test_expand_harmonics():
    """careless/utils/laue.py:9-81: rows = sum over spots of floor(d_0 / dmin); n H_0 = H; lambda_n = lambda_0 / n."""
    ds = U.load_fixture("pyp_off")
    ds.compute_dHKL()
    dmin = 2.0
    ex = expand_harmonics(ds, dmin, "Wavelength")
    H = ds.get_hkls()
    n_obs = np.gcd.reduce(H, axis=1)
    d0 = ds["dHKL"].astype(np.float64) * n_obs
    assert len(ex) == int(np.floor(d0 / dmin).sum())
    H0 = np.stack([ex["H_0"], ex["K_0"], ex["L_0"]], 1)
    assert np.all(np.gcd.reduce(H0, axis=1) == 1)
    n = np.gcd.reduce(ex.get_hkls(), axis=1)
    assert np.array_equal(ex.get_hkls(), n[:, None] * H0)
    assert np.all(ex["dHKL"] >= np.float32(dmin) * (1 - 1e-6))
    # every original observation is among the expanded rows with its own wavelength
    key = lambda h, w: set(zip(map(tuple, h.tolist()), np.round(w.astype(np.float64), 5).tolist()))
    assert key(H, ds["Wavelength"]) <= key(ex.get_hkls(), ex["Wavelength"])


def test_positional_encoding_shape_and_range():
    x = np.random.default_rng(0).random((50, 2)).astype(np.float32)
    e = positional_encoding(x, 4)
    assert e.shape == (50, 16) and np.all(np.abs(e) <= 1.0 + 1e-6)


def test_mtz_round_trip(tmp_path):
    ds = U.load_fixture("pyp_off")
    ds.hkl_to_asu()
    path = os.path.join(tmp_path, "rt.mtz")
    write_mtz(path, ds)
    back = read_mtz(path, to_observed=False)
    assert back.keys() == ds.keys()
    for k in ds.keys():
        assert np.array_equal(np.asarray(back[k], dtype=np.float32), np.asarray(ds[k], dtype=np.float32)), k
    assert back.cell.parameters == pytest.approx(ds.cell.parameters)
    assert [o.triplet() for o in back.spacegroup.sym_ops] == [o.triplet() for o in ds.spacegroup.sym_ops]
    obs = read_mtz(path)
    assert np.array_equal(obs.get_hkls(), U.load_fixture("pyp_off").get_hkls())


def test_crystfel_stream_reader(tmp_path):
    """rs.read_crystfel as careless uses it (tests/test_cli.py:112-119): observed indices, I / SigI, crystal number as BATCH."""
    from careless_b200.io.crystfel import read_crystfel
    path = os.path.join(tmp_path, "synthetic.stream")
    crystals = U.synthetic_stream(path, n_crystals=4, n_refl=50, seed=3)
    ds = read_crystfel(path)
    hkl = np.concatenate([c["hkl"] for c in crystals])
    assert np.array_equal(ds.get_hkls(), hkl)
    assert np.allclose(ds["I"], np.concatenate([c["I"] for c in crystals]), atol=5e-3)
    assert np.array_equal(ds["BATCH"], np.concatenate([np.full(len(c["hkl"]), i) for i, c in enumerate(crystals)]))
    assert ds.dtypes["BATCH"] == "B" and ds.dtypes["I"] == "J" and ds.dtypes["SigI"] == "Q" and not ds.merged
    assert ds.cell.a == pytest.approx(10 * np.mean([c["cell"][0] for c in crystals]), abs=1e-3) and ds.spacegroup is None
    f = MonoFormatter(None, None, None, ["dHKL", "image_id"], False, False, spacegroups=["1"])
    inputs, rac = f.format_files([path])
    assert inputs[0].shape[0] == len(hkl) and inputs[1].max() == 3
    with pytest.raises(ValueError):
        MonoFormatter(None, None, None, ["dHKL"], False, False).format_files([path])       # a stream has no space group
    with pytest.raises(ValueError):
        LaueFormatter("Wavelength", None, None, None, ["dHKL"], False, False).format_files([path])
    ref = "/root/reference/tests/data/crystfel.stream"
    if os.path.exists(ref):          # the reference's own fixture (build container only)
        real = read_crystfel(ref)
        assert len(real) == 618 and real["BATCH"].max() == 2 and real.cell.c == pytest.approx(38.76, abs=0.01)


def test_spacegroup_by_number():
    assert S.SpaceGroup.from_name("1").name == "P 1" and S.SpaceGroup.from_name("96").name == "P 43 21 2"
    assert S.SpaceGroup.from_name("P212121").number == 19
    with pytest.raises(ValueError):
        S.SpaceGroup.from_name("P -1")


# ---- property tests (hypothesis) --------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st


@settings(max_examples=40, deadline=None)
@given(st.sampled_from(sorted(S._SOHNCKE)), st.integers(0, 2 ** 31 - 1))
def test_asu_mapping_properties(name, seed):
    """For any tabulated group and any reflections: the image is in the wedge, the stored operation maps back to the
    observation, mapping is idempotent, and symmetry mates (Friedel mates included) share one ASU reflection."""
    sg = S.SpaceGroup.from_name(name)
    rng = np.random.default_rng(seed)
    hkl = rng.integers(-15, 16, size=(64, 3))
    hkl = hkl[np.any(hkl != 0, axis=1)]
    asu, isym = sg.hkl_to_asu(hkl)
    assert sg.in_asu(asu).all()
    assert np.array_equal(sg.hkl_to_observed(asu, isym), hkl)
    again, isym2 = sg.hkl_to_asu(asu)
    assert np.array_equal(again, asu)
    op = sg.sym_ops[int(rng.integers(len(sg.sym_ops)))]
    sign = int(rng.choice([-1, 1]))
    mates, _ = sg.hkl_to_asu(sign * op.apply_to_hkl(hkl))
    assert np.array_equal(mates, asu)
    # Friedel mates of acentric reflections are told apart by the parity of M/ISYM
    mates_minus, isym_minus = sg.hkl_to_asu(-hkl)
    acentric = ~sg.is_centric(hkl)
    assert np.all((isym_minus[acentric] % 2) != (isym[acentric] % 2))


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 400), n_extra=st.integers(1, 6), seed=st.integers(0, 2 ** 31 - 1))
def test_mtz_round_trip_property(n, n_extra, seed):
    import tempfile
    rng = np.random.default_rng(seed)
    sg = S.SpaceGroup.from_name(sorted(S._SOHNCKE)[seed % len(S._SOHNCKE)])
    cell = S.UnitCell(40. + seed % 7, 50., 60. + seed % 11, 90., 90., 90.)
    from careless_b200.io.mtz import DataSet
    cols = {"H": rng.integers(-30, 31, n).astype(np.int32), "K": rng.integers(-30, 31, n).astype(np.int32),
            "L": rng.integers(1, 31, n).astype(np.int32)}
    types = {"H": "H", "K": "H", "L": "H"}
    for j in range(n_extra):
        cols[f"C{j}"] = rng.standard_normal(n).astype(np.float32) * 10.0 ** rng.integers(-3, 6)
        types[f"C{j}"] = "R"
    ds = DataSet(cols, types, cell, sg, merged=True)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "p.mtz")
        write_mtz(path, ds)
        back = read_mtz(path)
    assert back.keys() == ds.keys() and len(back) == n
    for k in cols:
        assert np.array_equal(np.asarray(back[k], dtype=np.float32), cols[k].astype(np.float32)), k
    assert back.spacegroup.laue == sg.laue and len(back.spacegroup.sym_ops) == len(sg.sym_ops)
    assert len(back.spacegroup.cen_ops) == len(sg.cen_ops)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 5), st.integers(-3, 3)), min_size=1, max_size=60))
def test_ngroup_property(pairs):
    a = np.array([p[0] for p in pairs]); b = np.array([p[1] for p in pairs])
    g = ngroup(a, b)
    uniq = sorted(set(pairs))
    assert np.array_equal(g, np.array([uniq.index(p) for p in pairs]))


def test_command_line_options_match_the_reference_defaults():
    """careless/args/*.py: option names, destinations and defaults (the values DataManager.build_model reads)."""
    from careless_b200.careless import build_argparser, default_parser, make_formatter
    ns = build_argparser().parse_args(["mono", "dHKL,image_id", "a.mtz", "b.mtz", "out/base"])
    expect = dict(type="mono", mc_samples=1, epsilon=1e-7, iterations=10000, learning_rate=0.001, beta_1=0.9, beta_2=0.99,
                  mlp_layers=20, mlp_width=10, image_layers=0, use_image_scales=True, scale_bijector="exp", kl_weight=None,
                  studentt_likelihood_dof=None, refine_uncertainties=False, standardize_metadata=True, seed=1234,
                  validation_frequency=10, half_dataset_repeats=1, structure_factor_init_scale=1.0, anomalous=False,
                  separate_files=False, positional_encoding_frequencies=4, reflection_files=["a.mtz", "b.mtz"], output_base="out/base")
    for k, v in expect.items():
        assert getattr(ns, k) == v, k
    d = default_parser("mono")
    for k in expect:
        if k not in ("reflection_files", "output_base"):
            assert getattr(d, k) == getattr(ns, k), k
    ns = build_argparser().parse_args(["poly", "--disable-image-scales", "--disable-metadata-standardization", "-l", "0.95", "1.2",
                                       "--studentt-likelihood-dof", "16", "--double-wilson-parents", "None,0", "--double-wilson-r", "0.,0.99",
                                       "--separate-files", "-d", "2.0", "-c", "1.5", "dHKL,Wavelength", "a.mtz", "b.mtz", "o"])
    assert ns.use_image_scales is False and ns.standardize_metadata is False and ns.wavelength_range == [0.95, 1.2]
    assert ns.parents == "None,0" and ns.dwr == "0.,0.99" and ns.dmin == 2.0 and ns.isigi_cutoff == 1.5
    f = make_formatter(ns)
    assert type(f).__name__ == "LaueFormatter" and f.lam_min == 0.95 and f.lam_max == 1.2 and f.standardize is False
    with pytest.raises(TypeError):
        default_parser("mono", not_an_option=1)
    with pytest.raises(ValueError):
        ns.spacegroups = "P 1,P 1,P 1"
        make_formatter(ns)
