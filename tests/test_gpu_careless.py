"""GPU tests of the careless-shaped driver on the reference's own fixture data (BASELINE.json configs[0]):
files -> formatter -> DataManager.build_model -> train_model -> results / predictions / cross-validation outputs.
Modelled on tests/test_cli.py:20-120 of the reference (which only checks that the output files exist); here the
merged F / SigF are also compared with the oracle trained on the same Philox draws."""
import os

import numpy as np
import pytest

import _util as U
from careless_b200.careless import default_parser, run_careless
from careless_b200.io.mtz import read_mtz
from oracle import model as om
from oracle import philox

pytestmark = pytest.mark.gpu


def _problem(inputs, rac):
    return {"refl_id": inputs[0].reshape(-1), "image_id": inputs[1].reshape(-1), "metadata": inputs[3],
            "intensities": inputs[4].reshape(-1), "uncertainties": inputs[5].reshape(-1),
            "centric": rac.centric, "multiplicity": rac.multiplicity, "n_images": int(inputs[1].max()) + 1}


def test_config0_mono_fixture_matches_oracle(tmp_path):
    """configs[0]: careless mono on the unmerged fixture MTZ, WilsonPrior, NormalLikelihood, MLPScaler, 1 MC sample."""
    steps = 60
    parser = default_parser("mono", metadata_keys="dHKL,Hobs,Kobs,Lobs,X,Y", output_base=os.path.join(tmp_path, "pyp"),
                            iterations=steps, use_image_scales=False, disable_progress_bar=True, learning_rate=0.01)
    run = run_careless(parser, datasets=[U.load_fixture("pyp_off")])
    dm, hist = run["data_manager"], run["history"]
    assert len(hist["loss"]) == steps and np.all(np.isfinite(hist["loss"]))
    p = _problem(dm.inputs, dm.asu_collection)
    R, N = len(p["centric"]), len(p["refl_id"])
    ocfg = om.ModelConfig(n_refl=R, n_meta=6, mlp_width=10, mlp_layers=20)
    oprior = om.PriorData(p["centric"], p["multiplicity"])
    params = om.init_params(ocfg, oprior)
    draws = [(philox.refl_uniforms(parser.seed, s, 1, np.arange(R)), philox.obs_normals(parser.seed, s, 1, np.arange(N))) for s in range(steps)]
    oparams, ohist, _ = om.train(params, p, oprior, ocfg, om.AdamConfig(lr=0.01), draws)
    for i in (0, 1, steps // 2, steps - 1):
        for k in ("loss", "NLL", "F KLDiv", "Grad Norm"):
            assert abs(hist[k][i] - ohist[i][k]) <= 1e-3 * abs(ohist[i][k]) + 1e-6, (i, k, hist[k][i], ohist[i][k])
    loc = np.exp(oparams["sf_loc_raw"].numpy()); scale = np.exp(oparams["sf_scale_raw"].numpy()) + ocfg.eps
    low = 1e-32 * ~np.asarray(p["centric"], dtype=bool)
    Fo, So, _ = om.tn_moments(loc, scale, low, 1e10)
    res = run["results"][0]
    refl = dm.asu_collection.to_refl_id(np.zeros((len(res), 1)), res.get_hkls())
    eF, eS = U.rel_err(res["F"], Fo[refl]), U.rel_err(res["SigF"], So[refl])
    cc = np.corrcoef(res["F"], Fo[refl])[0, 1]
    print(f"\n[config0] steps={steps} F rel err {eF:.2e}  SigF rel err {eS:.2e}  CC {cc:.6f}  final loss gpu {hist['loss'][-1]:.6f} oracle {ohist[-1]['loss']:.6f}")
    assert cc >= 0.999 and eF <= 1e-3 and eS <= 1e-3
    assert np.array_equal(np.sort(refl), np.unique(p["refl_id"]))          # exactly the observed reflections, N > 0
    assert np.array_equal(res["N"], np.bincount(p["refl_id"], minlength=R)[refl])
    # files the reference driver writes (careless.py:66-98)
    out = parser.output_base
    merged = read_mtz(out + "_0.mtz")
    assert merged.keys()[:8] == ["H", "K", "L", "F", "SigF", "I", "SigI", "N"] and np.array_equal(merged["F"], res["F"])
    assert merged.spacegroup.name == "P 63"
    pred = read_mtz(out + "_predictions_0.mtz")
    assert len(pred) == N and {"Iobs", "SigIobs", "Ipred", "SigIpred", "Scale", "SigScale", "test"} <= set(pred.keys())
    assert np.all(np.isfinite(pred["Ipred"])) and np.all(pred["SigIpred"] > 0)
    assert os.path.exists(out + "_history.csv") and os.path.exists(out + "_structure_factor.npz")
    run["model"].close()


@pytest.mark.parametrize("mode", ["mono", "poly"])
@pytest.mark.parametrize("flags", [
    dict(),
    dict(separate_files=True, anomalous=True, studentt_likelihood_dof=16.0, test_fraction=0.2, merge_half_datasets=True),
    dict(refine_uncertainties=True, image_layers=1, mlp_width=8, mc_samples=2, positional_encoding_keys="X,Y", positional_encoding_frequencies=2),
    dict(separate_files=True, parents="None,0", dwr="0.,0.9", optimize_double_wilson_r=True, scale_bijector="softplus", dmin=2.5),
])
def test_run_careless_outputs(tmp_path, mode, flags):
    """tests/test_cli.py:63-120: every flag combination trains and writes readable, finite outputs."""
    parser = default_parser(mode, metadata_keys="dHKL,Hobs,Kobs,Lobs" + (",Wavelength" if mode == "poly" else ""),
                            output_base=os.path.join(tmp_path, "out"), iterations=12, mlp_layers=3, disable_progress_bar=True,
                            validation_frequency=4, **flags)
    run = run_careless(parser, datasets=[U.load_fixture("pyp_off"), U.load_fixture("pyp_2ms")])
    hist = run["history"]
    assert len(hist["loss"]) == 12 and all(np.all(np.isfinite(v)) for v in hist.values())
    n_asu = 2 if flags.get("separate_files") else 1
    assert len(run["results"]) == n_asu
    halves = set()
    for i in range(n_asu):
        ds = read_mtz(parser.output_base + f"_{i}.mtz")
        fkey = "F(+)" if flags.get("anomalous") else "F"
        assert len(ds) > 0 and np.all(np.isfinite(ds[fkey][~np.isnan(ds[fkey])])) and np.nanmin(ds[fkey]) > 0
        pred = read_mtz(parser.output_base + f"_predictions_{i}.mtz")
        assert np.all(np.isfinite(pred["Ipred"])) and np.all(np.isfinite(pred["SigIpred"]))
        if flags.get("test_fraction"):
            assert set(np.unique(pred["test"])) == {0, 1} and "NLL_val" in hist
        if flags.get("merge_half_datasets"):
            xv = read_mtz(parser.output_base + f"_xval_{i}.mtz")
            halves |= set(np.unique(xv["half"]).tolist())      # ten images: one file's images may all land in one half
            assert set(np.unique(xv["repeat"])) == {0} and len(xv) > 0
    if flags.get("merge_half_datasets"):
        assert halves == {0, 1}
    if flags.get("refine_uncertainties"):
        lik = run["model"].likelihood
        assert not np.allclose([lik.Sdfac, lik.Sdadd, lik.SdB], 1.0) and min(lik.Sdfac, lik.Sdadd, lik.SdB) > 0
    if flags.get("optimize_double_wilson_r"):
        assert run["model"].prior.r[0] == 0.0 and run["model"].prior.r[1] != np.float32(0.9)
    run["model"].close()


def test_command_line_entry_point(tmp_path):
    """`python -m careless_b200.careless mono ...` on MTZ files written to disk (argparse + read_mtz + the whole flow)."""
    from careless_b200.careless import main
    from careless_b200.io.mtz import write_mtz
    files = []
    for name in ("pyp_off", "pyp_2ms"):
        ds = U.load_fixture(name)
        ds.hkl_to_asu()                       # unmerged MTZ convention: ASU indices + M/ISYM, as the reference's fixtures
        path = os.path.join(tmp_path, name + ".mtz")
        write_mtz(path, ds)
        files.append(path)
    out = os.path.join(tmp_path, "cli")
    main(["mono", "--iterations", "8", "--mlp-layers", "2", "--disable-progress-bar", "--separate-files", "--studentt-likelihood-dof", "8",
          "--mc-samples", "2", "dHKL,Hobs,Kobs,Lobs,X,Y", files[0], files[1], out])
    for i in range(2):
        merged = read_mtz(out + f"_{i}.mtz")
        assert len(merged) > 50 and np.all(np.isfinite(merged["F"])) and np.all(merged["SigF"] > 0)
    lines = open(out + "_history.csv").read().strip().splitlines()
    assert lines[0].split(",")[:3] == ["step", "loss", "NLL"] and len(lines) == 9
    main(["poly", "--iterations", "4", "--mlp-layers", "2", "--disable-progress-bar", "-w", "Wavelength", "dHKL,Wavelength",
          files[0], out + "_laue"])
    assert len(read_mtz(out + "_laue_0.mtz")) > 50


def test_crystfel_stream_entry_point(tmp_path):
    """tests/test_cli.py:112-119: `careless mono --spacegroups=1 dHKL,image_id file.stream out`; poly refuses streams."""
    from careless_b200.careless import main
    path = os.path.join(tmp_path, "synthetic.stream")
    U.synthetic_stream(path, n_crystals=6, n_refl=150, seed=5)
    out = os.path.join(tmp_path, "sfx")
    main(["mono", "--iterations", "10", "--mlp-layers", "2", "--disable-progress-bar", "--spacegroups=1", "dHKL,image_id", path, out])
    merged = read_mtz(out + "_0.mtz")
    assert len(merged) > 100 and np.all(np.isfinite(merged["F"])) and merged.spacegroup.name == "P 1"
    with pytest.raises(ValueError):
        main(["poly", "--iterations", "2", "--disable-progress-bar", "--spacegroups=1", "dHKL,image_id", path, out + "_p"])


def _scaler_arrays(sm):
    """Every trainable array of a scaling model mirror (MLPScaler / HybridImageScaler / NeuralImageScaler)."""
    if hasattr(sm, "mlp_scaler"):
        return [w.copy() for w in sm.mlp_scaler.get_weights()] + [np.array(sm.image_scaler._scales, copy=True)]
    if hasattr(sm, "metadata_scaler"):
        return [w.copy() for w in sm.metadata_scaler.get_weights()] + [np.array(sm.flat(), copy=True)]
    return [w.copy() for w in sm.get_weights()]


@pytest.mark.parametrize("flags", [dict(), dict(use_image_scales=False), dict(image_layers=1, mlp_width=8)])
def test_scale_file_and_structure_factor_file_on_a_fresh_model(tmp_path, flags):
    """careless.py:48-56, 79-80, 104: `--scale-file` / `--structure-factor-file` load the weights of an earlier run into a model that has
    not been built yet; with `--freeze-scales --freeze-structure-factors` and no training noise the second run reproduces the first
    run's merged structure factors exactly and leaves both weight sets untouched."""
    common = dict(metadata_keys="dHKL,Hobs,Kobs,Lobs", iterations=10, mlp_layers=3, disable_progress_bar=True, **flags)
    first = default_parser("mono", output_base=os.path.join(tmp_path, "a"), **common)
    run1 = run_careless(first, datasets=[U.load_fixture("pyp_off")])
    w1 = _scaler_arrays(run1["model"].scaling_model)
    loc1 = run1["model"].surrogate_posterior.loc_raw.copy()
    f1 = read_mtz(first.output_base + "_0.mtz")
    run1["model"].close()
    second = default_parser("mono", output_base=os.path.join(tmp_path, "b"), scale_file=first.output_base + "_scale",
                            structure_factor_file=first.output_base + "_structure_factor", freeze_scales=True,
                            freeze_structure_factors=True, **common)
    run2 = run_careless(second, datasets=[U.load_fixture("pyp_off")])
    w2 = _scaler_arrays(run2["model"].scaling_model)
    assert len(w1) == len(w2) and all(np.array_equal(a, b) for a, b in zip(w1, w2))
    assert np.array_equal(loc1, run2["model"].surrogate_posterior.loc_raw)
    f2 = read_mtz(second.output_base + "_0.mtz")
    assert np.array_equal(np.asarray(f1["F"]), np.asarray(f2["F"])) and np.array_equal(np.asarray(f1["SigF"]), np.asarray(f2["SigF"]))
    assert len(run2["history"]["loss"]) == 10 and np.all(np.isfinite(run2["history"]["loss"]))
    run2["model"].close()
