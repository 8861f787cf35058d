#!/usr/bin/env python
"""`careless mono|poly` driver on the CUDA engine.

Mirror of careless/careless.py:11-139 (`run_careless`) and of the option names/defaults of careless/args/*.py:
format the reflection files, (optionally) hold out a test fraction, train, write `{out}_{i}.mtz`,
`{out}_history.csv`, weights and `{out}_predictions_{i}.mtz`, then (optionally) re-merge random half datasets with
the scale model frozen into `{out}_xval_{i}.mtz`.

    python -m careless_b200.careless mono [options] metadata_keys file.mtz [file.mtz ...] out_base
"""
from __future__ import annotations

import argparse
import csv
import sys

import numpy as np

# (flags, dest, type, default) -- careless/args/*.py
_OPTIONS = [
    (("--mc-samples",), "mc_samples", int, 1),
    (("--structure-factor-file",), "structure_factor_file", str, None),
    (("--freeze-structure-factors",), "freeze_structure_factors", bool, False),
    (("--structure-factor-init-scale",), "structure_factor_init_scale", float, 1.0),
    (("--epsilon",), "epsilon", float, 1e-7),
    (("--disable-metadata-standardization",), "standardize_metadata", "store_false", True),
    (("--disable-progress-bar",), "disable_progress_bar", bool, False),
    (("--test-fraction",), "test_fraction", float, None),
    (("--merge-half-datasets",), "merge_half_datasets", bool, False),
    (("--half-dataset-repeats",), "half_dataset_repeats", int, 1),
    (("--validation-frequency",), "validation_frequency", int, 10),
    (("-c", "--isigi-cutoff"), "isigi_cutoff", float, None),
    (("-d", "--dmin"), "dmin", float, None),
    (("--spacegroups",), "spacegroups", str, None),
    (("--image-key",), "image_key", str, None),
    (("--intensity-key",), "intensity_key", str, None),
    (("--uncertainty-key",), "uncertainty_key", str, None),
    (("--anomalous",), "anomalous", bool, False),
    (("--separate-files",), "separate_files", bool, False),
    (("--studentt-likelihood-dof",), "studentt_likelihood_dof", float, None),
    (("--refine-uncertainties",), "refine_uncertainties", bool, False),
    (("--iterations",), "iterations", int, 10000),
    (("--learning-rate",), "learning_rate", float, 0.001),
    (("--beta-1",), "beta_1", float, 0.9),
    (("--beta-2",), "beta_2", float, 0.99),
    (("--clipnorm",), "clipnorm", float, None),
    (("--clipvalue",), "clipvalue", float, None),
    (("--global-clipnorm",), "global_clipnorm", float, None),
    (("--positional-encoding-keys",), "positional_encoding_keys", str, None),
    (("--positional-encoding-frequencies", "-L"), "positional_encoding_frequencies", int, 4),
    (("--kl-weight",), "kl_weight", float, None),
    (("--wilson-prior-b",), "wilson_prior_b", float, None),
    (("--double-wilson-r",), "dwr", str, None),
    (("--double-wilson-parents",), "parents", str, None),
    (("--double-wilson-reindexing-ops",), "reindexing_ops", str, None),
    (("--optimize-double-wilson-r",), "optimize_double_wilson_r", bool, False),
    (("--scale-file",), "scale_file", str, None),
    (("--freeze-scales",), "freeze_scales", bool, False),
    (("--mlp-layers",), "mlp_layers", int, 20),
    (("--mlp-width",), "mlp_width", int, 10),
    (("--image-layers",), "image_layers", int, 0),
    (("--disable-image-scales",), "use_image_scales", "store_false", True),
    (("--scale-bijector",), "scale_bijector", str, "exp"),
    (("--gpu-id",), "gpu_id", int, 0),
    (("--seed",), "seed", int, 1234),
]
_POLY_OPTIONS = [
    (("-l", "--wavelength-range"), "wavelength_range", "two_floats", None),
    (("-w", "--wavelength-key"), "wavelength_key", str, "Wavelength"),
]


def default_parser(type="mono", **overrides):
    """A Namespace with every option at its careless default (what `parser.parse_args()` would return)."""
    ns = argparse.Namespace(type=type, metadata_keys=None, reflection_files=None, output_base=None,
                            run_eagerly=False, jit_compile=None, reduce_retracing=False, embed=False, save_data_manager=False)
    for _, dest, _, default in _OPTIONS + _POLY_OPTIONS:
        setattr(ns, dest, default)
    for k, v in overrides.items():
        if not hasattr(ns, k):
            raise TypeError(f"unknown careless option {k!r}")
        setattr(ns, k, v)
    return ns


def build_argparser():
    ap = argparse.ArgumentParser(prog="careless_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="type", required=True)
    for mode in ("mono", "poly"):
        sp = sub.add_parser(mode)
        for flags, dest, typ, default in _OPTIONS + (_POLY_OPTIONS if mode == "poly" else []):
            if typ is bool:
                sp.add_argument(*flags, dest=dest, action="store_true", default=default)
            elif typ == "store_false":
                sp.add_argument(*flags, dest=dest, action="store_false", default=default)
            elif typ == "two_floats":
                sp.add_argument(*flags, dest=dest, type=float, nargs=2, default=default)
            else:
                sp.add_argument(*flags, dest=dest, type=typ, default=default)
        sp.add_argument("metadata_keys", type=str)
        sp.add_argument("reflection_files", type=str, nargs="+")
        sp.add_argument("output_base", type=str)
    return ap


def make_formatter(parser):
    from .io.formatter import LaueFormatter, MonoFormatter
    pe = parser.positional_encoding_keys.split(",") if parser.positional_encoding_keys else None
    sgs = None
    if parser.spacegroups is not None:
        sgs = parser.spacegroups.split(",")
        if len(sgs) == 1:
            sgs = sgs * len(parser.reflection_files)
        elif len(sgs) != len(parser.reflection_files):
            raise ValueError("Multiple values provided for --spacegroups=, but the number of provided values does not match "
                             "the number of reflection files.")
    keys = parser.metadata_keys.split(",")
    if parser.type == "poly":
        lmin, lmax = parser.wavelength_range if parser.wavelength_range is not None else (None, None)
        return LaueFormatter(parser.wavelength_key, parser.intensity_key, parser.uncertainty_key, parser.image_key, keys,
                             parser.separate_files, parser.anomalous, lmin, lmax, parser.dmin, parser.isigi_cutoff, pe,
                             parser.positional_encoding_frequencies, sgs, standardize=parser.standardize_metadata)
    return MonoFormatter(parser.intensity_key, parser.uncertainty_key, parser.image_key, keys, parser.separate_files,
                         parser.anomalous, 0. if parser.dmin is None else parser.dmin, parser.isigi_cutoff, pe,
                         parser.positional_encoding_frequencies, sgs, standardize=parser.standardize_metadata)


def _concat_rows(a, b):
    from .io.mtz import concat
    return concat([a, b])


def run_careless(parser, datasets=None):
    """careless.py:11-139.  `datasets` (already loaded DataSet objects) may replace `parser.reflection_files`."""
    from . import parallel
    from .io.manager import DataManager
    from .io.mtz import write_mtz as _write_mtz

    # `torchrun --nproc-per-node N -m careless_b200.careless ...`: every process formats the same inputs (the splits below
    # draw from the same seeded generator), trains its share of the reflections on its own GPU and takes part in the
    # gathers; only rank 0 writes the output files.
    ctx = parallel.context()
    chief = ctx.rank == 0

    def write_mtz(path, ds):
        if chief:
            _write_mtz(path, ds)

    np.random.seed(parser.seed)
    df = make_formatter(parser)
    inputs, rac = df(datasets) if datasets is not None else df.format_files(parser.reflection_files)
    dm = DataManager(inputs, rac, parser=parser)
    if parser.test_fraction is not None:
        train, test = dm.split_data_by_refl(parser.test_fraction)
    else:
        train, test = dm.inputs, None

    model = dm.build_model()
    if parser.scale_file is not None:
        model.scaling_model.load_weights(parser.scale_file)
    if parser.freeze_scales:
        model.scaling_model.trainable = False
    if parser.structure_factor_file is not None:
        model.surrogate_posterior.load_weights(parser.structure_factor_file)
    if parser.freeze_structure_factors:
        model.surrogate_posterior.trainable = False
    progress = not parser.disable_progress_bar

    history = model.train_model(train, parser.iterations, message="Training", validation_data=test,
                                validation_frequency=parser.validation_frequency, progress=progress)
    out = parser.output_base
    results = dm.get_results(model.surrogate_posterior, inputs=train, model=model)
    for i, ds in enumerate(results):
        write_mtz(out + f"_{i}.mtz", ds)
    keys = list(history.keys())
    if chief:
        with open(out + "_history.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["step"] + keys)
            for step in range(len(history[keys[0]]) if keys else 0):
                w.writerow([step] + [history[k][step] for k in keys])
        model.surrogate_posterior.save_weights(out + "_structure_factor")
        model.scaling_model.save_weights(out + "_scale")

    if test is not None:
        pairs = zip(dm.get_predictions(model, train, test_value=0), dm.get_predictions(model, test, test_value=1))
        for file_id, (ds_train, ds_test) in enumerate(pairs):
            write_mtz(out + f"_predictions_{file_id}.mtz", _concat_rows(ds_train, ds_test))
    else:
        for file_id, ds_train in enumerate(dm.get_predictions(model, train, test_value=0)):
            write_mtz(out + f"_predictions_{file_id}.mtz", ds_train)

    xval = None
    if parser.merge_half_datasets:
        scaling_model = model.scaling_model
        scaling_model.trainable = False
        xval = [None] * len(dm.asu_collection)
        for repeat in range(parser.half_dataset_repeats):
            for half_id, half in enumerate(dm.split_data_by_image()):
                hmodel = dm.build_model(scaling_model=scaling_model)
                hmodel.train_model(half, parser.iterations, message=f"Merging repeat {repeat + 1} half {half_id + 1}", progress=progress)
                for file_id, ds in enumerate(dm.get_results(hmodel.surrogate_posterior, inputs=half, model=hmodel)):
                    ds["repeat"] = np.full(len(ds), repeat, dtype=np.int32)
                    ds["half"] = np.full(len(ds), half_id, dtype=np.int32)
                    xval[file_id] = ds if xval[file_id] is None else _concat_rows(xval[file_id], ds)
                hmodel.close()
        for file_id, ds in enumerate(xval):
            write_mtz(out + f"_xval_{file_id}.mtz", ds)
    ctx.barrier()
    return {"model": model, "data_manager": dm, "history": history, "results": results, "train": train, "test": test, "xval": xval}


def main(argv=None):
    parser = build_argparser().parse_args(argv)
    run = run_careless(parser)
    run["model"].close()


if __name__ == "__main__":
    main(sys.argv[1:])
