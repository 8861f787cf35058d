#!/usr/bin/env python
"""Single-GPU timing of BASELINE.json configs[2..4] (the parity-test configurations, NOT the bench line).

    python tools/bench_configs.py [--which laue,dw,stills] [--scale 1.0] [--steps 10]

configs[2]: synthetic Laue, 20 M harmonic rows, 1 M unique reflections
configs[3]: 4-dataset DoubleWilson merge, 40 M observations (4 x 10 M, 4 x 500 k reflections)
configs[4]: stills, one GPU's share of the 8-GPU job: 25 M observations, 250 k reflections, image layers
All with StudentT(12) + MLPScaler 32 x 20, d = 5 like configs[1].  Prints one JSON line per config.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from careless_b200 import synth
    from careless_b200.engine import Engine, EngineConfig

    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="laue,dw,stills")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--width", type=int, default=32)
    ap.add_argument("--layers", type=int, default=20)
    args = ap.parse_args()
    stream = torch.cuda.Stream(device=0)
    torch.cuda.set_stream(stream)
    s = args.scale
    for which in args.which.split(","):
        t0 = time.perf_counter()
        common = dict(n_meta=5, mlp_width=args.width, mlp_layers=args.layers, likelihood="studentt", dof=12.0, seed=1234,
                      device=0, stream=stream.cuda_stream)
        if which == "laue":
            N, R = int(20e6 * s), int(1e6 * s)
            p = synth.make_laue(N, R, d=5, n_images=max(2, int(10000 * s)), seed=5)
            cfg = EngineConfig(n_refl=R, laue=True, **common)
        elif which == "dw":
            n, r = int(10e6 * s), int(500e3 * s)
            p = synth.make_double_wilson(n, r, n_datasets=4, d=5, n_images=max(2, int(2500 * s)), r=0.99, seed=6)
            N, R = 4 * n, 4 * r
            cfg = EngineConfig(n_refl=R, prior="double_wilson", n_asu=4, **common)
        elif which == "stills":
            N, R = int(25e6 * s), int(250e3 * s)
            n_img = max(2, int(12500 * s))
            p = synth.make_mono(N, R, d=5, n_images=n_img, seed=7)
            cfg = EngineConfig(n_refl=R, n_images=n_img, image_layers=2, **common)
        elif which == "mono":
            N, R = int(10e6 * s), int(500e3 * s)
            p = synth.make_mono(N, R, d=5, n_images=5000, seed=8)
            cfg = EngineConfig(n_refl=R, **common)
        else:
            raise SystemExit(f"unknown config {which}")
        t_synth = time.perf_counter() - t0
        eng = Engine(cfg)
        t0 = time.perf_counter()
        eng.set_observations(p["refl_id"], p["image_id"], p["metadata"], p["intensities"], p["uncertainties"],
                             harmonic_id=p.get("harmonic_id") if which == "laue" else None)
        if which == "dw":
            eng.set_prior(p["centric"], p["multiplicity"], None, dw_parent=p["dw_parent"], asu_id=p["asu_id"], r=p["r"])
        else:
            eng.set_prior(p["centric"], p["multiplicity"])
        eng.synchronize()
        t_prep = time.perf_counter() - t0
        for _ in range(3):
            eng.step_begin(); eng.step_norms(); eng.step_end(False)
        torch.cuda.synchronize()
        eng.reset_timers(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            eng.step_begin(); eng.step_norms(); eng.step_end(False)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        kt = eng.kernel_times()
        eng.reset_timers(False)
        last = eng.step(1)[0]
        n_rows = len(p["refl_id"])
        print(json.dumps({"config": which, "rows": n_rows, "reflections": R, "ms_per_step": ms, "obs_per_s": n_rows / (ms * 1e-3),
                          "obs_kernel_ms": kt["obs_kernel_ms"] / max(1, kt["obs_kernel_launches"]),
                          "launches_per_step": kt["total_launches"] / args.steps, "host_prep_s": t_prep, "synth_s": t_synth,
                          "last_metrics": last, "mem_gb": torch.cuda.mem_get_info(0)}), flush=True)
        eng.close()
        del p


if __name__ == "__main__":
    main()
