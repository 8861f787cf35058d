#!/usr/bin/env python
"""Row preparation inside clb_set_observations: device (default) against host (CLB_DEVICE_PREP=0), wall time per call.

    python tools/bench_prep.py [--n 10000000,50000000] [--refl-per 20]
One JSON line per size and mode: rows, mode, prep_ms (row preparation only, measured inside the library), call_ms (whole call).
"""
import argparse, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(n, refl_per, mode):
    from careless_b200.engine import Engine, EngineConfig
    R = max(1, n // refl_per)
    rng = np.random.default_rng(0)
    refl = rng.integers(0, R, n, dtype=np.int64)
    img = rng.integers(0, 1000, n, dtype=np.int64)
    meta = rng.standard_normal((n, 5), dtype=np.float32)
    iobs = rng.standard_normal(n, dtype=np.float32); sig = np.abs(iobs) + 1.0
    eng = Engine(EngineConfig(n_refl=R, n_meta=5, mlp_width=32, mlp_layers=2, n_images=1000, image_scales=True))
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        eng.set_observations(refl, img, meta, iobs, sig)
        eng.synchronize()
        call_ms = 1e3 * (time.perf_counter() - t0)
        prep_ms = eng.download_rows_info()["prep_ms"]
        if best is None or call_ms < best[1]:
            best = (prep_ms, call_ms)
    eng.close()
    print(json.dumps({"rows": n, "reflections": R, "mode": mode, "prep_ms": round(best[0], 2), "call_ms": round(best[1], 2)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="10000000,50000000")
    ap.add_argument("--refl-per", type=int, default=20)
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child:
        one(int(a.n), a.refl_per, a.child)
    else:
        for n in a.n.split(","):
            for mode, env in (("device", {}), ("host", {"CLB_DEVICE_PREP": "0"})):
                subprocess.run([sys.executable, __file__, "--n", n, "--refl-per", str(a.refl_per), "--child", mode], env={**os.environ, **env})
