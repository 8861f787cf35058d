#!/usr/bin/env python
"""Per-phase cycle accounting of k_obs<32, studentt, TC> (debug build with -DCLB_PHASE_TIMING).

    python tools/phase_times.py [--obs N]

Builds tools/libclb_phases.so, runs a few steps of BASELINE configs[1] and prints, per phase, the share of the
cycles thread 0 of every CTA spent there (two CTAs share an SM, so a phase also absorbs the other CTA's work).
"""
import argparse, ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "libclb_phases.so")
NAMES = {0: "fwd: bias+leaky+scratch store of prev layer", 1: "fwd: issue (split, STTM, W image, sync, MMA issue)", 2: "fwd: wait + LDTM",
         3: "fwd: tail", 4: "epilogue (likelihood, dz reduction)", 5: "bwd: pre-layer ALU (mask, act loads, W prefetch)",
         6: "bwd: bias_partial shuffles", 7: "bwd: issue_backward (split x2, STTM, 32 STS.128, sync, MMA issue)",
         8: "bwd: wait dX + LDTM", 9: "bwd: wait dW + LDTM x2 + fold + STS", 10: "bwd: __syncthreads after collect_dw",
         11: "bwd: stage read + FP64 partial RMW + bias sum", 12: "bwd: final __syncthreads"}


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--obs", type=int, default=2_000_000); ap.add_argument("--build-only", action="store_true")
    args = ap.parse_args()
    if not os.path.exists(OUT) or args.build_only:
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                        "-DCLB_PHASE_TIMING", "-I", os.path.join(ROOT, "include"), "-o", OUT,
                        os.path.join(ROOT, "careless_b200", "csrc", "clb_api.cu")], check=True)
    if args.build_only:
        return
    os.environ["CLB_LIB_PATH"] = OUT
    import numpy as np
    from careless_b200 import synth, _lib
    from careless_b200.engine import Engine, EngineConfig
    N, R = args.obs, args.obs // 20
    p = synth.make_mono(N, R, d=5, n_images=1000, seed=1)
    eng = Engine(EngineConfig(n_refl=R, n_meta=5, mlp_width=32, mlp_layers=20, likelihood="studentt", dof=12.0))
    eng.set_observations(p["refl_id"], None, p["metadata"], p["intensities"], p["uncertainties"])
    eng.set_prior(p["centric"], p["multiplicity"])
    lib = _lib.load()
    buf = (C.c_ulonglong * 32)()
    eng.step(3)
    lib.clb_debug_phases.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.clb_debug_phases(buf)
    eng.step(3)
    eng.synchronize()
    lib.clb_debug_phases(buf)
    v = np.array(list(buf), dtype=np.float64)
    tot = v.sum()
    tiles = (N + 127) // 128 * 3
    print(f"total {tot:.3e} cycles over {tiles} tile passes = {tot / tiles:.0f} cycles per tile per CTA ({tot / tiles / 20:.0f} per layer)")
    for i in range(13):
        per_layer = v[i] / tiles / 20
        print(f"  phase {i:2d} {100 * v[i] / tot:6.2f}%  {per_layer:8.0f} cycles/layer   {NAMES[i]}")
    eng.close()


if __name__ == "__main__":
    main()
