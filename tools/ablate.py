#!/usr/bin/env python
"""Ablation timings of k_obs_tc2: what each part of the observation kernel costs on the wall clock.

    python tools/ablate.py --build        (here: nvcc, no GPU needed)
    python tools/ablate.py                (on the GPU box)

Each variant is the product library compiled with one -DCLB_ABL_* switch that removes one piece of work (the RESULTS
ARE WRONG; only the time is meaningful): CHAIN = no forward / dX tcgen05.mma, DW = no dW tcgen05.mma, STS = no
shared-memory stores of the dW operand images, SCR = no activation scratch traffic, PART = no FP64 partial
read-modify-write.  Measured on B200 (10 M obs, MLP 32x20): baseline 23.7 ms; CHAIN 22.7; DW 22.6; STS 18.7; SCR 21.0;
PART 21.2 -> the load/store + shared-memory pipe (ncu: "Mem Busy" 59 %), not the tensor pipe or instruction issue,
is what the kernel waits for.
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = ["CHAIN", "DW", "STS", "SCR", "PART"]


def lib(v):
    return os.path.join(ROOT, "tools", f"libclb_abl_{v}.so")


if "--build" in sys.argv:
    for v in VARIANTS:
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                        f"-DCLB_ABL_{v}", "-I", os.path.join(ROOT, "include"), "-o", lib(v),
                        os.path.join(ROOT, "careless_b200", "csrc", "clb_api.cu")], check=True)
    sys.exit(0)
for v in [None] + VARIANTS:
    env = dict(os.environ)
    if v:
        env["CLB_LIB_PATH"] = lib(v)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_configs.py"), "--which", "mono", "--steps", "8"],
                         env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    print(v or "baseline", round(json.loads(out)["ms_per_step"], 2), "ms/step")
