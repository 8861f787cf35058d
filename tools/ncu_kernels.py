#!/usr/bin/env python
"""ncu_kernels.py -- per-launch duration, DRAM bytes and achieved DRAM bandwidth of every kernel in an ncu report.

    python tools/ncu_kernels.py gpurun_out/r02_small_kernels_dw.ncu-rep [--peak 6551]

Used for the memory-bound kernels of the step (per-reflection chain, norms, Adam): achieved = (dram read + write) / duration."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    peak = float(sys.argv[sys.argv.index("--peak") + 1]) if "--peak" in sys.argv else 6551.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    col = {k: i for i, k in enumerate(h)}

    def val(r, k, scale):
        v = float(r[col[k]].replace(",", ""))
        return v * scale.get(u[col[k]].lower(), 1)
    print(f"{'kernel':44s} {'us':>8s} {'dram MB':>9s} {'GB/s':>7s} {'of HBM peak':>11s} {'issue %':>8s}")
    for r in rows[2:]:
        t = val(r, "gpu__time_duration.sum", {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6})
        by = val(r, "dram__bytes_read.sum", {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}) + \
             val(r, "dram__bytes_write.sum", {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9})
        iss = r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]] if "smsp__issue_active.avg.pct_of_peak_sustained_active" in col else "?"
        print(f"{r[col['Kernel Name']][:44]:44s} {t:8.1f} {by / 1e6:9.1f} {by / 1e3 / t:7.0f} {by / 1e3 / t / peak:11.2f} {iss:>8s}")


if __name__ == "__main__":
    main()
