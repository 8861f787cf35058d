#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters of one kernel to source lines.

    python tools/sass_lines.py <ncu source-page csv (--print-source sass)> <lib.so> <mangled kernel name> [top]

The cubin inside the .so must be the build that was profiled.  Offsets are matched by instruction order.
"""
import collections, csv, re, subprocess, sys, tempfile, os

csv_path, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = (".text." + kern + " ") in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
rows = list(csv.reader(open(csv_path)))
hdr, data = rows[1], rows[2:]
iS, iI, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [r for r in data if len(r) > iW and r[0].startswith("0x")]
if len(data) != len(lines):
    print(f"warning: {len(data)} profiled instructions vs {len(lines)} in the cubin", file=sys.stderr)
inst, samp = collections.Counter(), collections.Counter()
for (off, loc, txt), r in zip(lines, data):
    inst[loc] += int(r[iI]); samp[loc] += int(r[iW])
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}, stall samples {ts}")
print("by instructions:")
for loc, n in inst.most_common(top):
    print(f"  {str(loc):36s} inst {100*n/ti:6.2f}%   samples {100*samp[loc]/ts:6.2f}%")
print("by samples:")
for loc, n in samp.most_common(top):
    print(f"  {str(loc):36s} samples {100*n/ts:6.2f}%   inst {100*inst[loc]/ti:6.2f}%")
