#!/usr/bin/env python
"""Per-SASS-instruction executed counts and stall-sample shares from an .ncu-rep source page.

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep <units (e.g. pages in the launch)> [min_count_per_unit] [min_pct]
"""
import csv
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
minc = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data)
ts = sum(int(r[isamp]) for r in data)
print(f"total warp-instructions {tot} ({tot / units:.0f} per unit), samples {ts}, SASS lines {len(data)}")
for i, r in enumerate(data):
    n = int(r[ia]) / units
    pct = int(r[isamp]) * 100 / ts
    if n >= minc or pct >= minp:
        print(f"{i:4d} {n:8.1f} {pct:5.2f}%  {r[isrc].strip()[:100]}")
