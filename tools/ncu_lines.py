#!/usr/bin/env python
"""Executed warp-instructions and stall samples of one kernel PER SOURCE LINE of the kernel body.

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <cubin name part> <kernel symbol part> <units> [src file]

The ncu source page lists SASS in address order; `nvdisasm -gi` of the same cubin gives, for every instruction, the
inline chain.  Instructions are attributed to the OUTERMOST frame (the line of the kernel function that, through
inlining, produced them), so helpers like lds32u_a are charged to their call sites.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, so, cubin_part, sym_part, units = sys.argv[1:6]
units = float(units)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if cubin_part in f and "sm_100a" in f][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()

insts = []  # (outer file:line, text)
in_fn = False
chain = []
pending = []
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        in_fn = sym_part in m.group(1)
        continue
    if not in_fn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)( inlined at)?', ln)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2)), bool(m.group(3))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        if pending:
            chain = pending
            pending = []
        outer = chain[-1] if chain else ("?", 0, False)
        inner = chain[0] if chain else ("?", 0, False)
        insts.append((f"{outer[0]}:{outer[1]}", f"{inner[0]}:{inner[1]}", m.group(2).strip()))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
assert len(data) == len(insts), (len(data), len(insts))
tot = sum(int(r[ia]) for r in data)
ts = sum(int(r[isamp]) for r in data)
agg = collections.OrderedDict()
for (outer, inner, txt), r in zip(insts, data):
    a = agg.setdefault(outer, [0, 0, 0])
    a[0] += int(r[ia])
    a[1] += int(r[isamp])
    a[2] += 1
print(f"total warp-instructions {tot} ({tot / units:.0f} per unit), stall samples {ts}, SASS lines {len(data)}")
print("line                       instr/unit  %instr  %samples  sass")
for k in sorted(agg, key=lambda k: (k.split(":")[0], int(k.split(":")[1]))):
    a = agg[k]
    if a[0] / units >= 0.5 or a[1] * 100 / ts >= 0.3:
        print(f"{k:26s} {a[0] / units:10.1f} {a[0] * 100 / tot:6.2f}% {a[1] * 100 / ts:7.2f}%  {a[2]:4d}")
