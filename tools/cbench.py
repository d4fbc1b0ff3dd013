#!/usr/bin/env python
"""Compress-only timing of one page class (kernel experiments; output is NOT checked -- use the tests for that).
    python tools/cbench.py [text|zero|random|mixed] [pages] [unit] [wm]"""
import sys

import torch

sys.path.insert(0, ".")
import csnappy_b200 as cs
from csnappy_b200 import synth

only = sys.argv[1] if len(sys.argv) > 1 else "text"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
L = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
wm = int(sys.argv[4]) if len(sys.argv) > 4 else 13
stage = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # compress_stage_input: 0 auto, 1 staged, 2 read from global
cs.set_tuning("compress_stage_input", stage)
if L > 4096:
    d = synth.text_fragments(n, L, device="cuda")
else:
    d = synth.mixed_pages(n, L, device="cuda", text="urls", only="" if only == "mixed" else only)
ostride = cs.api.out_stride_for(L)
out = torch.empty(n * ostride, dtype=torch.uint8, device="cuda")
olen = torch.empty(n, dtype=torch.int32, device="cuda")
f = lambda: cs.batch_compress_fragments(d, L, n, wm, out=out, out_len=olen, out_stride=ostride)
for _ in range(3):
    f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    f()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{only} unit {L} wm {wm} n {n} stage {stage}: compress {n * L / ms / 1e6:.1f} GB/s  ({ms:.3f} ms)  ratio {float(olen.sum()) / (n * L):.4f}", flush=True)
