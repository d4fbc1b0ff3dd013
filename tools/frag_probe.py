#!/usr/bin/env python
"""BASELINE.json configs[2]/[3] as a diagnostic: 32 KiB text fragments, wm 15 / 16 (and 4 KiB for comparison).
    python tools/frag_probe.py [decompress_stage_input]   (1: always stage, 0: default routing)"""
import sys

import torch

sys.path.insert(0, ".")
import csnappy_b200 as cs
from csnappy_b200 import synth


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1  # multiplies the number of blocks (1: 0.5 GiB per shape)
cs.set_tuning("decompress_stage_input", stage)
print("decompress_stage_input", stage)
for L, wm, n in ((32768, 15, 16384 * scale), (32768, 16, 16384 * scale), (16384, 14, 32768 * scale), (4096, 13, 131072 * scale)):
    d = synth.text_fragments(n, L, device="cuda")
    ostride = cs.api.out_stride_for(L)
    out = torch.empty(n * ostride, dtype=torch.uint8, device="cuda")
    olen = torch.empty(n, dtype=torch.int32, device="cuda")
    back = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    blen = torch.empty(n, dtype=torch.int32, device="cuda")
    st = torch.empty(n, dtype=torch.int32, device="cuda")
    tc = timed(lambda: cs.batch_compress_fragments(d, L, n, wm, out=out, out_len=olen, out_stride=ostride))
    td = timed(lambda: cs.batch_decompress(out, olen, n, L, in_stride=ostride, out=back, out_stride=L, out_len=blen, status=st))
    assert int((st != 0).sum()) == 0 and torch.equal(back, d)
    print(f"block {L} wm {wm}: ratio {float(olen.sum()) / (n * L):.3f}  compress {n * L / tc / 1e9:.1f} GB/s  decompress {n * L / td / 1e9:.1f} GB/s",
          flush=True)
