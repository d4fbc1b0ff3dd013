#!/usr/bin/env python
"""One decode shape for profiling: python tools/lane_probe.py <block> <wm> <n_blocks> <stage> [reps]"""
import sys

import torch

sys.path.insert(0, ".")
import csnappy_b200 as cs
from csnappy_b200 import synth

L, wm, n, stage = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
only = sys.argv[6] if len(sys.argv) > 6 else ""
ctas = int(sys.argv[7]) if len(sys.argv) > 7 else 0
d = synth.text_fragments(n, L, device="cuda") if L > 4096 else synth.mixed_pages(n, L, device="cuda", text="urls", only=only)
ostride = cs.api.out_stride_for(L)
out, olen = cs.batch_compress_fragments(d, L, n, wm)
back = torch.empty(n * L, dtype=torch.uint8, device="cuda")
blen = torch.empty(n, dtype=torch.int32, device="cuda")
st = torch.empty(n, dtype=torch.int32, device="cuda")
cs.set_tuning("decompress_stage_input", stage)
cs.set_tuning("decompress_lane_warps", ctas)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    e0.record()
    cs.batch_decompress(out, olen, n, L, in_stride=ostride, out=back, out_stride=L, out_len=blen, status=st)
    e1.record()
    torch.cuda.synchronize()
    print(f"block {L} n {n} stage {stage} lane_warps {ctas}: decompress {n * L / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s", flush=True)
assert int((st != 0).sum()) == 0 and torch.equal(back, d)
