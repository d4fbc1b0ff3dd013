#!/usr/bin/env python
"""Container calls on PAGEABLE caller memory for several copy-thread counts (one process per count: the pool reads
the knob when it starts).   python tools/pageable_probe.py <copy_threads> [pages]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import csnappy_b200 as cs
from csnappy_b200 import synth

PAGE = 4096
threads = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 19
cs.set_tuning("copy_threads", threads)
pages = synth.mixed_pages(B, PAGE, seed=0x5EED0001, device="cuda", text="urls")
h_in = pages.cpu()
h_cont = torch.empty(cs.api.bc_max_container_length(B * PAGE, PAGE), dtype=torch.uint8)
h_back = torch.empty(B * PAGE, dtype=torch.uint8)
assert not h_in.is_pinned()
best_c = best_d = 1e9
for it in range(4):
    t0 = time.perf_counter()
    clen = cs.api.bc_compress_host(h_in, B * PAGE, h_cont, 13, PAGE)
    t1 = time.perf_counter()
    rc, olen, _ = cs.api.bc_decompress_host(h_cont, clen, h_back, PAGE)
    t2 = time.perf_counter()
    assert rc == 0 and olen == B * PAGE
    if it:
        best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
assert torch.equal(h_back, h_in)
print(f"copy_threads {threads}: compress {B * PAGE / best_c / 1e9:.1f} GB/s  decompress {B * PAGE / best_d / 1e9:.1f} GB/s  "
      f"e2e {2 * B * PAGE / (best_c + best_d) / 1e9:.1f} GB/s (pageable buffers, {B} pages)", flush=True)
