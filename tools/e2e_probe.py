#!/usr/bin/env python
"""Time the host-buffer container calls separately at several batch sizes (diagnostic)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import csnappy_b200 as cs
from csnappy_b200 import synth

PAGE = 4096
import os
if os.environ.get("CHUNK_MB"):
    cs.set_tuning("chunk_mb", int(os.environ["CHUNK_MB"]))
for B in [int(x) for x in (sys.argv[1:] or ["262144", "524288", "1048576"])]:
    pages = synth.mixed_pages(B, PAGE, seed=0x5EED0001, device="cuda")
    h_in = torch.empty(B * PAGE, dtype=torch.uint8, pin_memory=True)
    h_cont = torch.empty(cs.api.bc_max_container_length(B * PAGE, PAGE), dtype=torch.uint8, pin_memory=True)
    h_back = torch.empty(B * PAGE, dtype=torch.uint8, pin_memory=True)
    h_in.copy_(pages)
    torch.cuda.synchronize()
    for it in range(3):
        t0 = time.perf_counter()
        clen = cs.api.bc_compress_host(h_in, B * PAGE, h_cont, 13, PAGE)
        t1 = time.perf_counter()
        rc, olen, _ = cs.api.bc_decompress_host(h_cont, clen, h_back, PAGE)
        t2 = time.perf_counter()
        print(f"B={B} it={it}: compress {1e3 * (t1 - t0):.1f} ms ({B * PAGE / (t1 - t0) / 1e9:.1f} GB/s)  "
              f"decompress {1e3 * (t2 - t1):.1f} ms ({B * PAGE / (t2 - t1) / 1e9:.1f} GB/s)", flush=True)
    del pages, h_in, h_cont, h_back
    torch.cuda.empty_cache()
