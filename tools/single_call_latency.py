#!/usr/bin/env python
"""Latency of the per-call drop-in entry points on one 4 KiB page (host pointers)."""
import sys
import time

sys.path.insert(0, ".")
import gzip

import csnappy_b200 as cs

page = gzip.open("csnappy_b200/data/urls.10K.gz").read()[:4096]
comp = cs.csnappy_compress_fragment(page, 13)
for name, fn in (("csnappy_compress_fragment(4 KiB)", lambda: cs.csnappy_compress_fragment(page, 13)),
                 ("csnappy_decompress_noheader(4 KiB)", lambda: cs.csnappy_decompress_noheader(comp, 4096))):
    for _ in range(50):
        fn()
    t0 = time.perf_counter()
    for _ in range(1000):
        fn()
    print(f"{name}: {1e3 * (time.perf_counter() - t0):.1f} us per call")
