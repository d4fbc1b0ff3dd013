#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import csnappy_b200 as cs
from cases import fuzz_pages
from csnappy_b200 import synth

B, L = 96, 4096
pages = synth.mixed_pages(B, L, seed=0x5EED0001, device="cuda", pool_bytes=1 << 20)
extra = torch.from_numpy(np.frombuffer(b"".join(fuzz_pages(5, 28, L)), dtype=np.uint8).copy()).cuda()
pages = torch.cat([pages, extra])
B = pages.numel() // L
for lanes in (32, 16, 8):
    cs.set_tuning("compress_lanes", lanes)
    cs.set_tuning("decompress_lanes", lanes)
    out, out_len = cs.batch_compress_fragments(pages, L, B, 13)
    ostride = cs.api.out_stride_for(L)
    back, back_len, status = cs.batch_decompress(out, out_len, B, L, in_stride=ostride)
    torch.cuda.synchronize()
    assert int((status != 0).sum()) == 0 and torch.equal(back.view(B, L), pages.view(B, L)), lanes
    packed, off = cs.batch_pack(out, ostride, out_len, B)
    back2, _, st2 = cs.batch_decompress(packed, out_len, B, L, in_off=off[:-1].contiguous())
    torch.cuda.synchronize()
    assert int((st2 != 0).sum()) == 0 and torch.equal(back2.view(B, L), pages.view(B, L)), lanes
# compress reading the block from global memory (the form large fragments take in large batches)
cs.set_tuning("compress_lanes", 0)
cs.set_tuning("compress_stage_input", 2)
out_u, out_len_u = cs.batch_compress_fragments(pages, L, B, 13)
cs.set_tuning("compress_stage_input", 0)
out_s, out_len_s = cs.batch_compress_fragments(pages, L, B, 13)
torch.cuda.synchronize()
assert torch.equal(out_len_u, out_len_s)
# the other decoder families on the same batch: 3 = warp per block against global memory, 4 = lane per block
cs.set_tuning("compress_lanes", 0)
cs.set_tuning("decompress_lanes", 0)
out, out_len = cs.batch_compress_fragments(pages, L, B, 13)
for stage in (3, 4, 2, 1):
    cs.set_tuning("decompress_stage_input", stage)
    back, back_len, status = cs.batch_decompress(out, out_len, B, L, in_stride=cs.api.out_stride_for(L))
    torch.cuda.synchronize()
    assert int((status != 0).sum()) == 0 and torch.equal(back.view(B, L), pages.view(B, L)), stage
cs.set_tuning("decompress_stage_input", 0)
# 32 KiB fragments
frag = synth.text_fragments(6, 32768, device="cuda", pool_bytes=1 << 20)
o, ol = cs.batch_compress_fragments(frag, 32768, 6, 15)
b, bl, st = cs.batch_decompress(o, ol, 6, 32768, in_stride=cs.api.out_stride_for(32768))
torch.cuda.synchronize()
assert int((st != 0).sum()) == 0 and torch.equal(b.view(6, 32768), frag.view(6, 32768))
# whole multi-chunk streams through the drop-in API (global decode path) and the page container
import gzip

urls = gzip.open("csnappy_b200/data/urls.10K.gz").read()[:200000]
c = cs.csnappy_compress(urls, 15)
rc, outb = cs.csnappy_decompress(c, len(urls))
assert rc == 0 and outb == urls
h_in = np.frombuffer(urls, dtype=np.uint8).copy()
cont = np.zeros(cs.api.bc_max_container_length(len(urls), 4096), dtype=np.uint8)
clen = cs.api.bc_compress_host(h_in, len(urls), cont, 13, 4096)
outp = np.zeros((len(urls) + 4095) // 4096 * 4096, dtype=np.uint8)
rc, olen, _ = cs.api.bc_decompress_host(cont, clen, outp, 4096)
assert rc == 0 and olen == len(urls) and outp[:olen].tobytes() == urls
bad = bytes.fromhex("086162630104")
assert cs.csnappy_decompress_noheader(bad, 100)[0] == -5
print("sanitize_run ok")
