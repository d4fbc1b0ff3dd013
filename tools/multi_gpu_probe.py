#!/usr/bin/env python
"""One process, all GPUs of the box (VERDICT r1 task 3):

  1. copy-only ceiling: G = 1, 2, 4, 8 devices moving pinned host memory H2D and D2H concurrently
     (what any host-buffer codec call on this box is bounded by), with the NUMA node of every GPU and CPU set;
  2. strong scaling of ONE pinned host buffer through csnappy_bc_compress_host_multi /
     csnappy_bc_decompress_host_multi on 1, 2, 4, 8 devices (same bytes as the single-device call).

    python tools/multi_gpu_probe.py [GiB of pages, default 8]
"""
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, ".")
import csnappy_b200 as cs
from csnappy_b200 import synth

PAGE = 4096
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
ndev = torch.cuda.device_count()
print(f"devices {ndev}  host cores {len(os.sched_getaffinity(0))}", flush=True)
for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "cat /sys/devices/system/node/node*/cpulist 2>/dev/null | head -8"]):
    try:
        print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout.strip(), flush=True)
    except Exception as e:  # noqa: BLE001
        print("n/a:", e)

# ---- 1. copy-only ceiling -----------------------------------------------------------------------------------
N = 1 << 30
bufs = []
for d in range(ndev):
    with torch.cuda.device(d):
        bufs.append((torch.empty(N, dtype=torch.uint8, pin_memory=True), torch.empty(N, dtype=torch.uint8, pin_memory=True),
                     torch.empty(N, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(N, dtype=torch.uint8, device=f"cuda:{d}"),
                     torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))


def copy_round(G, h2d=True, d2h=True, reps=4):
    for d in range(G):
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        for d in range(G):
            hi, ho, di, do, s1, s2 = bufs[d]
            if h2d:
                with torch.cuda.stream(s1):
                    di.copy_(hi, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    ho.copy_(do, non_blocking=True)
    for d in range(G):
        torch.cuda.synchronize(d)
    return N * reps * G / (time.perf_counter() - t0) / 1e9


G = 1
while G <= ndev:
    copy_round(G, reps=1)
    print(f"copy-only G={G}: H2D alone {copy_round(G, True, False):.1f} GB/s  D2H alone {copy_round(G, False, True):.1f} GB/s  "
          f"both directions at once {copy_round(G):.1f} GB/s per direction (aggregate over the {G} devices)", flush=True)
    G *= 2
del bufs
torch.cuda.empty_cache()

# ---- 2. strong scaling of one host buffer through the multi-device container calls ---------------------------------
B = int(gib * (1 << 30)) // PAGE
h_in = torch.empty(B * PAGE, dtype=torch.uint8, pin_memory=True)
step = 1 << 18
for s in range(0, B, step):
    n = min(step, B - s)
    h_in[s * PAGE:(s + n) * PAGE].copy_(synth.mixed_pages(n, PAGE, seed=0x5EED0001, device="cuda:0", first_page=s, text="urls"))
torch.cuda.synchronize()
h_cont = torch.empty(cs.api.bc_max_container_length(B * PAGE, PAGE), dtype=torch.uint8, pin_memory=True)
h_back = torch.empty(B * PAGE, dtype=torch.uint8, pin_memory=True)
ref_len = None
G = 1
while G <= ndev:
    devs = list(range(G))
    best_c = best_d = 1e9
    for it in range(3):
        t0 = time.perf_counter()
        clen = cs.api.bc_compress_host_multi(h_in, B * PAGE, h_cont, 13, PAGE, devices=devs)
        t1 = time.perf_counter()
        rc, olen, _ = cs.api.bc_decompress_host_multi(h_cont, clen, h_back, PAGE, devices=devs)
        t2 = time.perf_counter()
        assert rc == 0 and olen == B * PAGE
        if it:
            best_c, best_d = min(best_c, t1 - t0), min(best_d, t2 - t1)
    if ref_len is None:
        ref_len = clen
        assert torch.equal(h_back, h_in)
    assert clen == ref_len, "container length differs between device counts"
    print(f"strong scaling, {gib:g} GiB of pages, G={G}: compress {B * PAGE / best_c / 1e9:.1f} GB/s  decompress "
          f"{B * PAGE / best_d / 1e9:.1f} GB/s  e2e {2 * B * PAGE / (best_c + best_d) / 1e9:.1f} GB/s  (container {clen} B)", flush=True)
    G *= 2
