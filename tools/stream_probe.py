#!/usr/bin/env python
"""f-4 numbers: ONE long stream (the reference's urls.10K fixture, and a 64 MiB text stream) through
  device   csnappy_stream_decompress, stream resident in HBM, CUDA events
  drop-in  csnappy_decompress on host pointers (pageable memory, H2D + D2H inside)
  serial   the warp-per-stream path of round 1 (stream_decode_min = -1)
  cpu      the unmodified reference on one host core
and the snappy_unittest-style adapter line (tools/snappy_unittest_csnappy.cc)."""
import gzip
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csnappy_b200 as cs
import oracle


def fixture(name):
    with gzip.open(os.path.join(ROOT, "tests", "golden", name + ".gz"), "rb") as f:
        return f.read()


def timed_host(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def probe(label, data, comp):
    hlen, n = cs.csnappy_get_uncompressed_length(comp)
    raw = comp[hlen:]
    d_src = torch.from_numpy(np.frombuffer(raw, dtype=np.uint8).copy()).cuda()
    ws = torch.empty(cs.api.lib().csnappy_stream_decompress_workspace(len(raw), n), dtype=torch.uint8, device="cuda")
    out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    res = torch.zeros(2, dtype=torch.int32, device="cuda")
    for _ in range(3):
        cs.api.stream_decompress(d_src, len(raw), n, out=out, workspace=ws, result=res)
    torch.cuda.synchronize()
    assert res.tolist() == [n, 0] and out[:n].cpu().numpy().tobytes() == data
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        cs.api.stream_decompress(d_src, len(raw), n, out=out, workspace=ws, result=res)
    e1.record()
    torch.cuda.synchronize()
    dev_s = e0.elapsed_time(e1) * 1e-3 / reps
    drop_s = timed_host(lambda: cs.csnappy_decompress(comp, n), 10)
    cs.set_tuning("stream_decode_min", -1)
    serial_s = timed_host(lambda: cs.csnappy_decompress(comp, n), 2)
    cs.set_tuning("stream_decode_min", 0)
    ref = oracle.best()
    cpu_s = timed_host(lambda: ref.decompress(comp, n), 5)
    comp_s = timed_host(lambda: cs.csnappy_compress(data, 15), 10)
    cpu_comp_s = timed_host(lambda: ref.compress(data, 15), 3)
    print(f"{label}: {len(comp)} -> {n} bytes | decode GB/s: device {n / dev_s / 1e9:.2f}  drop-in {n / drop_s / 1e9:.3f}  "
          f"serial-path drop-in {n / serial_s / 1e9:.4f}  cpu({ref.kind}, 1 core) {n / cpu_s / 1e9:.3f} | "
          f"csnappy_compress GB/s: drop-in {n / comp_s / 1e9:.3f}  cpu 1 core {n / cpu_comp_s / 1e9:.3f}", flush=True)


urls = fixture("urls.10K")
probe("urls.10K", urls, fixture("urls.10K.snappy"))
big = (urls * 96)[: 64 << 20]
probe("64 MiB text", big, oracle.best().compress(big, 15))
exe = "/tmp/snappy_unittest_csnappy"
subprocess.run(["g++", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "snappy_unittest_csnappy.cc"),
                "-o", exe, "-L" + os.path.join(ROOT, "csnappy_b200"), "-lcsnappy_b200",
                "-Wl,-rpath," + os.path.join(ROOT, "csnappy_b200")], check=True)
for name in ("urls.10K",):
    p = f"/tmp/{name}"
    open(p, "wb").write(fixture(name))
    open(p + ".snappy15", "wb").write(fixture("urls.10K.snappy"))
    subprocess.run([exe, "--wm", "15", "--expect", p + ".snappy15", p], check=True)
    subprocess.run([exe, p], check=True)
