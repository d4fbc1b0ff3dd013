#!/usr/bin/env python
"""Run bench.py over tuning knobs on the GPU box and print one line per setting.

    python tools/sweep.py --pages 262144 --lanes-c 8,16,32 --lanes-d 8,16,32
"""
import argparse
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--pages", type=int, default=262144)
ap.add_argument("--lanes-c", default="16")
ap.add_argument("--lanes-d", default="32")
ap.add_argument("--ctas", default="0")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--only", default="")
ap.add_argument("--text", default="words")
ap.add_argument("--stage-d", default="0")
ap.add_argument("--smem-d", default="0")
args = ap.parse_args()
for lc, ld, ct, sd, kb in itertools.product(args.lanes_c.split(","), args.lanes_d.split(","), args.ctas.split(","),
                                         args.stage_d.split(","), args.smem_d.split(",")):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--pages", str(args.pages), "--steps", str(args.steps),
           "--warmup", "3", "--no-e2e", "--no-cpu", "--lanes-c", lc, "--lanes-d", ld, "--ctas-per-sm", ct, "--stage-d", sd, "--smem-d", kb] + (["--only", args.only] if args.only else []) + ["--text", args.text]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"lanes_c={lc} lanes_d={ld} ctas={ct} stage_d={sd} smem_d={kb}: compress {d['compress_gbs']} GB/s  decompress {d['decompress_gbs']} GB/s  "
              f"value {d['value']}", flush=True)
    except Exception:
        print(f"lanes_c={lc} lanes_d={ld} ctas={ct}: FAILED rc={r.returncode}\n{r.stdout[-500:]}\n{r.stderr[-1500:]}", flush=True)
