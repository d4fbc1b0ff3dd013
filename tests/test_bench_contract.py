"""bench.py's one-line JSON contract: the reference arm runs on CPU here, the B200 arm on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def run_bench(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--pages", "2048")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "GB/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "pages" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"] and d["gpu_launches"] == 0
    # the same config keys as the B200 arm prints (the driver compares the two dicts)
    assert set(d["config"]) == {"workload", "pages_per_gpu", "page_bytes", "wm", "ratio", "parallelism", "l2", "value_definition"}
    assert d["config"]["pages_per_gpu"] == 2048 and 0.3 < d["config"]["ratio"] < 0.7


@pytest.mark.gpu
def test_b200_arm_line_on_gpu():
    d = run_bench("--pages", "32768", "--steps", "3", "--warmup", "3", "--frag-gib", "0.0625", "--decode-gib", "0.25",
                  "--wave-gib", "0.125")
    assert "impl" not in d and BASE_KEYS | {"roofline", "clocks", "compress_gbs", "decompress_gbs"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["gpu_launches"] == 6  # one compress + one decompress kernel per step
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    assert rf["traffic"] is None or rf["traffic"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 32768 * 4096 and e["d2h_bytes_per_step"] > 32768 * 4096
    assert e["value"] < d["value"]  # host buffers and PCIe inside the timed region
    assert 0 < d["e2e_pageable"]["value"] < d["value"] and 0 < d["e2e_concurrent"]["value"] < d["value"]
    # BASELINE.json configs[2] / [3] ride in the same line
    w = d["workloads"]
    for wm in ("wm15", "wm16"):
        f = w["fragments_32k"][wm]
        assert f["compress_gbs"] > 0 and f["decompress_gbs"] > 0 and 0 < f["roofline_frac_compress"] < 1
        assert f["units_checked_against_cpu_reference"] > 0 and 0.4 < f["ratio"] < 0.6
    for k in ("pages_4k", "fragments_32k"):
        x = w["decode_only"][k]
        assert x["waves"] == 2 and x["decompress_gbs"] > 0 and 0 < x["roofline_frac_decompress"] < 1
        assert x["bytes_decoded_per_gpu"] == 2 * x["units_per_wave"] * x["unit_bytes"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["alt_workload"]["value"] > 0
    # per-class ratios of the zram-style batch (text ~0.60, zero ~0.047, random ~1.001)
    pc = d["per_class"]
    assert 0.55 < pc["text"]["ratio"] < 0.65 and 0.04 < pc["zero"]["ratio"] < 0.055 and 1.0 < pc["random"]["ratio"] < 1.002
