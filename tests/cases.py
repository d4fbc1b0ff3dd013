"""Seeded input generators shared by the oracle tests and the GPU parity tests.

synth_cases() must stay identical to tests/golden/make_golden.py::synth_cases
(golden.json records the reference's output for every entry)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import synth_cases  # noqa: E402,F401


def fuzz_pages(seed: int, count: int, size: int):
    """Mixed-entropy pages: random, zero, periodic, low-entropy, word text, copies-with-noise."""
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 10)), dtype=np.uint8)) for _ in range(300)]
    out = []
    for i in range(count):
        k = i % 7
        if k == 0:
            b = rng.integers(0, 256, size, dtype=np.uint8).tobytes()
        elif k == 1:
            b = bytes(size)
        elif k == 2:
            p = int(rng.integers(1, 80))
            pat = rng.integers(0, 256, p, dtype=np.uint8).tobytes()
            b = (pat * (size // p + 1))[:size]
        elif k == 3:
            b = rng.integers(0, int(rng.integers(2, 6)), size, dtype=np.uint8).tobytes()
        elif k == 4:
            b = b" ".join(words[int(j)] for j in rng.integers(0, 300, size // 2 + 8))[:size]
        elif k == 5:
            base = bytearray(b" ".join(words[int(j)] for j in rng.integers(0, 40, size // 2 + 8))[:size])
            for j in rng.integers(0, max(size, 1), size // 50):
                base[int(j)] = int(rng.integers(0, 256))
            b = bytes(base)
        else:
            # long matches at long offsets + runs
            half = rng.integers(0, 256, max(size // 3, 1), dtype=np.uint8).tobytes()
            b = (half + bytes(int(rng.integers(0, 200))) + half + half[::-1] + half)[:size]
            b = b + bytes(size - len(b))
        out.append(b)
    return out
