"""Synthetic workloads of BASELINE.json's configs.

zram-style batch: pages of `page_len` bytes, class per page from
splitmix64(seed ^ page_index): 50 % text, 25 % zero, 25 % random.
  text    text="urls": a page_len slice of the reference's urls.10K corpus (committed copy:
          csnappy_b200/data/urls.10K.gz) at offset h mod (702087 - page_len) -- the primary definition
          of SURVEY.md 8d config 2;  text="words": the purely synthetic alternative, a page_len
          slice of a Zipf(1.1) word stream over a 4096-word lowercase vocabulary
  zero    all zero bytes
  random  uniform bytes
Nothing reads /root/reference: the urls corpus comes from the committed fixture.
Generation uses torch only as device-memory plumbing (gather / randint); it is not
part of the measured path.
"""
from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(_MASK)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(_MASK)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(_MASK)
    return z ^ (z >> np.uint64(31))


def text_pool(nbytes: int, seed: int = 0x5EED0002) -> np.ndarray:
    """Zipf(1.1) words over a 4096-word lowercase vocabulary -> uint8 array of nbytes."""
    rng = np.random.default_rng(seed)
    V = 4096
    wlen = rng.integers(2, 11, V)
    letters = rng.integers(97, 123, (V, 10), dtype=np.uint8)
    p = 1.0 / np.arange(1, V + 1) ** 1.1
    p /= p.sum()
    n_words = int(nbytes / 5.0) + 1024
    ranks = rng.choice(V, size=n_words, p=p)
    seps = np.where(rng.random(n_words) < 0.08, 10, 32).astype(np.uint8)
    lens = wlen[ranks] + 1
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    total = int(lens.sum())
    out = np.empty(total, dtype=np.uint8)
    for j in range(10):
        m = wlen[ranks] > j
        out[starts[m] + j] = letters[ranks[m], j]
    out[starts + wlen[ranks]] = seps
    assert total >= nbytes, (total, nbytes)
    return out[:nbytes]


def urls_pool() -> np.ndarray:
    """The reference's own text corpus (testdata/urls.10K, committed as csnappy_b200/data/urls.10K.gz): the
    primary text class of SURVEY.md 8d config 2 is a 4096-byte slice of it at a seeded offset."""
    import gzip
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "urls.10K.gz")
    with gzip.open(path, "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8).copy()


def page_classes(n_pages: int, seed: int, first_page: int = 0):
    """-> (cls uint8 [n]: 0 text, 1 zero, 2 random ; h uint64 [n])."""
    with np.errstate(over="ignore"):
        idx = np.arange(first_page, first_page + n_pages, dtype=np.uint64)
        h = splitmix64(idx ^ np.uint64(seed))
    q = (h & np.uint64(3)).astype(np.uint8)
    cls = np.where(q < 2, 0, q - 1).astype(np.uint8)
    return cls, h


def mixed_pages(n_pages: int, page_len: int = 4096, seed: int = 0x5EED0001, device="cuda", first_page: int = 0,
                pool_bytes: int = 8 << 20, text_only: bool = False, only: str = "", text: str = "words"):
    """uint8 tensor [n_pages * page_len] on `device` holding the zram-style mixed batch.
    `first_page` lets each rank of a sharded run generate exactly its slice of the global batch."""
    import torch

    cls, h = page_classes(n_pages, seed, first_page)
    if text_only or only == "text":
        cls[:] = 0
    elif only:
        cls[:] = {"zero": 1, "random": 2}[only]
    if text == "urls":
        host_pool = urls_pool()
        pool_bytes = len(host_pool)
    else:
        host_pool = text_pool(pool_bytes)
    pool = torch.from_numpy(host_pool).to(device)
    windows = pool.unfold(0, page_len, 1)  # [pool - page_len + 1, page_len] overlapping view
    pages = torch.zeros((n_pages, page_len), dtype=torch.uint8, device=device)
    text_idx = np.nonzero(cls == 0)[0]
    rand_idx = np.nonzero(cls == 2)[0]
    offs = ((h[text_idx] >> np.uint64(8)) % np.uint64(pool_bytes - page_len)).astype(np.int64)
    step = 1 << 16
    for s in range(0, len(text_idx), step):
        ti = torch.from_numpy(text_idx[s:s + step]).to(device)
        to = torch.from_numpy(offs[s:s + step]).to(device)
        pages[ti] = windows.index_select(0, to)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) ^ (first_page * 0x9E37 + 1))
    for s in range(0, len(rand_idx), step):
        ri = torch.from_numpy(rand_idx[s:s + step]).to(device)
        pages[ri] = torch.randint(0, 256, (len(ri), page_len), dtype=torch.uint8, device=device, generator=gen)
    return pages.view(-1)


_POOLS: dict = {}


def text_fragments(n_frag: int, frag_len: int = 32768, seed: int = 0x5EED0002, device="cuda", first: int = 0,
                   pool_bytes: int = 32 << 20):
    """Config 3: text-like 32 KiB fragments (slices of the word stream at seeded offsets; fragment i of the global
    stream is the same whatever rank generates it: `first` is the rank's first fragment)."""
    import torch

    with np.errstate(over="ignore"):
        h = splitmix64(np.arange(first, first + n_frag, dtype=np.uint64) ^ np.uint64(seed))
    key = (pool_bytes, str(device))
    if key not in _POOLS:  # the pool itself is fixed (seed 0x5EED0002); `seed` picks the offsets
        _POOLS[key] = torch.from_numpy(text_pool(pool_bytes)).to(device)
    pool = _POOLS[key]
    windows = pool.unfold(0, frag_len, 1)
    offs = torch.from_numpy((h % np.uint64(pool_bytes - frag_len)).astype(np.int64)).to(device)
    out = torch.empty((n_frag, frag_len), dtype=torch.uint8, device=device)
    step = 1 << 13
    for s in range(0, n_frag, step):
        out[s:s + step] = windows.index_select(0, offs[s:s + step])
    return out.view(-1)
