"""csnappy_b200 -- B200-native batched Snappy block codec behind the csnappy.h API.

The product is csnappy_b200/libcsnappy_b200.so (plain-C shim + sm_100a CUDA kernels,
sources under csnappy_b200/csrc, public ABI in include/).  This package is the thin
Python host side over that C-ABI: `api` mirrors the reference interface, `shard`
partitions batches across ranks, `synth` makes the benchmark workloads.
"""
from . import api  # noqa: F401
from .api import *  # noqa: F401,F403
