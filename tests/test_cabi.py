"""The C-ABI boundary on a machine WITHOUT a GPU: the library loads, exports every symbol the
headers declare, does its pure-host arithmetic, and fails loudly (never falls back to a CPU
codec) when asked to compress or decompress."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", h) for h in ("csnappy.h", "csnappy_batch.h")]


@pytest.fixture(scope="module")
def lib():
    from csnappy_b200 import build

    build.build()
    from csnappy_b200._lib import lib as load

    return load()


def declared_functions():
    names = []
    for h in HEADERS:
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        text = re.sub(r"#.*", "", text)
        names += re.findall(r"\b(csnappy_\w+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(lib):
    from csnappy_b200._lib import LIB_PATH

    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    decl = declared_functions()
    assert len(decl) >= 17, decl
    missing = [n for n in decl if n not in exported]
    assert not missing, missing
    # the six drop-in symbols of the reference (csnappy.h:30-119) are all there
    for n in ("csnappy_compress", "csnappy_compress_fragment", "csnappy_decompress", "csnappy_decompress_noheader",
              "csnappy_get_uncompressed_length", "csnappy_max_compressed_length"):
        assert n in exported
    # and nothing of the oracle or of a CPU codec leaks into the product library
    assert not [s for s in exported if s.startswith("oracle_") or s.startswith("harness_")]
    # every ctypes signature in _lib.py resolves
    assert set(lib._signatures) >= set(decl)


def test_host_arithmetic_matches_reference_pins(lib, golden):
    # csnappy_compress.c:612-616 (SURVEY.md 8b pins)
    assert [lib.csnappy_max_compressed_length(n) for n in (0, 4096, 32768)] == [32, 4810, 38261]
    for e in golden["varint"]:
        data = bytes.fromhex(e["hex"])
        buf = C.create_string_buffer(data, max(len(data), 1))
        val = C.c_uint32(0)
        rc = lib.csnappy_get_uncompressed_length(buf, len(data), C.byref(val))
        assert rc == e["rc"], e
        if rc > 0:
            assert val.value == e["value"], e
    assert lib.csnappy_bc_max_container_length(10000, 4096) == 4 + 4 * 3 + 10000


def test_no_cpu_fallback_without_a_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device behaviour cannot be observed")
    assert lib.csnappy_b200_device_ok() == 0
    src = C.create_string_buffer(bytes.fromhex("0861626301"), 8)
    dst = C.create_string_buffer(64)
    n = C.c_uint32(64)
    assert lib.csnappy_decompress_noheader(src, 5, dst, C.byref(n)) == -100  # CSNAPPY_E_DEVICE
    assert n.value == 64
    assert lib.csnappy_decompress(src, 5, dst, 64) == -100
    out_len = C.c_uint64(0)
    cont = C.create_string_buffer(64)
    assert lib.csnappy_bc_compress_host(src, 5, 4096, cont, 64, C.byref(out_len), 13) == -100
    assert b"csnappy_b200" in lib.csnappy_b200_last_error()
    # bad arguments are rejected before any device work
    assert lib.csnappy_b200_set_tuning(b"compress_lanes", 7) == -101
    assert lib.csnappy_bc_compress_host(src, 5, 0, cont, 64, C.byref(out_len), 13) == -101


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "csnappy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "liboracle" not in text and "snappy_oracle" not in text, f


C_MAIN = os.path.join(ROOT, "tests", "c", "dropin_main.c")


def build_c_caller(tmpdir) -> str:
    """A plain C program built only against include/ and linked against the library,
    like the reference's own callers (cl_tester, block_compressor, the zram glue)."""
    from csnappy_b200._lib import LIB_PATH

    exe = os.path.join(str(tmpdir), "dropin_main")
    libdir = os.path.dirname(LIB_PATH)
    subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Wextra", "-Werror", "-O2", "-I" + os.path.join(ROOT, "include"),
                    "-o", exe, C_MAIN, "-L" + libdir, "-lcsnappy_b200", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_c_caller_compiles_and_links(lib, tmp_path):
    assert os.path.exists(build_c_caller(tmp_path))
