"""ctypes bindings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
--impl reference legs.  Nothing under csnappy_b200/ imports this package.

Two implementations sit behind the same Python surface:

* ``port``       oracle/snappy_oracle.c, our restatement (liboracle.so)
* ``reference``  the unmodified reference compiled by oracle/Makefile into
                 oracle/_ref/libcsnappy_ref.so (present when /root/reference was
                 available at build time; the prebuilt file travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libcsnappy_ref.so")

E_OK, E_HEADER_BAD, E_OUTPUT_INSUF, E_OUTPUT_OVERRUN, E_DATA_MALFORMED = 0, -1, -2, -3, -5


def build() -> None:
    """Compile liboracle.so and, when /root/reference exists, _ref/libcsnappy_ref.so."""
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def max_compressed_length(n: int) -> int:
    return 32 + n + n // 6


class _Port:
    kind = "port"

    def __init__(self):
        if not os.path.exists(_PORT_SO):
            build()
        L = C.CDLL(_PORT_SO)
        self.L = L
        L.oracle_compress_fragment.restype = C.c_uint32
        L.oracle_compress_fragment.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.oracle_compress.restype = None
        L.oracle_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_int]
        L.oracle_decompress_noheader.restype = C.c_int
        L.oracle_decompress_noheader.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        L.oracle_decompress.restype = C.c_int
        L.oracle_decompress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.oracle_get_uncompressed_length.restype = C.c_int
        L.oracle_get_uncompressed_length.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oracle_chunk_wm.restype = C.c_int
        L.oracle_chunk_wm.argtypes = [C.c_uint32, C.c_int]
        L.harness_open_ref.restype = C.c_int
        L.harness_open_ref.argtypes = [C.c_char_p]
        L.harness_compress_pages.restype = C.c_double
        L.harness_compress_pages.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                             C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        L.harness_decompress_pages.restype = C.c_double
        L.harness_decompress_pages.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                               C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                               C.c_int]

    # -- single-buffer calls (bytes in, bytes out) -------------------------
    def compress_fragment(self, data: bytes, wm: int) -> bytes:
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max_compressed_length(len(data)) + 16, dtype=np.uint8)
        n = self.L.oracle_compress_fragment(src.ctypes.data, len(data), out.ctypes.data, wm)
        return out[:n].tobytes()

    def compress(self, data: bytes, wm: int) -> bytes:
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max_compressed_length(len(data)) + 16 + 64 * (len(data) // 32768 + 1), dtype=np.uint8)
        n = C.c_uint32(0)
        self.L.oracle_compress(src.ctypes.data, len(data), out.ctypes.data, C.byref(n), wm)
        return out[: n.value].tobytes()

    def decompress_noheader(self, data: bytes, cap: int):
        """-> (rc, produced bytes or None); dst_len is only meaningful when rc == 0."""
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max(cap, 1) + 64, dtype=np.uint8)
        n = C.c_uint32(cap)
        rc = self.L.oracle_decompress_noheader(src.ctypes.data, len(data), out.ctypes.data, C.byref(n))
        return rc, (out[: n.value].tobytes() if rc == 0 else None)

    def decompress(self, data: bytes, dst_len: int):
        """Mirrors csnappy_decompress: -> (rc, dst buffer bytes of dst_len)."""
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max(dst_len, 1) + 64, dtype=np.uint8)
        rc = self.L.oracle_decompress(src.ctypes.data, len(data), out.ctypes.data, dst_len)
        return rc, out[:dst_len].tobytes()

    def get_uncompressed_length(self, data: bytes):
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        r = C.c_uint32(0)
        rc = self.L.oracle_get_uncompressed_length(src.ctypes.data, len(data), C.byref(r))
        return rc, r.value

    def chunk_wm(self, chunk_len: int, wm: int) -> int:
        return self.L.oracle_chunk_wm(chunk_len, wm)


class _Reference:
    """The unmodified reference library, same Python surface as _Port."""

    kind = "reference"

    def __init__(self):
        L = C.CDLL(_REF_SO)
        self.L = L
        L.csnappy_compress_fragment.restype = C.c_void_p
        L.csnappy_compress_fragment.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.csnappy_compress.restype = None
        L.csnappy_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_int]
        L.csnappy_decompress_noheader.restype = C.c_int
        L.csnappy_decompress_noheader.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        L.csnappy_decompress.restype = C.c_int
        L.csnappy_decompress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.csnappy_get_uncompressed_length.restype = C.c_int
        L.csnappy_get_uncompressed_length.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.csnappy_max_compressed_length.restype = C.c_uint32
        L.csnappy_max_compressed_length.argtypes = [C.c_uint32]
        self._wm = np.zeros(1 << 16, dtype=np.uint8)

    def compress_fragment(self, data: bytes, wm: int) -> bytes:
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max_compressed_length(len(data)) + 16, dtype=np.uint8)
        end = self.L.csnappy_compress_fragment(src.ctypes.data, len(data), out.ctypes.data, self._wm.ctypes.data, wm)
        return out[: end - out.ctypes.data].tobytes()

    def compress(self, data: bytes, wm: int) -> bytes:
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max_compressed_length(len(data)) + 16 + 64 * (len(data) // 32768 + 1), dtype=np.uint8)
        n = C.c_uint32(0)
        self.L.csnappy_compress(src.ctypes.data, len(data), out.ctypes.data, C.byref(n), self._wm.ctypes.data, wm)
        return out[: n.value].tobytes()

    def decompress_noheader(self, data: bytes, cap: int):
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max(cap, 1) + 64, dtype=np.uint8)
        n = C.c_uint32(cap)
        rc = self.L.csnappy_decompress_noheader(src.ctypes.data, len(data), out.ctypes.data, C.byref(n))
        return rc, (out[: n.value].tobytes() if rc == 0 else None)

    def decompress(self, data: bytes, dst_len: int):
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        out = np.zeros(max(dst_len, 1) + 64, dtype=np.uint8)
        rc = self.L.csnappy_decompress(src.ctypes.data, len(data), out.ctypes.data, dst_len)
        return rc, out[:dst_len].tobytes()

    def get_uncompressed_length(self, data: bytes):
        src = np.frombuffer(bytes(data) + b"\0" * 16, dtype=np.uint8)
        r = C.c_uint32(0)
        rc = self.L.csnappy_get_uncompressed_length(src.ctypes.data, len(data), C.byref(r))
        return rc, r.value


_port = None
_ref = None


def port() -> _Port:
    global _port
    if _port is None:
        _port = _Port()
    return _port


def have_reference() -> bool:
    return os.path.exists(_REF_SO)


def reference() -> _Reference:
    global _ref
    if _ref is None:
        _ref = _Reference()
    return _ref


def best():
    """The strongest checker available: the real reference if built, else the port."""
    return reference() if have_reference() else port()


# -- bulk (pthread) entry points over strided numpy batches ----------------------
def _impl_id(impl: str) -> int:
    if impl == "reference":
        if port().L.harness_open_ref(_REF_SO.encode()) != 0:
            raise RuntimeError("oracle/_ref/libcsnappy_ref.so not loadable")
        return 1
    return 0


def batch_compress(pages: np.ndarray, wm: int, impl: str = "port", threads: int = 1, lens=None):
    """pages: uint8 [B, stride].  -> (out uint8 [B, out_stride], out_len uint32 [B], seconds)."""
    assert pages.dtype == np.uint8 and pages.ndim == 2 and pages.flags.c_contiguous
    B, stride = pages.shape
    out_stride = (max_compressed_length(stride) + 15) // 16 * 16
    out = np.zeros((B, out_stride), dtype=np.uint8)
    out_len = np.zeros(B, dtype=np.uint32)
    lp = None
    if lens is not None:
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        lp = lens.ctypes.data
    # the reference's literal fast path reads up to 15 B past a short literal: pad the source
    src = np.concatenate([pages.reshape(-1), np.zeros(64, np.uint8)])
    sec = port().L.harness_compress_pages(_impl_id(impl), src.ctypes.data, stride, lp, stride, B,
                                          out.ctypes.data, out_stride, out_len.ctypes.data, wm, threads)
    if sec < 0:
        raise RuntimeError("cpu harness failed")
    return out, out_len, sec


def batch_decompress(comp: np.ndarray, comp_len: np.ndarray, cap: int, impl: str = "port", threads: int = 1):
    """comp: uint8 [B, in_stride].  -> (out uint8 [B, out_stride], out_len, status, seconds)."""
    assert comp.dtype == np.uint8 and comp.ndim == 2 and comp.flags.c_contiguous
    B, in_stride = comp.shape
    comp_len = np.ascontiguousarray(comp_len, dtype=np.uint32)
    out_stride = (cap + 64 + 15) // 16 * 16
    out = np.zeros((B, out_stride), dtype=np.uint8)
    out_len = np.zeros(B, dtype=np.uint32)
    status = np.zeros(B, dtype=np.int32)
    src = np.concatenate([comp.reshape(-1), np.zeros(64, np.uint8)])
    sec = port().L.harness_decompress_pages(_impl_id(impl), src.ctypes.data, in_stride, comp_len.ctypes.data, B,
                                            out.ctypes.data, out_stride, cap, out_len.ctypes.data,
                                            status.ctypes.data, threads)
    if sec < 0:
        raise RuntimeError("cpu harness failed")
    return out, out_len, status, sec


class BatchRunner:
    """compress + decompress one strided batch on the CPU, repeatedly, with buffers allocated and touched ONCE
    (a fresh np.zeros per pass would put first-touch page faults of the output into every timed pass).
    units: uint8 [B, unit_len].  The one CPU-baseline protocol of bench.py: `measure(warmup, steps)` = mean seconds
    of `steps` passes after `warmup` passes, compress and decompress timed separately inside the C harness."""

    def __init__(self, units: np.ndarray, wm: int, impl: str = "port", threads: int = 1):
        assert units.dtype == np.uint8 and units.ndim == 2 and units.flags.c_contiguous
        self.B, self.unit = units.shape
        self.wm, self.threads, self.impl_id = wm, threads, _impl_id(impl)
        self.units = units
        # the reference's literal fast path reads up to 15 B past a short literal: pad the source
        self.src = np.concatenate([units.reshape(-1), np.zeros(64, np.uint8)])
        self.out_stride = (max_compressed_length(self.unit) + 15) // 16 * 16
        self.comp = np.zeros(self.B * self.out_stride + 64, dtype=np.uint8)
        self.comp_len = np.zeros(self.B, dtype=np.uint32)
        self.back_stride = (self.unit + 64 + 15) // 16 * 16
        self.back = np.zeros(self.B * self.back_stride, dtype=np.uint8)
        self.back_len = np.zeros(self.B, dtype=np.uint32)
        self.status = np.zeros(self.B, dtype=np.int32)
        self.comp[::4096] = 1  # touch every page of the output buffers before anything is timed
        self.back[::4096] = 1

    def compress(self) -> float:
        sec = port().L.harness_compress_pages(self.impl_id, self.src.ctypes.data, self.unit, None, self.unit, self.B,
                                              self.comp.ctypes.data, self.out_stride, self.comp_len.ctypes.data,
                                              self.wm, self.threads)
        if sec < 0:
            raise RuntimeError("cpu harness failed")
        return sec

    def decompress(self) -> float:
        sec = port().L.harness_decompress_pages(self.impl_id, self.comp.ctypes.data, self.out_stride,
                                                self.comp_len.ctypes.data, self.B, self.back.ctypes.data,
                                                self.back_stride, self.unit, self.back_len.ctypes.data,
                                                self.status.ctypes.data, self.threads)
        if sec < 0:
            raise RuntimeError("cpu harness failed")
        return sec

    def measure(self, warmup: int, steps: int):
        """-> (mean compress seconds, mean decompress seconds) over `steps` passes; checks the round trip once."""
        tc = td = 0.0
        for it in range(warmup + steps):
            c, d = self.compress(), self.decompress()
            if it >= warmup:
                tc, td = tc + c, td + d
        back = self.back.reshape(self.B, self.back_stride)[:, : self.unit]
        assert (self.status == 0).all() and (self.back_len == self.unit).all() and (back == self.units).all()
        return tc / steps, td / steps

    def compressed(self, i: int) -> bytes:
        return self.comp[i * self.out_stride: i * self.out_stride + int(self.comp_len[i])].tobytes()
