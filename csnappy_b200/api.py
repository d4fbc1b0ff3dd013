"""Python mirror of the reference's C interface, bound to libcsnappy_b200.so.

Names, argument meaning and error behaviour follow /root/reference/csnappy.h:30-129
so that the parity tests read like the reference's own callers (cl_tester.c:14-114,
167-238; block_compressor.c:113-134).  Every function below goes through the C-ABI
(ctypes) into the CUDA kernels; nothing here compresses or decompresses on the CPU.

Host-buffer calls (bytes in / bytes out):
    csnappy_max_compressed_length, csnappy_compress_fragment, csnappy_compress,
    csnappy_get_uncompressed_length, csnappy_decompress, csnappy_decompress_noheader
Device-resident batches (torch CUDA tensors, asynchronous on the current stream):
    batch_compress_fragments, batch_decompress, batch_pack
Host-buffer batches (numpy / pinned torch tensors, synchronous, pipelined H2D/D2H):
    batch_compress_fragments_host, batch_decompress_host
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import last_error, lib

CSNAPPY_VERSION = 5
CSNAPPY_WORKMEM_BYTES_POWER_OF_TWO = 16
CSNAPPY_WORKMEM_BYTES = 1 << 16
CSNAPPY_E_OK = 0
CSNAPPY_E_HEADER_BAD = -1
CSNAPPY_E_OUTPUT_INSUF = -2
CSNAPPY_E_OUTPUT_OVERRUN = -3
CSNAPPY_E_INPUT_NOT_CONSUMED = -4
CSNAPPY_E_DATA_MALFORMED = -5
CSNAPPY_E_DEVICE = -100
CSNAPPY_E_BAD_ARG = -101
BATCH_SHRINK_TABLE = 1
BATCH_WITH_HEADER = 2
BATCH_RAW_IF_FULL = 4
FRAGMENT_MAX = 32768


class CsnappyDeviceError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise CsnappyDeviceError(f"{what} failed (rc={rc}): {last_error()}")


def _in_buf(data) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(a)


# ----------------------------------------------------------------------------- csnappy.h mirror
def csnappy_max_compressed_length(source_len: int) -> int:
    return lib().csnappy_max_compressed_length(source_len)


def csnappy_compress_fragment(data: bytes, workmem_bytes_power_of_two: int = 15) -> bytes:
    """One fragment (<= 32 KiB), no length prefix.  csnappy.h:46-52."""
    n = len(data)
    src = _in_buf(data)
    out = np.empty(csnappy_max_compressed_length(n), dtype=np.uint8)
    end = lib().csnappy_compress_fragment(src.ctypes.data, n, out.ctypes.data, None, workmem_bytes_power_of_two)
    return out[: end - out.ctypes.data].tobytes()


def csnappy_compress(data: bytes, workmem_bytes_power_of_two: int = CSNAPPY_WORKMEM_BYTES_POWER_OF_TWO) -> bytes:
    """Whole buffer with varint32 prefix.  csnappy.h:65-72."""
    n = len(data)
    src = _in_buf(data)
    out = np.empty(csnappy_max_compressed_length(n) + 32 * (n // FRAGMENT_MAX + 1), dtype=np.uint8)
    olen = C.c_uint32(0)
    lib().csnappy_compress(src.ctypes.data, n, out.ctypes.data, C.byref(olen), None, workmem_bytes_power_of_two)
    return out[: olen.value].tobytes()


def csnappy_get_uncompressed_length(data: bytes):
    """-> (rc, value): rc = header bytes consumed (1..5) or CSNAPPY_E_HEADER_BAD.  csnappy.h:83-87."""
    src = _in_buf(data)
    val = C.c_uint32(0)
    rc = lib().csnappy_get_uncompressed_length(src.ctypes.data, len(data), C.byref(val))
    return rc, val.value


def csnappy_decompress(data: bytes, dst_len: int):
    """-> (rc, dst bytes[0:dst_len]).  csnappy.h:99-104."""
    src = _in_buf(data)
    out = np.zeros(max(dst_len, 1), dtype=np.uint8)
    rc = lib().csnappy_decompress(src.ctypes.data, len(data), out.ctypes.data, dst_len)
    return rc, out[:dst_len].tobytes()


def csnappy_decompress_noheader(data: bytes, dst_capacity: int):
    """-> (rc, produced bytes or None).  *dst_len is only meaningful on success.  csnappy.h:114-119."""
    src = _in_buf(data)
    out = np.zeros(max(dst_capacity, 1), dtype=np.uint8)
    dl = C.c_uint32(dst_capacity)
    rc = lib().csnappy_decompress_noheader(src.ctypes.data, len(data), out.ctypes.data, C.byref(dl))
    if rc != 0:
        assert dl.value == dst_capacity, "*dst_len must be untouched on error"
        return rc, None
    return rc, out[: dl.value].tobytes()


# ----------------------------------------------------------------------------- device batches
def _stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def out_stride_for(block_len: int) -> int:
    """Output slot stride for compression: max_compressed_length rounded up to 16 bytes."""
    return (32 + block_len + block_len // 6 + 15) // 16 * 16


def batch_compress_fragments(d_in, block_len: int, n_blocks: int, wm: int, *, in_stride: int | None = None,
                             in_len=None, in_off=None, out=None, out_len=None, out_stride: int | None = None,
                             flags: int = 0):
    """Device-resident batch of csnappy_compress_fragment.  d_in: uint8 CUDA tensor.
    Returns (out uint8 [n_blocks*out_stride], out_len int32-as-uint32 [n_blocks]) CUDA tensors."""
    import torch

    in_stride = block_len if in_stride is None else in_stride
    out_stride = out_stride_for(block_len) if out_stride is None else out_stride
    if out is None:
        out = torch.empty(max(n_blocks * out_stride, 16), dtype=torch.uint8, device=d_in.device)
    if out_len is None:
        out_len = torch.empty(max(n_blocks, 1), dtype=torch.int32, device=d_in.device)
    rc = lib().csnappy_batch_compress_fragments(_ptr(d_in), _ptr(in_off), in_stride, _ptr(in_len), block_len,
                                                 n_blocks, _ptr(out), out_stride, _ptr(out_len), wm, flags,
                                                 _stream_ptr())
    _check(rc, "csnappy_batch_compress_fragments")
    return out, out_len


def batch_decompress(d_in, in_len, n_blocks: int, out_cap: int, *, in_stride: int = 0, in_off=None, out=None,
                     out_stride: int | None = None, out_caps=None, out_len=None, status=None, flags: int = 0):
    """Device-resident batch of csnappy_decompress_noheader (or csnappy_decompress with BATCH_WITH_HEADER).
    Returns (out, out_len, status) CUDA tensors."""
    import torch

    out_stride = (out_cap + 15) // 16 * 16 if out_stride is None else out_stride
    if out is None:
        out = torch.empty(max(n_blocks * out_stride, 16), dtype=torch.uint8, device=d_in.device)
    if out_len is None:
        out_len = torch.empty(max(n_blocks, 1), dtype=torch.int32, device=d_in.device)
    if status is None:
        status = torch.empty(max(n_blocks, 1), dtype=torch.int32, device=d_in.device)
    rc = lib().csnappy_batch_decompress(_ptr(d_in), _ptr(in_off), in_stride, _ptr(in_len), n_blocks, _ptr(out),
                                         out_stride, _ptr(out_caps), out_cap, _ptr(out_len), _ptr(status), flags,
                                         _stream_ptr())
    _check(rc, "csnappy_batch_decompress")
    return out, out_len, status


def batch_pack(d_slots, slot_stride: int, d_len, n_blocks: int, packed=None):
    """Exclusive scan of sizes + gather of slots.  Returns (packed uint8, off int64 [n_blocks+1])."""
    import torch

    off = torch.empty(n_blocks + 1, dtype=torch.int64, device=d_slots.device)
    if packed is None:
        packed = torch.empty(max(n_blocks * slot_stride, 16), dtype=torch.uint8, device=d_slots.device)
    rc = lib().csnappy_batch_pack(_ptr(d_slots), slot_stride, _ptr(d_len), n_blocks, _ptr(packed), _ptr(off),
                                   _stream_ptr())
    _check(rc, "csnappy_batch_pack")
    return packed, off


# ----------------------------------------------------------------------------- host batches
def _host_ptr(x):
    return x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr()


def batch_compress_fragments_host(h_in, block_len: int, n_blocks: int, wm: int, h_out, h_out_len, *,
                                  in_stride: int | None = None, out_stride: int | None = None):
    """Host buffers (numpy arrays or pinned torch CPU tensors); synchronous."""
    in_stride = block_len if in_stride is None else in_stride
    out_stride = out_stride_for(block_len) if out_stride is None else out_stride
    rc = lib().csnappy_batch_compress_fragments_host(_host_ptr(h_in), in_stride, block_len, n_blocks,
                                                      _host_ptr(h_out), out_stride, _host_ptr(h_out_len), wm)
    _check(rc, "csnappy_batch_compress_fragments_host")


def batch_decompress_host(h_in, in_stride: int, h_in_len, n_blocks: int, h_out, out_stride: int, out_cap: int,
                          h_out_len, h_status, flags: int = 0):
    rc = lib().csnappy_batch_decompress_host(_host_ptr(h_in), in_stride, _host_ptr(h_in_len), n_blocks,
                                              _host_ptr(h_out), out_stride, out_cap, _host_ptr(h_out_len),
                                              _host_ptr(h_status), flags)
    _check(rc, "csnappy_batch_decompress_host")


# ----------------------------------------------------------------------------- block_compressor container
def bc_max_container_length(input_length: int, page_size: int = 4096) -> int:
    return lib().csnappy_bc_max_container_length(input_length, page_size)


def bc_compress_host(h_in, input_length: int, h_container, wm: int = 13, page_size: int = 4096) -> int:
    """block_compressor-style container of `input_length` bytes at h_in (numpy / pinned torch) into
    h_container; returns the container length.  Reference: block_compressor.c:275-345."""
    clen = C.c_uint64(0)
    cap = h_container.nbytes if isinstance(h_container, np.ndarray) else h_container.numel() * h_container.element_size()
    rc = lib().csnappy_bc_compress_host(_host_ptr(h_in), input_length, page_size, _host_ptr(h_container), cap,
                                        C.byref(clen), wm)
    _check(rc, "csnappy_bc_compress_host")
    return clen.value


def bc_decompress_host(h_container, container_length: int, h_out, page_size: int = 4096):
    """-> (rc, bytes produced, failed page or None).  Reference: block_compressor.c:347-394."""
    olen = C.c_uint64(0)
    bad = C.c_uint32(0xFFFFFFFF)
    cap = h_out.nbytes if isinstance(h_out, np.ndarray) else h_out.numel() * h_out.element_size()
    rc = lib().csnappy_bc_decompress_host(_host_ptr(h_container), container_length, page_size, _host_ptr(h_out), cap,
                                          C.byref(olen), C.byref(bad))
    if rc in (CSNAPPY_E_DEVICE, CSNAPPY_E_BAD_ARG):
        _check(rc, "csnappy_bc_decompress_host")
    return rc, olen.value, (bad.value if rc != 0 and bad.value != 0xFFFFFFFF else None)


def bc_compress_host_multi(h_in, input_length: int, h_container, wm: int = 13, page_size: int = 4096, devices=None) -> int:
    """bc_compress_host over several devices from this one process (devices: list of device indices, None = all
    visible).  Same bytes as the single-device call."""
    clen = C.c_uint64(0)
    cap = h_container.nbytes if isinstance(h_container, np.ndarray) else h_container.numel() * h_container.element_size()
    dev = (C.c_int * len(devices))(*devices) if devices else None
    rc = lib().csnappy_bc_compress_host_multi(_host_ptr(h_in), input_length, page_size, _host_ptr(h_container), cap,
                                              C.byref(clen), wm, dev, len(devices) if devices else 0)
    _check(rc, "csnappy_bc_compress_host_multi")
    return clen.value


def bc_decompress_host_multi(h_container, container_length: int, h_out, page_size: int = 4096, devices=None):
    """-> (rc, bytes produced, failed page or None) over several devices."""
    olen = C.c_uint64(0)
    bad = C.c_uint32(0xFFFFFFFF)
    cap = h_out.nbytes if isinstance(h_out, np.ndarray) else h_out.numel() * h_out.element_size()
    dev = (C.c_int * len(devices))(*devices) if devices else None
    rc = lib().csnappy_bc_decompress_host_multi(_host_ptr(h_container), container_length, page_size, _host_ptr(h_out),
                                                cap, C.byref(olen), C.byref(bad), dev, len(devices) if devices else 0)
    if rc in (CSNAPPY_E_DEVICE, CSNAPPY_E_BAD_ARG):
        _check(rc, "csnappy_bc_decompress_host_multi")
    return rc, olen.value, (bad.value if rc != 0 and bad.value != 0xFFFFFFFF else None)


def batch_compress(d_in, in_off, in_len, wm: int, *, out_stride: int | None = None):
    """Device-resident batch of csnappy_compress (whole buffers, varint32 header + 32 KiB fragments each).
    d_in: uint8 CUDA tensor; in_off / in_len: HOST sequences (offsets into d_in, lengths).
    Returns (out uint8 [n * out_stride], out_len int32 [n], out_stride)."""
    import torch

    n = len(in_len)
    h_len = np.ascontiguousarray(np.asarray(in_len, dtype=np.uint32))
    h_off = np.ascontiguousarray(np.asarray(in_off, dtype=np.uint64))
    longest = int(h_len.max()) if n else 0
    if out_stride is None:
        out_stride = (5 + longest + longest // 6 + 32 * max(1, -(-longest // FRAGMENT_MAX)) + 15) // 16 * 16
    ws_bytes = lib().csnappy_batch_compress_workspace(h_len.ctypes.data, 0, n)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=d_in.device)
    out = torch.empty(max(n * out_stride, 16), dtype=torch.uint8, device=d_in.device)
    out_len = torch.empty(max(n, 1), dtype=torch.int32, device=d_in.device)
    rc = lib().csnappy_batch_compress(_ptr(d_in), h_off.ctypes.data, 0, h_len.ctypes.data, 0, n, _ptr(out), out_stride,
                                      _ptr(out_len), wm, _ptr(ws), ws_bytes, _stream_ptr())
    _check(rc, "csnappy_batch_compress")
    torch.cuda.current_stream().synchronize()  # the workspace dies with this frame
    return out, out_len, out_stride


def stream_decompress(d_src, src_len: int, out_cap: int, *, out=None, workspace=None, result=None):
    """ONE long raw stream, device resident, through the parallel stream decoder.
    Returns (out uint8 [out_cap], result int32 [2] = [out_len, status]) CUDA tensors (asynchronous)."""
    import torch

    ws_bytes = lib().csnappy_stream_decompress_workspace(src_len, out_cap)
    if workspace is None:
        workspace = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=d_src.device)
    if out is None:
        out = torch.empty(max(out_cap, 16), dtype=torch.uint8, device=d_src.device)
    if result is None:
        result = torch.empty(2, dtype=torch.int32, device=d_src.device)
    rc = lib().csnappy_stream_decompress(_ptr(d_src), src_len, _ptr(out), out_cap, result.data_ptr(),
                                         result.data_ptr() + 4, _ptr(workspace), workspace.numel(), _stream_ptr())
    _check(rc, "csnappy_stream_decompress")
    return out, result


def device_count() -> int:
    return lib().csnappy_b200_device_count()


def set_tuning(key: str, value: int):
    rc = lib().csnappy_b200_set_tuning(key.encode(), value)
    if rc != 0:
        raise ValueError(f"bad tuning {key}={value}")


def kernel_launches() -> int:
    return lib().csnappy_b200_kernel_launches()


def device_ok() -> bool:
    return bool(lib().csnappy_b200_device_ok())
