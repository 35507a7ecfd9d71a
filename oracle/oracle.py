"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/blosc_oracle.h.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by
the product package.  Every function works on numpy uint8 arrays (or bytes).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

OK, EINVALID_DATA, EINVALID_HEADER, EINVALID_VERSION, EINVALID_CODEC = 0, 1, 2, 3, 4
ESIZE_MISMATCH, EDATA_TOO_LARGE, ECOMPRESSION_FAILED, EDECOMPRESSION_FAILED = 5, 6, 7, 8
EUNSUPPORTED, EDST_TOO_SMALL = 10, 11
NOSHUFFLE, SHUFFLE, BITSHUFFLE = 0, 1, 2
BLOSCLZ, LZ4, LZ4HC, SNAPPY, ZLIB, ZSTD = 0, 1, 2, 3, 4, 5
FLAG_SHUFFLE, FLAG_MEMCPY, FLAG_BITSHUFFLE = 1, 2, 4
MEMCPY_SHUFFLED, MEMCPY_REF_QUIRK = 0, 1


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with gcc (seconds)."""
    deps = [os.path.join(_HERE, f) for f in ("blosc_oracle.c", "blosc1_blocks.c", "blosc_oracle.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in deps)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, i64, u32p, u64p = C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.c_void_p
        for name in ("orc_shuffle", "orc_unshuffle", "orc_bitshuffle", "orc_bitunshuffle",
                     "orc_shuffle_fast", "orc_unshuffle_fast"):
            getattr(L, name).argtypes = [vp, vp, sz, i64]
            getattr(L, name).restype = None
        L.orc_lz4_bound.argtypes = [sz]; L.orc_lz4_bound.restype = sz
        L.orc_lz4_compress.argtypes = [vp, sz, vp, sz]; L.orc_lz4_compress.restype = sz
        L.orc_lz4_decompress.argtypes = [vp, sz, vp, sz]; L.orc_lz4_decompress.restype = i64
        L.orc_compress.argtypes = [vp, sz, C.c_int, C.c_int, C.c_int, i64, C.c_int, vp, sz,
                                   C.POINTER(sz)]
        L.orc_compress.restype = C.c_int
        L.orc_decompress.argtypes = [vp, sz, i64, vp, sz, C.POINTER(sz)]
        L.orc_decompress.restype = C.c_int
        L.orc_compress_batch_mt.argtypes = [vp, u64p, u32p, C.c_uint32, C.c_int, i64, vp, u64p,
                                            u32p, C.c_int, C.c_int]
        L.orc_compress_batch_mt.restype = C.c_int
        L.orc_decompress_batch_mt.argtypes = [vp, u64p, u32p, C.c_uint32, vp, u64p, u32p, C.c_int,
                                              C.c_int]
        L.orc_decompress_batch_mt.restype = C.c_int
        L.orc_shuffle_mt.argtypes = [C.c_int, C.c_int, i64, vp, vp, sz, C.c_int, C.c_int]
        L.orc_shuffle_mt.restype = C.c_int
        L.orc_blocks_blocksize.argtypes = [sz, i64, C.c_uint32]; L.orc_blocks_blocksize.restype = C.c_uint32
        L.orc_blocks_compress.argtypes = [vp, sz, C.c_int, i64, C.c_uint32, C.c_int, vp, sz, C.POINTER(sz)]
        L.orc_blocks_compress.restype = C.c_int
        L.orc_blocks_decompress.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
        L.orc_blocks_decompress.restype = C.c_int
        _lib = L
    return _lib


def _u8(a) -> np.ndarray:
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(a), dtype=np.uint8)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _filter(name, data, typesize):
    s = _u8(data)
    d = np.empty_like(s)
    getattr(lib(), name)(_ptr(s), _ptr(d), s.size, int(typesize))
    return d


def shuffle(data, typesize):
    return _filter("orc_shuffle", data, typesize)


def unshuffle(data, typesize):
    return _filter("orc_unshuffle", data, typesize)


def bitshuffle(data, typesize):
    return _filter("orc_bitshuffle", data, typesize)


def bitunshuffle(data, typesize):
    return _filter("orc_bitunshuffle", data, typesize)


def shuffle_fast(data, typesize):
    return _filter("orc_shuffle_fast", data, typesize)


def unshuffle_fast(data, typesize):
    return _filter("orc_unshuffle_fast", data, typesize)


def lz4_bound(n: int) -> int:
    return int(lib().orc_lz4_bound(n))


def lz4_compress(data) -> np.ndarray:
    s = _u8(data)
    d = np.empty(lz4_bound(s.size), dtype=np.uint8)
    c = lib().orc_lz4_compress(_ptr(s), s.size, _ptr(d), d.size)
    assert c > 0
    return d[:c].copy()


def lz4_decompress(payload, cap: int):
    """Returns the decoded bytes, or None if the stream is malformed / overruns cap."""
    s = _u8(payload)
    d = np.empty(max(cap, 1), dtype=np.uint8)
    r = lib().orc_lz4_decompress(_ptr(s), s.size, _ptr(d), cap)
    if r < 0:
        return None
    return d[:r].copy()


def compress(data, codec=LZ4, level=5, shuffle=SHUFFLE, typesize=4, memcpy_policy=MEMCPY_SHUFFLED):
    """Returns (status, frame-or-None)."""
    s = _u8(data)
    d = np.empty(16 + s.size + 16, dtype=np.uint8)
    out = C.c_size_t(0)
    rc = lib().orc_compress(_ptr(s) if s.size else None, s.size, codec, level, shuffle,
                            int(typesize), memcpy_policy, _ptr(d), d.size, C.byref(out))
    return rc, (d[:out.value].copy() if rc == 0 else None)


def decompress(frame, typesize_override=0, cap=None):
    """Returns (status, data-or-None)."""
    s = _u8(frame)
    if cap is None:
        cap = int(np.frombuffer(s[4:8].tobytes(), dtype="<u4")[0]) if s.size >= 16 else 0
        cap = min(cap, 1 << 31)
    d = np.empty(max(cap, 1), dtype=np.uint8)
    out = C.c_size_t(0)
    rc = lib().orc_decompress(_ptr(s) if s.size else None, s.size, int(typesize_override), _ptr(d),
                              cap, C.byref(out))
    return rc, (d[:out.value].copy() if rc == 0 else None)


# ---- opt-in Blosc-1 multi-block frames (oracle/blosc1_blocks.c; parity unpinned) -----------
B1_FLAG_DONTSPLIT, B1_LZ4_FORMAT = 0x10, 1


def blocks_blocksize(n: int, typesize: int, blocksize: int = 0) -> int:
    return int(lib().orc_blocks_blocksize(n, int(typesize), int(blocksize)))


def blocks_compress(data, shuffle=SHUFFLE, typesize=4, blocksize=0, split=False):
    """Returns (status, frame-or-None)."""
    s = _u8(data)
    d = np.empty(16 + s.size + 64, dtype=np.uint8)
    out = C.c_size_t(0)
    rc = lib().orc_blocks_compress(_ptr(s) if s.size else None, s.size, int(shuffle), int(typesize),
                                   int(blocksize), int(bool(split)), _ptr(d), d.size, C.byref(out))
    return rc, (d[:out.value].copy() if rc == 0 else None)


def blocks_decompress(frame, cap=None):
    """Returns (status, data-or-None)."""
    s = _u8(frame)
    if cap is None:
        cap = int(np.frombuffer(s[4:8].tobytes(), dtype="<u4")[0]) if s.size >= 16 else 0
        cap = min(cap, 1 << 31)
    d = np.empty(max(cap, 1), dtype=np.uint8)
    out = C.c_size_t(0)
    rc = lib().orc_blocks_decompress(_ptr(s) if s.size else None, s.size, _ptr(d), cap, C.byref(out))
    return rc, (d[:out.value].copy() if rc == 0 else None)


def compress_batch_mt(src, src_off, src_len, shuffle, typesize, threads, fast=1):
    """Compress frames into fixed slots of 16+len bytes. Returns (rc, dst, dst_off, dst_len)."""
    s = _u8(src)
    src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
    src_len = np.ascontiguousarray(src_len, dtype=np.uint32)
    slots = src_len.astype(np.uint64) + 16
    dst_off = np.zeros(len(slots), dtype=np.uint64)
    np.cumsum(slots[:-1], out=dst_off[1:])
    dst = np.empty(int(slots.sum()), dtype=np.uint8)
    dst_len = np.zeros(len(slots), dtype=np.uint32)
    rc = lib().orc_compress_batch_mt(_ptr(s), _ptr(src_off), _ptr(src_len), len(src_len), shuffle,
                                     int(typesize), _ptr(dst), _ptr(dst_off), _ptr(dst_len),
                                     threads, fast)
    return rc, dst, dst_off, dst_len


def decompress_batch_mt(frames, frame_off, frame_len, dst_off, dst_total, threads, fast=1):
    s = _u8(frames)
    frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
    frame_len = np.ascontiguousarray(frame_len, dtype=np.uint32)
    dst_off = np.ascontiguousarray(dst_off, dtype=np.uint64)
    dst = np.empty(dst_total, dtype=np.uint8)
    out_len = np.zeros(len(frame_len), dtype=np.uint32)
    rc = lib().orc_decompress_batch_mt(_ptr(s), _ptr(frame_off), _ptr(frame_len), len(frame_len),
                                       _ptr(dst), _ptr(dst_off), _ptr(out_len), threads, fast)
    return rc, dst, out_len


def shuffle_mt(mode, inverse, typesize, data, threads, fast=1):
    s = _u8(data)
    d = np.empty_like(s)
    lib().orc_shuffle_mt(mode, inverse, int(typesize), _ptr(s), _ptr(d), s.size, threads, fast)
    return d


# ---- independent LZ4 format referee: the system liblz4 (runtime only, no headers) ---------
_lz4 = None


def liblz4():
    """ctypes handle on liblz4.so.1 (1.9.4 in this image) or None when absent."""
    global _lz4
    if _lz4 is None:
        try:
            L = C.CDLL("liblz4.so.1")
            L.LZ4_decompress_safe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            L.LZ4_decompress_safe.restype = C.c_int
            L.LZ4_compress_default.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            L.LZ4_compress_default.restype = C.c_int
            L.LZ4_compressBound.argtypes = [C.c_int]
            L.LZ4_compressBound.restype = C.c_int
            _lz4 = L
        except OSError:
            _lz4 = False
    return _lz4 or None


def liblz4_decompress(payload, cap: int):
    L = liblz4()
    s = _u8(payload)
    d = np.empty(max(cap, 1), dtype=np.uint8)
    r = L.LZ4_decompress_safe(_ptr(s), _ptr(d), s.size, cap)
    return None if r < 0 else d[:r].copy()


def liblz4_compress(data) -> np.ndarray:
    L = liblz4()
    s = _u8(data)
    d = np.empty(L.LZ4_compressBound(s.size), dtype=np.uint8)
    r = L.LZ4_compress_default(_ptr(s), _ptr(d), s.size, d.size)
    assert r > 0
    return d[:r].copy()
