"""go-blosc_b200: Python host-side mirror of go-blosc's public API over the b2b C ABI.

The reference is a Go package (`blosc`); there is no Go toolchain in this image, so the
same C ABI (include/b2b.h -> lib/libb2b.so) that the cgo backend in go/blosc binds is driven
from here through ctypes.  Names and behaviour follow the reference:

    reference (blosc.go / shuffle.go)          here
    Compress(data, codec, level, shuffle, ts)  compress(...)
    CompressWithOptions(data, opts)            compress_with_options(...)
    Decompress / DecompressWithSize            decompress / decompress_with_size
    GetInfo / GetDecompressedSize / ParseHeader get_info / get_decompressed_size / parse_header
    ShuffleBuffer / UnshuffleBuffer            shuffle_buffer / unshuffle_buffer
    Codec, Shuffle, Options, Header, Err*      Codec, Shuffle, Options, Header, Err*

There is NO CPU fallback: importing works without a GPU (so the ABI can be inspected), but
every call that does work needs libb2b.so and a CUDA device and raises otherwise.
(The directory name has a hyphen; load it with importlib -- see __graft_entry__.load_package.)
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2B_LIB_PATH") or os.path.join(_HERE, "lib", "libb2b.so")   # override: A/B builds of the same ABI
VERSION = "1.0.0"          # blosc.go:49
FORMAT_VERSION = 2         # blosc.go:50
HEADER_SIZE = 16           # blosc.go:118-121
MIN_HEADER_SIZE = 16

# status codes of include/b2b.h
(OK, EINVALID_DATA, EINVALID_HEADER, EINVALID_VERSION, EINVALID_CODEC, ESIZE_MISMATCH,
 EDATA_TOO_LARGE, ECOMPRESSION_FAILED, EDECOMPRESSION_FAILED, ECUDA, EUNSUPPORTED,
 EDST_TOO_SMALL, EINVAL) = range(13)
OPT_REF_MEMCPY_QUIRK, OPT_FILTER_CTAS_PER_SM, OPT_HOST_STAGE_BYTES, OPT_HASH_LOG, OPT_KERNEL_TIMING, OPT_HASH_BYTES = 1, 2, 3, 4, 5, 6
OPT_HOST_THREADS, OPT_NO_HOST_STAGING, OPT_DECODER, OPT_FUSE_UNSHUFFLE, OPT_DECODE_STREAMS = 7, 8, 9, 10, 11


class Codec(enum.IntEnum):   # blosc.go:55-64
    BloscLZ = 0
    LZ4 = 1
    LZ4HC = 2
    Snappy = 3
    ZLIB = 4
    ZSTD = 5

    def __str__(self):       # blosc.go:67-84
        return {0: "blosclz", 1: "lz4", 2: "lz4hc", 3: "snappy", 4: "zlib", 5: "zstd"}[int(self)]


class Shuffle(enum.IntEnum):  # blosc.go:86-92
    NoShuffle = 0
    Shuffle1 = 1
    BitShuffle = 2

    def __str__(self):        # blosc.go:95-107
        return {0: "noshuffle", 1: "shuffle", 2: "bitshuffle"}[int(self)]


# ---- the reference's sentinel errors (blosc.go:125-149) ------------------------------------
class BloscError(Exception):
    status = -1


class ErrInvalidData(BloscError):
    status = EINVALID_DATA


class ErrInvalidHeader(BloscError):
    status = EINVALID_HEADER


class ErrInvalidVersion(BloscError):
    status = EINVALID_VERSION


class ErrInvalidCodec(BloscError):
    status = EINVALID_CODEC


class ErrSizeMismatch(BloscError):
    status = ESIZE_MISMATCH


class ErrDataTooLarge(BloscError):
    status = EDATA_TOO_LARGE


class ErrCompressionFailed(BloscError):
    status = ECOMPRESSION_FAILED


class ErrDecompressionFailed(BloscError):
    status = EDECOMPRESSION_FAILED


class ErrCuda(BloscError):
    status = ECUDA


class ErrUnsupported(BloscError):
    """Codec that the reference implements on the CPU and that is outside the GPU path."""
    status = EUNSUPPORTED


class ErrDstTooSmall(BloscError):
    status = EDST_TOO_SMALL


class ErrInval(BloscError):
    status = EINVAL


_ERRORS = {c.status: c for c in (ErrInvalidData, ErrInvalidHeader, ErrInvalidVersion, ErrInvalidCodec,
                                 ErrSizeMismatch, ErrDataTooLarge, ErrCompressionFailed,
                                 ErrDecompressionFailed, ErrCuda, ErrUnsupported, ErrDstTooSmall,
                                 ErrInval)}


@dataclass
class Options:               # blosc.go:227-234
    codec: int = Codec.LZ4
    level: int = 5
    shuffle: int = Shuffle.Shuffle1
    typesize: int = 4
    blocksize: int = 0       # declared and never read by the reference
    numthreads: int = 0      # reserved


def default_options() -> Options:   # blosc.go:237-245
    return Options()


class _CHeader(C.Structure):
    _fields_ = [("version", C.c_uint8), ("versionlz", C.c_uint8), ("flags", C.c_uint8),
                ("typesize", C.c_uint8), ("nbytes_orig", C.c_uint32), ("blocksize", C.c_uint32),
                ("nbytes_comp", C.c_uint32)]


@dataclass
class Header:                # blosc.go:154-162
    version: int
    versionlz: int
    flags: int
    typesize: int
    nbytes_orig: int
    blocksize: int
    nbytes_comp: int

    def has_shuffle(self) -> bool:      # blosc.go:201-203
        return bool(self.flags & 0x1)

    def has_bitshuffle(self) -> bool:   # blosc.go:206-208
        return bool(self.flags & 0x4)

    def is_memcpy(self) -> bool:        # blosc.go:211-213
        return bool(self.flags & 0x2)

    def shuffle_mode(self) -> Shuffle:  # blosc.go:216-224
        if self.has_bitshuffle():
            return Shuffle.BitShuffle
        if self.has_shuffle():
            return Shuffle.Shuffle1
        return Shuffle.NoShuffle

    def to_bytes(self) -> bytes:        # blosc.go:188-198
        h = _CHeader(self.version & 0xFF, self.versionlz & 0xFF, self.flags & 0xFF, self.typesize & 0xFF,
                     self.nbytes_orig & 0xFFFFFFFF, self.blocksize & 0xFFFFFFFF,
                     self.nbytes_comp & 0xFFFFFFFF)
        out = (C.c_uint8 * 16)()
        lib().b2b_header_bytes(C.byref(h), out)
        return bytes(out)


# ---- library loading ------------------------------------------------------------------------
# every symbol include/b2b.h declares: (name, restype, argtypes)
_vp, _sz, _i64, _u64, _u32, _int = C.c_void_p, C.c_size_t, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
ABI = [
    ("b2b_init", _int, [_int, C.POINTER(_vp)]),
    ("b2b_destroy", None, [_vp]),
    ("b2b_strerror", C.c_char_p, [_int]),
    ("b2b_version", C.c_char_p, []),
    ("b2b_last_error", C.c_char_p, [_vp]),
    ("b2b_set_option", _int, [_vp, _int, _i64]),
    ("b2b_reserve", _int, [_vp, _u64, _u32]),
    ("b2b_launch_count", _u64, [_vp]),
    ("b2b_kernel_stats", _int, [_vp, _int, C.POINTER(C.c_char_p), C.POINTER(_u64), C.POINTER(C.c_double)]),
    ("b2b_kernel_stats_reset", _int, [_vp]),
    ("b2b_max_frame_size", _sz, [_sz]),
    ("b2b_parse_header", _int, [_vp, _sz, C.POINTER(_CHeader)]),
    ("b2b_header_bytes", None, [C.POINTER(_CHeader), C.POINTER(C.c_uint8 * 16)]),
    ("b2b_compress", _int, [_vp, _vp, _sz, _int, _int, _int, _i64, _vp, _sz, C.POINTER(_sz)]),
    ("b2b_decompress", _int, [_vp, _vp, _sz, _i64, _vp, _sz, C.POINTER(_sz)]),
    ("b2b_shuffle", _int, [_vp, _int, _int, _i64, _vp, _vp, _sz]),
    ("b2b_lz4_bound", _sz, [_sz]),
    ("b2b_lz4_block_compress", _int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    ("b2b_lz4_block_decompress", _int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    ("b2b_compress_batch", _int, [_vp, _vp, _vp, _vp, _u32, _int, _i64, _vp, _u64, _vp, _vp, _vp,
                                  C.POINTER(_u64)]),
    ("b2b_decompress_batch", _int, [_vp, _vp, _vp, _vp, _u32, _i64, _vp, _u64, _vp, _vp, _vp]),
    ("b2b_shuffle_dev", _int, [_vp, _int, _int, _i64, _vp, _vp, _sz, _vp]),
    ("b2b_compress_batch_dev", _int, [_vp, _vp, _vp, _vp, _u32, _u64, _u32, _int, _i64, _vp, _u64,
                                      _vp, _vp, _vp, _vp, _vp]),
    ("b2b_frame_info_batch_dev", _int, [_vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    ("b2b_decompress_batch_dev", _int, [_vp, _vp, _vp, _vp, _u32, _i64, _vp, _vp, _vp, _u64, _u32,
                                        _vp, _vp, _vp]),
    ("b2b_scan_offsets_dev", _int, [_vp, _vp, _u32, _vp, _vp, _vp]),
    ("b2b_allgather_sizes", _int, [_vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp, _int, _vp]),
    ("b2b_blocks_blocksize", _u32, [_sz, _i64, _u32]),
    ("b2b_compress_blocks", _int, [_vp, _vp, _sz, _int, _i64, _u32, _vp, _sz, C.POINTER(_sz)]),
    ("b2b_decompress_blocks", _int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    ("b2b_compress_blocks_batch", _int, [_vp, _vp, _vp, _vp, _u32, _int, _i64, _u32, _vp, _u64, _vp, _vp, _vp,
                                         C.POINTER(_u64)]),
    ("b2b_decompress_blocks_batch", _int, [_vp, _vp, _vp, _vp, _u32, _u32, _vp, _u64, _vp, _vp, _vp]),
    ("b2b_compress_blocks_batch_dev", _int, [_vp, _vp, _vp, _vp, _u32, _u64, _u32, _int, _i64, _u32, _vp, _u64,
                                             _vp, _vp, _vp, _vp, _vp]),
    ("b2b_decompress_blocks_batch_dev", _int, [_vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _u64, _u32, _u32,
                                               _vp, _vp, _vp]),
    ("b2b_index_segments", _u32, [_u32]),
    ("b2b_compress_batch_dev_indexed", _int, [_vp, _vp, _vp, _vp, _u32, _u64, _u32, _int, _i64, _vp, _u64,
                                              _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    ("b2b_decompress_batch_dev_indexed", _int, [_vp, _vp, _vp, _vp, _u32, _i64, _vp, _vp, _vp, _u64, _u32,
                                                _vp, _vp, _vp, _u32, _vp]),
]

_lib = None


def lib():
    """The C-ABI library. Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in ABI:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _raise(status: int, ctx=None):
    msg = lib().b2b_strerror(status).decode()
    if status == ECUDA and ctx is not None:
        msg += ": " + lib().b2b_last_error(ctx).decode()
    raise _ERRORS.get(status, BloscError)(msg)


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data,
                         dtype=np.uint8)


def _np_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data) if a.size else None


def _dev_ptr(t):
    """Device pointer of a torch tensor (or a raw int address)."""
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


# ---- host-only helpers (never touch the GPU) -----------------------------------
def parse_header(data) -> Header:            # blosc.go:165-185
    a = _as_u8(data)
    h = _CHeader()
    rc = lib().b2b_parse_header(_np_ptr(a), a.size, C.byref(h))
    if rc:
        _raise(rc)
    return Header(h.version, h.versionlz, h.flags, h.typesize, h.nbytes_orig, h.blocksize, h.nbytes_comp)


def get_info(data) -> Header:                # blosc.go:306-308
    return parse_header(data)


def get_decompressed_size(data) -> int:      # blosc.go:311-317
    return parse_header(data).nbytes_orig


def max_frame_size(n: int) -> int:
    return int(lib().b2b_max_frame_size(n))


# ---- context ---------------------------------------------------------------------------------
class Context:
    """One b2b_ctx: a CUDA device, a scratch arena and a stream.  Calls are serialised."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        rc = lib().b2b_init(device, C.byref(h))
        if rc:
            raise ErrCuda(f"b2b_init(device={device}) failed: no usable CUDA device "
                          "(this backend has no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().b2b_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option: int, value: int):
        rc = lib().b2b_set_option(self._h, option, value)
        if rc:
            _raise(rc, self._h)

    def reserve(self, total_bytes: int, nframes: int):
        rc = lib().b2b_reserve(self._h, total_bytes, nframes)
        if rc:
            _raise(rc, self._h)

    def launch_count(self) -> int:
        return int(lib().b2b_launch_count(self._h))

    def kernel_stats(self) -> dict:
        """{kernel name: (launches, total device ms)}; times need OPT_KERNEL_TIMING."""
        out = {}
        for k in range(64):
            name, n, ms = C.c_char_p(), C.c_uint64(0), C.c_double(0)
            if lib().b2b_kernel_stats(self._h, k, C.byref(name), C.byref(n), C.byref(ms)) != 0:
                break                                               # B2B_EINVAL beyond the last kernel id
            out[name.value.decode()] = (int(n.value), float(ms.value))
        return out

    def kernel_stats_reset(self):
        lib().b2b_kernel_stats_reset(self._h)

    # -- host-pointer, one frame (compressBackend / decompressBackend) ------------------------
    def compress(self, data, codec=Codec.LZ4, level=5, shuffle=Shuffle.Shuffle1, typesize=4) -> bytes:
        a = _as_u8(data)
        out = np.empty(max_frame_size(a.size) + 64, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = lib().b2b_compress(self._h, _np_ptr(a), a.size, int(codec), int(level), int(shuffle),
                                int(typesize), _np_ptr(out), out.size, C.byref(n))
        if rc:
            _raise(rc, self._h)
        return out[:n.value].tobytes()

    def decompress(self, frame, typesize: int = 0) -> bytes:
        a = _as_u8(frame)
        if a.size < HEADER_SIZE:
            raise ErrInvalidHeader(lib().b2b_strerror(EINVALID_HEADER).decode())   # blosc.go:297-299
        cap = int.from_bytes(a[4:8].tobytes(), "little")
        # like codec.go:78 the output buffer is sized from the header; an LZ4 block cannot
        # expand more than 255x, which bounds what a hostile header can make us allocate
        cap_alloc = min(cap, 255 * a.size + 64)
        out = np.empty(max(cap_alloc, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = lib().b2b_decompress(self._h, _np_ptr(a), a.size, int(typesize), _np_ptr(out), cap_alloc,
                                  C.byref(n))
        if rc == EDST_TOO_SMALL and cap_alloc < cap:
            rc = ESIZE_MISMATCH
        if rc:
            _raise(rc, self._h)
        return out[:n.value].tobytes()

    def shuffle(self, data, typesize: int, mode=Shuffle.Shuffle1, inverse: bool = False) -> np.ndarray:
        a = _as_u8(data)
        out = np.empty_like(a)
        rc = lib().b2b_shuffle(self._h, int(mode), int(bool(inverse)), int(typesize), _np_ptr(a),
                               _np_ptr(out), a.size)
        if rc:
            _raise(rc, self._h)
        return out

    def lz4_block_compress(self, data) -> bytes:
        a = _as_u8(data)
        out = np.empty(int(lib().b2b_lz4_bound(a.size)), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = lib().b2b_lz4_block_compress(self._h, _np_ptr(a), a.size, _np_ptr(out), out.size, C.byref(n))
        if rc:
            _raise(rc, self._h)
        return out[:n.value].tobytes()

    def lz4_block_decompress(self, data, expected_size: int) -> bytes:
        a = _as_u8(data)
        out = np.empty(max(expected_size, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = lib().b2b_lz4_block_decompress(self._h, _np_ptr(a), a.size, _np_ptr(out), expected_size,
                                            C.byref(n))
        if rc:
            _raise(rc, self._h)
        return out[:n.value].tobytes()

    # -- host-pointer batches ------------------------------------------------------------------
    def compress_batch(self, src, src_off, src_len, shuffle=Shuffle.Shuffle1, typesize=4, dst=None):
        """Returns (dst, frame_off, frame_len, status, total)."""
        a = _as_u8(src)
        src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
        src_len = np.ascontiguousarray(src_len, dtype=np.uint32)
        nf = len(src_len)
        span = int((src_off + src_len).max() - src_off.min()) if nf else 0
        if dst is None:
            dst = np.empty(span + 32 * nf + 64, dtype=np.uint8)
        frame_off = np.zeros(nf, dtype=np.uint64)
        frame_len = np.zeros(nf, dtype=np.uint32)
        status = np.zeros(nf, dtype=np.uint32)
        total = C.c_uint64(0)
        rc = lib().b2b_compress_batch(self._h, _np_ptr(a), _np_ptr(src_off), _np_ptr(src_len), nf,
                                      int(shuffle), int(typesize), _np_ptr(dst), dst.size,
                                      _np_ptr(frame_off), _np_ptr(frame_len), _np_ptr(status),
                                      C.byref(total))
        if rc:
            _raise(rc, self._h)
        return dst, frame_off, frame_len, status, total.value

    def decompress_batch(self, frames, frame_off, frame_len, dst_off, dst_total, typesize=0, dst=None):
        """Returns (dst, out_len, status)."""
        a = _as_u8(frames)
        frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
        frame_len = np.ascontiguousarray(frame_len, dtype=np.uint32)
        dst_off = np.ascontiguousarray(dst_off, dtype=np.uint64)
        nf = len(frame_len)
        if dst is None:
            dst = np.empty(max(int(dst_total), 1), dtype=np.uint8)
        out_len = np.zeros(nf, dtype=np.uint32)
        status = np.zeros(nf, dtype=np.uint32)
        rc = lib().b2b_decompress_batch(self._h, _np_ptr(a), _np_ptr(frame_off), _np_ptr(frame_len), nf,
                                        int(typesize), _np_ptr(dst), int(dst_total), _np_ptr(dst_off),
                                        _np_ptr(out_len), _np_ptr(status))
        if rc:
            _raise(rc, self._h)
        return dst, out_len, status

    # -- device-pointer entry points (torch tensors or raw addresses; stream = cudaStream_t) ---
    def shuffle_dev(self, mode, inverse, typesize, d_src, d_dst, n, stream=0):
        rc = lib().b2b_shuffle_dev(self._h, int(mode), int(bool(inverse)), int(typesize), _dev_ptr(d_src),
                                   _dev_ptr(d_dst), n, C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    def compress_batch_dev(self, d_src, d_src_off, d_src_len, nframes, total_src_bytes, max_frame_len,
                           shuffle, typesize, d_dst, dst_cap, d_frame_off, d_frame_len, d_status,
                           d_total_out, stream=0):
        rc = lib().b2b_compress_batch_dev(self._h, _dev_ptr(d_src), _dev_ptr(d_src_off),
                                          _dev_ptr(d_src_len), nframes, total_src_bytes, max_frame_len,
                                          int(shuffle), int(typesize), _dev_ptr(d_dst), dst_cap,
                                          _dev_ptr(d_frame_off), _dev_ptr(d_frame_len),
                                          _dev_ptr(d_status), _dev_ptr(d_total_out), C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    def frame_info_batch_dev(self, d_frames, d_frame_off, d_frame_len, nframes, d_orig_len, d_dst_off,
                             d_total, d_status, stream=0):
        rc = lib().b2b_frame_info_batch_dev(self._h, _dev_ptr(d_frames), _dev_ptr(d_frame_off),
                                            _dev_ptr(d_frame_len), nframes, _dev_ptr(d_orig_len),
                                            _dev_ptr(d_dst_off), _dev_ptr(d_total), _dev_ptr(d_status),
                                            C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    def decompress_batch_dev(self, d_frames, d_frame_off, d_frame_len, nframes, typesize_override, d_dst,
                             d_dst_off, d_dst_cap, total_dst_bytes, max_orig_len, d_out_len, d_status,
                             stream=0):
        rc = lib().b2b_decompress_batch_dev(self._h, _dev_ptr(d_frames), _dev_ptr(d_frame_off),
                                            _dev_ptr(d_frame_len), nframes, int(typesize_override),
                                            _dev_ptr(d_dst), _dev_ptr(d_dst_off), _dev_ptr(d_dst_cap),
                                            total_dst_bytes, max_orig_len, _dev_ptr(d_out_len),
                                            _dev_ptr(d_status), C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    # ---- side-car decode index (b2b.h): standard frames, independent 64 KiB segments, an index of
    # sequence boundaries outside the frames; decode with one warp per index entry
    @staticmethod
    def index_segments(max_frame_len: int) -> int:
        return int(lib().b2b_index_segments(int(max_frame_len)))

    def compress_batch_dev_indexed(self, d_src, d_src_off, d_src_len, nframes, total_src_bytes, max_frame_len,
                                   shuffle, typesize, d_dst, dst_cap, d_frame_off, d_frame_len, d_status,
                                   d_total_out, d_index, segs_per_frame, stream=0):
        rc = lib().b2b_compress_batch_dev_indexed(self._h, _dev_ptr(d_src), _dev_ptr(d_src_off),
                                                  _dev_ptr(d_src_len), nframes, total_src_bytes, max_frame_len,
                                                  int(shuffle), int(typesize), _dev_ptr(d_dst), dst_cap,
                                                  _dev_ptr(d_frame_off), _dev_ptr(d_frame_len),
                                                  _dev_ptr(d_status), _dev_ptr(d_total_out), _dev_ptr(d_index),
                                                  int(segs_per_frame), C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    def decompress_batch_dev_indexed(self, d_frames, d_frame_off, d_frame_len, nframes, typesize_override, d_dst,
                                     d_dst_off, d_dst_cap, total_dst_bytes, max_orig_len, d_out_len, d_status,
                                     d_index, segs_per_frame, stream=0):
        rc = lib().b2b_decompress_batch_dev_indexed(self._h, _dev_ptr(d_frames), _dev_ptr(d_frame_off),
                                                    _dev_ptr(d_frame_len), nframes, int(typesize_override),
                                                    _dev_ptr(d_dst), _dev_ptr(d_dst_off), _dev_ptr(d_dst_cap),
                                                    total_dst_bytes, max_orig_len, _dev_ptr(d_out_len),
                                                    _dev_ptr(d_status), _dev_ptr(d_index), int(segs_per_frame),
                                                    C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    # ---- opt-in Blosc-1 multi-block frames (b2b.h; the reference ignores Options.BlockSize and
    # cannot read these): int32 bstarts + one LZ4 block per block, filter per block
    @staticmethod
    def blocks_blocksize(n: int, typesize: int, blocksize: int = 0) -> int:
        return int(lib().b2b_blocks_blocksize(int(n), int(typesize), int(blocksize)))

    def compress_blocks(self, data, shuffle=Shuffle.Shuffle1, typesize=4, blocksize=0) -> bytes:
        a = _as_u8(data)
        out = np.empty(a.size + 16 + 64, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = lib().b2b_compress_blocks(self._h, _np_ptr(a), a.size, int(shuffle), int(typesize), int(blocksize),
                                       _np_ptr(out), out.size, C.byref(n))
        if rc:
            _raise(rc, self._h)
        return out[:n.value].tobytes()

    def decompress_blocks(self, frame) -> bytes:
        a = _as_u8(frame)
        if a.size < HEADER_SIZE:
            raise ErrInvalidHeader(lib().b2b_strerror(EINVALID_HEADER).decode())
        cap = int.from_bytes(a[4:8].tobytes(), "little")
        cap_alloc = min(cap, 255 * a.size + 64)          # an LZ4 block expands at most 255x
        out = np.empty(max(cap_alloc, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        rc = lib().b2b_decompress_blocks(self._h, _np_ptr(a), a.size, _np_ptr(out), cap_alloc, C.byref(n))
        if rc == EDST_TOO_SMALL and cap_alloc < cap:
            rc = EDECOMPRESSION_FAILED
        if rc:
            _raise(rc, self._h)
        return out[:n.value].tobytes()

    def compress_blocks_batch(self, src, src_off, src_len, shuffle=Shuffle.Shuffle1, typesize=4, blocksize=0, dst=None):
        """Host-pointer batch of multi-block frames. Returns (dst, frame_off, frame_len, status, total)."""
        a = _as_u8(src)
        src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
        src_len = np.ascontiguousarray(src_len, dtype=np.uint32)
        nf = len(src_len)
        span = int((src_off + src_len).max() - src_off.min()) if nf else 0
        if dst is None:
            dst = np.empty(span + 32 * nf + 64, dtype=np.uint8)
        frame_off = np.zeros(nf, dtype=np.uint64)
        frame_len = np.zeros(nf, dtype=np.uint32)
        status = np.zeros(nf, dtype=np.uint32)
        total = C.c_uint64(0)
        rc = lib().b2b_compress_blocks_batch(self._h, _np_ptr(a), _np_ptr(src_off), _np_ptr(src_len), nf,
                                             int(shuffle), int(typesize), int(blocksize), _np_ptr(dst), dst.size,
                                             _np_ptr(frame_off), _np_ptr(frame_len), _np_ptr(status), C.byref(total))
        if rc:
            _raise(rc, self._h)
        return dst, frame_off, frame_len, status, total.value

    def decompress_blocks_batch(self, frames, frame_off, frame_len, dst_off, dst_total, blocksize=0, dst=None):
        """Returns (dst, out_len, status)."""
        a = _as_u8(frames)
        frame_off = np.ascontiguousarray(frame_off, dtype=np.uint64)
        frame_len = np.ascontiguousarray(frame_len, dtype=np.uint32)
        dst_off = np.ascontiguousarray(dst_off, dtype=np.uint64)
        nf = len(frame_len)
        if dst is None:
            dst = np.empty(max(int(dst_total), 1), dtype=np.uint8)
        out_len = np.zeros(nf, dtype=np.uint32)
        status = np.zeros(nf, dtype=np.uint32)
        rc = lib().b2b_decompress_blocks_batch(self._h, _np_ptr(a), _np_ptr(frame_off), _np_ptr(frame_len), nf,
                                               int(blocksize), _np_ptr(dst), int(dst_total), _np_ptr(dst_off),
                                               _np_ptr(out_len), _np_ptr(status))
        if rc:
            _raise(rc, self._h)
        return dst, out_len, status

    def compress_blocks_batch_dev(self, d_src, d_src_off, d_src_len, nframes, total_src_bytes, max_frame_len,
                                  shuffle, typesize, blocksize, d_dst, dst_cap, d_frame_off, d_frame_len,
                                  d_status, d_total_out, stream=0):
        rc = lib().b2b_compress_blocks_batch_dev(self._h, _dev_ptr(d_src), _dev_ptr(d_src_off),
                                                 _dev_ptr(d_src_len), nframes, total_src_bytes, max_frame_len,
                                                 int(shuffle), int(typesize), int(blocksize), _dev_ptr(d_dst),
                                                 dst_cap, _dev_ptr(d_frame_off), _dev_ptr(d_frame_len),
                                                 _dev_ptr(d_status), _dev_ptr(d_total_out), C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    def decompress_blocks_batch_dev(self, d_frames, d_frame_off, d_frame_len, nframes, d_dst, d_dst_off,
                                    d_dst_cap, total_dst_bytes, max_orig_len, blocksize, d_out_len, d_status,
                                    stream=0):
        rc = lib().b2b_decompress_blocks_batch_dev(self._h, _dev_ptr(d_frames), _dev_ptr(d_frame_off),
                                                   _dev_ptr(d_frame_len), nframes, _dev_ptr(d_dst),
                                                   _dev_ptr(d_dst_off), _dev_ptr(d_dst_cap), total_dst_bytes,
                                                   max_orig_len, int(blocksize), _dev_ptr(d_out_len),
                                                   _dev_ptr(d_status), C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    def allgather_sizes(self, nccl_comm, d_local_len, n_local, world, d_all_len, d_all_off, d_total, align16=False, stream=0):
        """One ncclAllGather of the per-frame sizes + the K5 scan, all on `stream` (include/b2b.h)."""
        rc = lib().b2b_allgather_sizes(self._h, C.c_void_p(nccl_comm), _dev_ptr(d_local_len), n_local, world,
                                       _dev_ptr(d_all_len), _dev_ptr(d_all_off), _dev_ptr(d_total), int(bool(align16)),
                                       C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)

    def scan_offsets_dev(self, d_len, n, d_off, d_total, stream=0):
        rc = lib().b2b_scan_offsets_dev(self._h, _dev_ptr(d_len), n, _dev_ptr(d_off), _dev_ptr(d_total),
                                        C.c_void_p(stream))
        if rc:
            _raise(rc, self._h)


# ---- package-level API with the reference's signatures (uses a lazily created default ctx) ----
_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def compress_with_options(data, opts: Options) -> bytes:     # blosc.go:268-286
    if len(data) == 0:
        raise ErrInvalidData(lib().b2b_strerror(EINVALID_DATA).decode())   # bare sentinel, before any device work
    level = min(max(opts.level, 1), 9)
    return default_context().compress(data, opts.codec, level, opts.shuffle, opts.typesize)


def compress(data, codec=Codec.LZ4, level=5, shuffle=Shuffle.Shuffle1, typesize=4) -> bytes:  # blosc.go:257-265
    return compress_with_options(data, Options(codec=codec, level=level, shuffle=shuffle, typesize=typesize))


def decompress_with_size(data, typesize: int) -> bytes:     # blosc.go:296-303
    return default_context().decompress(data, typesize)


def decompress(data) -> bytes:                               # blosc.go:291-293
    return decompress_with_size(data, 0)


def shuffle_buffer(data, typesize: int, mode) -> None:       # shuffle.go:298-309 (in place)
    _filter_in_place(data, typesize, mode, False)


def unshuffle_buffer(data, typesize: int, mode) -> None:     # shuffle.go:312-323 (in place)
    _filter_in_place(data, typesize, mode, True)


def _filter_in_place(data, typesize, mode, inverse):
    if int(mode) not in (Shuffle.Shuffle1, Shuffle.BitShuffle):
        return                                               # default arm: untouched
    view = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
    if view.size == 0:
        return
    out = default_context().shuffle(view, typesize, mode, inverse)
    view[:] = out
