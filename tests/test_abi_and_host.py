"""CPU: the C-ABI library loads, exports every symbol include/b2b.h declares, its host-only
entry points (headers, sizes, error strings) behave like the reference, and it FAILS LOUDLY
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "b2b.h")).read()
    declared = set(re.findall(r"B2B_API\s+[\w\s\*]+?\b(b2b_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    L = ctypes.CDLL(pkg.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/b2b.h but not exported"
    bound = {n for n, _, _ in pkg.ABI}
    assert declared == bound, f"python binding and header disagree: {declared ^ bound}"


def test_library_is_self_contained_native_code(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, "libb2b.so must carry sm_100a SASS"


def test_header_roundtrip_and_layout(pkg):
    """blosc.go:165-198; fuzz_test.go:366-424 (field-by-field little-endian layout)."""
    raw = struct.pack("<BBBBIII", 2, 1, 5, 8, 0x01020304, 0x0A0B0C0D, 0x11223344)
    h = pkg.parse_header(raw)
    assert (h.version, h.versionlz, h.flags, h.typesize) == (2, 1, 5, 8)
    assert (h.nbytes_orig, h.blocksize, h.nbytes_comp) == (0x01020304, 0x0A0B0C0D, 0x11223344)
    assert h.to_bytes() == raw
    assert h.has_shuffle() and h.has_bitshuffle() and not h.is_memcpy()
    assert h.shuffle_mode() == pkg.Shuffle.BitShuffle               # blosc_test.go:457-478: bitshuffle wins
    assert pkg.Header(2, 1, 1, 4, 0, 0, 16).shuffle_mode() == pkg.Shuffle.Shuffle1
    assert pkg.Header(2, 1, 2, 4, 0, 0, 16).shuffle_mode() == pkg.Shuffle.NoShuffle
    assert pkg.get_info(raw) == h and pkg.get_decompressed_size(raw) == 0x01020304
    assert pkg.parse_header(raw + b"trailing").nbytes_comp == 0x11223344
    with pytest.raises(pkg.ErrInvalidHeader):
        pkg.parse_header(raw[:15])
    with pytest.raises(pkg.ErrInvalidHeader):
        pkg.parse_header(b"")
    for v in (0, 1, 3, 99, 255):
        with pytest.raises(pkg.ErrInvalidVersion):
            pkg.parse_header(bytes([v]) + raw[1:])


def test_enums_strings_defaults(pkg):
    """blosc.go:55-107, 237-245; blosc_test.go:314-349, 419-435, 480-498."""
    assert [int(c) for c in pkg.Codec] == [0, 1, 2, 3, 4, 5]
    assert [str(c) for c in pkg.Codec] == ["blosclz", "lz4", "lz4hc", "snappy", "zlib", "zstd"]
    assert [str(s) for s in pkg.Shuffle] == ["noshuffle", "shuffle", "bitshuffle"]
    o = pkg.default_options()
    assert (o.codec, o.level, o.shuffle, o.typesize, o.blocksize) == (pkg.Codec.LZ4, 5, pkg.Shuffle.Shuffle1, 4, 0)
    assert (pkg.VERSION, pkg.FORMAT_VERSION, pkg.HEADER_SIZE, pkg.MIN_HEADER_SIZE) == ("1.0.0", 2, 16, 16)
    assert pkg.max_frame_size(1000) == 1016
    assert pkg.lib().b2b_lz4_bound(1000) == 1000 + 3 + 16


def test_error_strings_match_reference(pkg):
    """blosc.go:125-149."""
    want = {1: "blosc: invalid compressed data", 2: "blosc: invalid header",
            3: "blosc: unsupported format version", 4: "blosc: unsupported codec",
            5: "blosc: decompressed size mismatch", 6: "blosc: data too large",
            7: "blosc: compression failed", 8: "blosc: decompression failed"}
    for code, msg in want.items():
        assert pkg.lib().b2b_strerror(code).decode() == msg


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ErrCuda):
        pkg.Context(0)
    with pytest.raises(pkg.ErrCuda):
        pkg.compress(b"abcd" * 100)
    with pytest.raises(pkg.ErrInvalidData):          # argument checks come first, like the reference
        pkg.compress(b"")


def test_product_does_not_touch_the_oracle():
    """The product path must never import/link oracle/ (tests, smoke and bench only)."""
    bad = []
    walk = list(os.walk(os.path.join(ROOT, "go-blosc_b200"))) + list(os.walk(os.path.join(ROOT, "tools")))   # probes that use the checker live under tests/tools/
    for base, _, files in walk:
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".go")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle[/.]|liboracle|blosc_oracle", txt):
                    bad.append(f)
    assert not bad, bad
    assert "oracle" not in open(os.path.join(ROOT, "include", "b2b.h")).read().lower()


def test_host_kernel_check_builds_and_passes():
    """Per-thread permutation code of the kernels, executed on the CPU, equals the oracle."""
    exe = os.path.join(ROOT, "tests", "_build", "host_kernel_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "host_kernel_check.cu"),
                           os.path.join(ROOT, "oracle", "blosc_oracle.c"), "-Xcompiler", "-pthread"],
                          stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


def test_cpp_host_mirror_abi_only(pkg):
    """C++ host mirror, host-only entry points (no device): header, strings, argument errors."""
    exe = os.path.join(ROOT, "go-blosc_b200", "lib", "blosc_host_test")
    out = subprocess.run([exe, "--abi"], capture_output=True, text=True)
    assert out.returncode == 0 and "abi: ok" in out.stdout, out.stdout


def test_go_package_mirrors_every_exported_identifier():
    """the Go host package keeps the reference's exported surface."""
    go = ""
    d = os.path.join(ROOT, "go-blosc_b200", "go", "blosc")
    for f in os.listdir(d):
        go += open(os.path.join(d, f)).read()
    for ident in ["func Compress(", "func CompressWithOptions(", "func Decompress(", "func DecompressWithSize(",
                  "func GetInfo(", "func GetDecompressedSize(", "func ParseHeader(", "func DefaultOptions(",
                  "func ShuffleBuffer(", "func UnshuffleBuffer(", "func RegisterCodec(", "func GetCodec(",
                  "func ListCodecs(", "type CodecInterface interface", "type Codec uint8", "type Shuffle uint8",
                  "type Options struct", "type Header struct", "FormatVersion", "HeaderSize", "MinHeaderSize",
                  "ErrInvalidData", "ErrInvalidHeader", "ErrInvalidVersion", "ErrInvalidCodec", "ErrSizeMismatch",
                  "ErrDataTooLarge", "ErrCompressionFailed", "ErrDecompressionFailed", "HasShuffle", "HasBitShuffle",
                  "IsMemcpy", "ShuffleMode", "func (h *Header) Bytes("]:
        assert ident in go, ident
    for sym in re.findall(r"C\.(b2b_\w+)\(", go):
        assert sym in open(os.path.join(ROOT, "include", "b2b.h")).read(), sym
