"""The reference's own test assertions for the shuffle + LZ4 path, restated (file:line cited)
as functions of an implementation adapter (tests/adapters.py).  Run against the oracle on the
CPU (test_oracle_pins.py) and against the CUDA path on the GPU (test_gpu_parity.py)."""
import struct

import numpy as np
import pytest

import datagen as dg

LZ4, LZ4HC, SNAPPY, ZLIB, ZSTD = 1, 2, 3, 4, 5
NOSHUFFLE, SHUFFLE, BITSHUFFLE = 0, 1, 2


def hdr(frame):
    v, c, fl, ts, no, bs, nc = struct.unpack("<BBBBIII", bytes(frame[:16]))
    return dict(version=v, codec=c, flags=fl, typesize=ts, norig=no, blocksize=bs, ncomp=nc)


def make_header(version=2, codec=LZ4, flags=0, typesize=4, norig=0, blocksize=None, ncomp=16):
    return struct.pack("<BBBBIII", version, codec, flags, typesize, norig, norig if blocksize is None else blocksize, ncomp)


# ---- shuffle_test.go -------------------------------------------------------------------------
def case_shuffle_formula(impl):
    """shuffle_amd64_test.go:47-61,152-174: dst[j*E+i] == src[i*T+j]; shuffle.go:14-15 doc example."""
    for n in (32, 64, 128, 1000):
        src = dg.ramp(n)
        E = n // 4
        want = src[:E * 4].reshape(E, 4).T.reshape(-1)
        got = impl.shuffle(src, 4, SHUFFLE)
        assert np.array_equal(got[:E * 4], want)
    doc = np.frombuffer(b"\xA0\xA1\xA2\xA3\xB0\xB1\xB2\xB3\xC0\xC1\xC2\xC3", dtype=np.uint8)
    assert impl.shuffle(doc, 4, SHUFFLE).tobytes() == b"\xA0\xB0\xC0\xA1\xB1\xC1\xA2\xB2\xC2\xA3\xB3\xC3"


def case_shuffle_roundtrips(impl):
    """shuffle_test.go:13-40 (T in 1,2,4,8,16), 63-91 (bitshuffle T in 2,4,8), 146-168 and 410-435
    (remainders), 284-316 and 382-408 (partial groups), 364-380 (group boundaries), 170-184 (length)."""
    cases = [(1024, t, m) for t in (1, 2, 4, 8, 16) for m in (SHUFFLE, BITSHUFFLE)]
    cases += [(1003, 4, SHUFFLE), (13, 4, SHUFFLE), (103, 8, SHUFFLE), (10, 4, SHUFFLE)]
    cases += [(n, 4, BITSHUFFLE) for n in (28, 35, 12, 37, 97)] + [(127, 8, BITSHUFFLE)]
    cases += [(e * 4, 4, BITSHUFFLE) for e in range(8, 65, 8)]
    for n, t, m in cases:
        src = dg.ramp(n)
        sh = impl.shuffle(src, t, m)
        assert len(sh) == n
        back = impl.shuffle(sh, t, m, inverse=True)
        assert np.array_equal(back, src), (n, t, m)


def case_shuffle_noops(impl):
    """shuffle_test.go:113-144, 318-362, 437-452: T=1, NoShuffle, len<T and unknown modes are identity."""
    src = dg.ramp(100)
    assert np.array_equal(impl.shuffle(src, 1, SHUFFLE), src)
    assert np.array_equal(impl.shuffle(src, 1, BITSHUFFLE), src)
    assert np.array_equal(impl.shuffle(src, 4, NOSHUFFLE), src)
    assert np.array_equal(impl.shuffle(src, 4, 99), src)
    small = dg.ramp(3)
    assert np.array_equal(impl.shuffle(small, 4, SHUFFLE), small)
    assert np.array_equal(impl.shuffle(small, 8, BITSHUFFLE), small)
    assert np.array_equal(impl.shuffle(small, 4, SHUFFLE, inverse=True), small)


# ---- blosc_test.go / codec_test.go / example_test.go -------------------------------------------
def case_roundtrip_basic(impl):
    """blosc_test.go:13-105 (LZ4 arm), 107-134 (float32*0.1 + Shuffle), 136-163 (float64*0.1 + BitShuffle)."""
    for data, sh, ts in ((dg.ramp(1000), NOSHUFFLE, 1), (dg.f32_ramp(1000), SHUFFLE, 4), (dg.f64_ramp(1000), BITSHUFFLE, 8)):
        fr = impl.compress(data, LZ4, 5, sh, ts)
        assert impl.decompress(fr) == data.tobytes()


def case_header_fields(impl):
    """blosc_test.go:165-192; example_test.go:136-148; README quick start."""
    fr = impl.compress(dg.ramp(1000), LZ4, 5, SHUFFLE, 4)
    h = hdr(fr)
    assert (h["version"], h["typesize"], h["norig"], h["codec"]) == (2, 4, 1000, LZ4)
    assert h["flags"] & 0x1 and not h["flags"] & 0x4
    assert h["blocksize"] == 1000 and h["ncomp"] == len(fr)       # blosc.go:364-365
    fr = impl.compress(dg.ramp(10000), LZ4, 5, SHUFFLE, 4)
    h = hdr(fr)
    assert (h["version"], h["codec"], h["norig"], h["typesize"]) == (2, 1, 10000, 4) and h["flags"] & 1
    assert len(fr) < 10000                                          # "Compressed smaller: true"
    fr = impl.compress(dg.ramp(100000), LZ4, 5, SHUFFLE, 4)         # BASELINE config C1
    assert fr[:12] == bytes.fromhex("02010104a0860100a0860100")


def case_errors(impl, pkg):
    """Sentinels: blosc_test.go:211-241, 437-455; codec_test.go:37-79, 203-233, 275-295, 452-470."""
    with pytest.raises(pkg.ErrInvalidData):
        impl.compress(b"", LZ4, 5, SHUFFLE, 4)
    with pytest.raises(pkg.ErrInvalidHeader):
        impl.decompress(b"\x01\x02\x03")
    with pytest.raises(pkg.ErrInvalidHeader):
        impl.decompress(b"")
    fr = bytearray(impl.compress(dg.ramp(1000), LZ4, 5, NOSHUFFLE, 1))
    bad = bytearray(fr); bad[0] = 99
    with pytest.raises(pkg.ErrInvalidVersion):
        impl.decompress(bytes(bad))
    for v in (0, 1, 3):
        with pytest.raises(pkg.ErrInvalidVersion):
            impl.decompress(make_header(version=v, norig=100, ncomp=116) + bytes(100))
    with pytest.raises(pkg.ErrInvalidCodec):                         # codec_test.go:37-58
        impl.decompress(make_header(codec=99, norig=100, ncomp=66) + bytes(50))
    with pytest.raises(pkg.ErrInvalidCodec):                         # BloscLZ is declared, never registered
        impl.decompress(make_header(codec=0, norig=100, ncomp=66) + bytes(50))
    with pytest.raises(pkg.ErrInvalidCodec):
        impl.compress(dg.ramp(100), 0, 5, SHUFFLE, 4)
    with pytest.raises(pkg.ErrInvalidCodec):
        impl.compress(dg.ramp(100), 99, 5, SHUFFLE, 4)
    mism = bytearray(fr); mism[4:8] = struct.pack("<I", 2000)      # codec_test.go:60-79
    with pytest.raises(pkg.ErrSizeMismatch):
        impl.decompress(bytes(mism))
    with pytest.raises(pkg.ErrInvalidData):                          # codec_test.go:452-470
        impl.decompress(make_header(flags=0x2, norig=100, ncomp=1000) + bytes(10))
    with pytest.raises(pkg.ErrInvalidData):                          # fuzz seed: NBytesComp beyond the data
        impl.decompress(make_header(norig=1000, ncomp=1000))
    with pytest.raises(pkg.ErrInvalidData):                          # NBytesComp < 16
        impl.decompress(make_header(norig=10, ncomp=8) + bytes(10))
    with pytest.raises(pkg.BloscError):                              # codec_test.go:276-284
        impl.decompress(make_header(norig=100, ncomp=20) + b"\xff\xff\xff\xff")
    corrupt = bytearray(fr)                                          # blosc_test.go:593-611
    for i in range(16, len(corrupt)):
        corrupt[i] ^= 0xFF
    with pytest.raises(pkg.BloscError):
        impl.decompress(bytes(corrupt))
    with pytest.raises(pkg.ErrSizeMismatch):                         # memcpy payload shorter than NBytesOrig
        impl.decompress(make_header(flags=0x2, norig=100, ncomp=26) + bytes(10))


def case_memcpy_path(impl):
    """blosc_test.go:243-266, 560-591, 657-681, 764-800: incompressible data, NoShuffle -> memcpy frame."""
    data = dg.random_bytes(1000, 11)
    fr = impl.compress(data, LZ4, 1, NOSHUFFLE, 1)
    h = hdr(fr)
    assert h["flags"] & 0x2 and h["ncomp"] == 1016 and fr[16:] == data.tobytes()
    assert impl.decompress(fr) == data.tobytes()
    tiny = b"ab"                                                      # shorter than any LZ4 gain
    assert impl.decompress(impl.compress(tiny, LZ4, 5, NOSHUFFLE, 1)) == tiny


def case_typesize_shuffle_matrix(impl):
    """blosc_test.go:290-312: ramp(1024), T in 1,2,4,8,16 x {NoShuffle, Shuffle, BitShuffle}."""
    data = dg.ramp(1024)
    for ts in (1, 2, 4, 8, 16):
        for sh in (NOSHUFFLE, SHUFFLE, BITSHUFFLE):
            fr = impl.compress(data, LZ4, 5, sh, ts)
            h = hdr(fr)
            assert h["typesize"] == ts
            assert (h["flags"] & 0x5) == {NOSHUFFLE: 0, SHUFFLE: 1, BITSHUFFLE: 4}[sh]   # blosc.go:348-353
            assert impl.decompress(fr) == data.tobytes(), (ts, sh)


def case_clamping_and_override(impl):
    """blosc_test.go:613-655 (levels/typesizes out of range do not error), 683-719 (override)."""
    data = dg.ramp(1000)
    for level in (-1, 0, 1, 5, 9, 10, 100):                          # fuzz_test.go:255-266
        assert impl.decompress(impl.compress(data, LZ4, level, NOSHUFFLE, 1)) == data.tobytes()
    for ts in (-1, 0):
        fr = impl.compress(data, LZ4, 5, SHUFFLE, ts)                # TypeSize <= 0 -> 1 (blosc.go:274-276)
        assert hdr(fr)["typesize"] == 1 and impl.decompress(fr) == data.tobytes()
    for ts in (3, 7, 16, 32):                                        # fuzz_test.go:239-274: must not fail
        fr = impl.compress(data, LZ4, 5, SHUFFLE, ts)
        assert impl.decompress(fr) == data.tobytes()
    f = dg.f32_ramp(250)
    fr = impl.compress(f, LZ4, 5, SHUFFLE, 4)
    assert impl.decompress(fr, 4) == f.tobytes() and impl.decompress(fr, 0) == f.tobytes()
    # an override that differs from the header changes the unshuffle (blosc.go:417-426)
    assert impl.decompress(fr, 2) != f.tobytes()


def case_shuffle_improves(impl):
    """example_test.go:208-231: the 4000-byte correlated pattern compresses better with Shuffle."""
    i = np.arange(0, 4000, 4)
    data = np.zeros(4000, dtype=np.uint8)
    data[0::4] = (i // 100) & 0xFF; data[1::4] = (i // 50) & 0xFF
    data[2::4] = (i // 10) & 0xFF; data[3::4] = i & 0xFF
    assert len(impl.compress(data, LZ4, 5, SHUFFLE, 4)) < len(impl.compress(data, LZ4, 5, NOSHUFFLE, 4))


def case_fuzz_seeds(impl, pkg):
    """fuzz_test.go:25-159: hostile headers never crash; a success implies len(out) == NBytesOrig."""
    seeds = [b"", b"\x02", b"\x02\x01", b"\x02\x01\x00\x04",
             make_header(version=99, norig=100, ncomp=116), make_header(version=0), make_header(version=1),
             make_header(norig=1000, ncomp=1000),
             make_header(flags=0x2, norig=100, ncomp=26) + bytes(10),
             make_header(codec=255, flags=0, typesize=1, norig=50, ncomp=66) + bytes(50),
             make_header(norig=0, ncomp=16),
             make_header(norig=0xFFFFFFFF, blocksize=0xFFFFFFFF, ncomp=0xFFFFFFFF)]
    for ts in (0, 1, 2, 4, 8, 16, 255):
        seeds.append(make_header(flags=0x1, typesize=ts, norig=20, ncomp=36) + bytes(20))
    seeds.append(make_header(flags=0x4, typesize=4, norig=20, ncomp=36) + bytes(20))
    seeds.append(make_header(flags=0xFF, typesize=4, norig=20, ncomp=36) + bytes(20))
    for s in seeds:
        try:
            out = impl.decompress(s)
        except pkg.BloscError:
            continue
        assert len(out) == hdr(s)["norig"], s.hex()
    # zero-length original with an empty payload decodes to nothing (pierrec: empty src -> 0 bytes)
    assert impl.decompress(make_header(norig=0, ncomp=16)) == b""
    # the all-flags frame is a memcpy frame (0x2 set): 20 zero bytes, bit-unshuffled -> zeros
    assert impl.decompress(make_header(flags=0xFF, typesize=4, norig=20, ncomp=36) + bytes(20)) == bytes(20)


def case_memcpy_shuffle_policy(impl):
    """SURVEY F4 / DESIGN.md: incompressible input with a shuffle flag.  Default policy stores the
    shuffled bytes (round-trips through the reference decoder); the quirk policy stores the
    original bytes exactly like blosc.go:342-345 (and then does not round-trip, like the reference)."""
    data = dg.random_bytes(1000, 5)
    fr = impl.compress(data, LZ4, 5, SHUFFLE, 2)
    assert hdr(fr)["flags"] == 0x3 and impl.decompress(fr) == data.tobytes()
    q = impl.compress(data, LZ4, 5, SHUFFLE, 2, quirk=True)
    assert q[:16] == fr[:16] and q[16:] == data.tobytes()
    assert impl.decompress(q) != data.tobytes()


ALL_CASES = [case_shuffle_formula, case_shuffle_roundtrips, case_shuffle_noops, case_roundtrip_basic,
             case_header_fields, case_memcpy_path, case_typesize_shuffle_matrix,
             case_clamping_and_override, case_shuffle_improves, case_memcpy_shuffle_policy]
CASES_WITH_PKG = [case_errors, case_fuzz_seeds]
