"""CPU: pin the oracle.  (1) every assertion the reference's own tests hold for this path
(reference_suite.py, file:line cited there) must hold for the oracle; (2) the oracle's filters
must equal an independent pure-Python restatement of shuffle.go's loops; (3) its LZ4 streams
must be valid for -- and its decoder must agree with -- the system liblz4 (format referee);
(4) it must reproduce the committed golden fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

import datagen as dg
import reference_suite as rs
from adapters import OracleImpl

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.json")


@pytest.fixture(scope="module")
def impl(orc, pkg):
    return OracleImpl(orc, pkg)


@pytest.mark.parametrize("case", rs.ALL_CASES, ids=lambda c: c.__name__)
def test_reference_case(case, impl):
    case(impl)


@pytest.mark.parametrize("case", rs.CASES_WITH_PKG, ids=lambda c: c.__name__)
def test_reference_case_errors(case, impl, pkg):
    case(impl, pkg)


# ---- independent restatement of the scalar Go loops (slow, small sizes only) -------------------
def py_shuffle(src, T):            # shuffle.go:16-73
    n = len(src)
    if T <= 1 or n < T:
        return bytes(src)
    E = n // T
    dst = bytearray(n)
    for i in range(E):
        for j in range(T):
            dst[j * E + i] = src[i * T + j]
    dst[E * T:] = src[E * T:]
    return bytes(dst)


def py_bitshuffle(src, T, inverse=False):   # shuffle.go:145-295
    n = len(src)
    if T <= 1 or n < T:
        return bytes(src)
    E = n // T
    dst = bytearray(src)             # leftovers / tail stay raw
    for g in range(E // 8):
        base = g * 8 * T
        for j in range(T):
            if not inverse:
                b = [src[base + m * T + j] for m in range(8)]
                for k in range(8):
                    o = 0
                    for m in range(8):
                        if b[m] & (1 << (7 - k)):
                            o |= 1 << (7 - m)
                    dst[base + j * 8 + k] = o
            else:
                s = [src[base + j * 8 + k] for k in range(8)]
                for e in range(8):
                    o = 0
                    for k in range(8):
                        if s[k] & (1 << (7 - e)):
                            o |= 1 << (7 - k)
                    dst[base + e * T + j] = o
    return bytes(dst)


@pytest.mark.parametrize("n", [1, 3, 10, 13, 28, 35, 64, 97, 127, 256, 1003])
@pytest.mark.parametrize("T", [1, 2, 3, 4, 7, 8, 16, 17])
def test_filters_match_python_restatement(orc, n, T):
    src = dg.lcg_bytes(n, seed=n * 31 + T)
    b = bytes(src)
    assert orc.shuffle(src, T).tobytes() == py_shuffle(b, T)
    assert orc.bitshuffle(src, T).tobytes() == py_bitshuffle(b, T)
    assert orc.bitunshuffle(src, T).tobytes() == py_bitshuffle(b, T, inverse=True)
    assert orc.unshuffle(orc.shuffle(src, T), T).tobytes() == b
    assert orc.shuffle_fast(src, T).tobytes() == py_shuffle(b, T)
    assert orc.unshuffle_fast(orc.shuffle(src, T), T).tobytes() == b


def test_derived_vectors(orc):
    """literals derived from the Go loops by hand."""
    r = dg.ramp
    assert orc.shuffle(r(16), 4).tobytes().hex() == "0004080c0105090d02060a0e03070b0f"
    assert orc.shuffle(r(11), 4).tobytes().hex() == "000401050206030708090a"
    assert orc.bitshuffle(r(16), 2).tobytes().hex() == "000000000f335500000000000f3355ff"
    assert orc.bitshuffle(r(22), 2).tobytes().hex() == "000000000f335500000000000f3355ff101112131415"
    assert orc.bitshuffle(r(32), 4).tobytes().hex() == ("0000000f335500000000000f335500ff"
                                                        "0000000f3355ff000000000f3355ffff")
    one = np.array([0x80, 0, 0, 0, 0, 0, 0, 0], dtype=np.uint8)
    assert orc.bitshuffle(one, 1).tobytes() == one.tobytes()          # T=1: identity
    col = np.zeros(16, dtype=np.uint8); col[0] = 0xFF                 # T=2: b=[ff 00 ...] for j=0
    assert orc.bitshuffle(col, 2).tobytes()[:8] == bytes([0x80] * 8)


# ---- LZ4 format referee -------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 5, 12, 13, 14, 64, 1000, 4096, 65536, 65537, 200000])
def test_lz4_streams_are_valid_for_liblz4(orc, n):
    if orc.liblz4() is None:
        pytest.skip("liblz4.so.1 not present")
    for name, data in dg.corpus(n).items():
        comp = orc.lz4_compress(data)
        assert len(comp) <= orc.lz4_bound(n)
        back = orc.liblz4_decompress(comp, n)
        assert back is not None and np.array_equal(back, data), name
        assert np.array_equal(orc.lz4_decompress(comp, n), data), name
        third = orc.liblz4_compress(data)                 # third-party stream -> oracle decoder
        assert np.array_equal(orc.lz4_decompress(third, n), data), name
        assert orc.lz4_decompress(comp, n - 1) is None if n > 1 else True   # would overrun dst
        got = orc.lz4_decompress(comp, n + 7)             # roomier dst: decodes n bytes (short)
        assert got is not None and len(got) == n


def test_lz4_decoder_rejects_malformed(orc):
    bad = [b"\xff\xff\xff\xff", b"\x10", b"\x00\x00\x00", b"\x0f\x01\x00", b"\x10a\x05\x00",
           b"\x11a\x00\x00\x00", b"\x1fa\x01\x00\xff", b"\x40abcd\x00"]
    for s in bad:
        assert orc.lz4_decompress(np.frombuffer(s, dtype=np.uint8), 64) is None, s.hex()
    assert len(orc.lz4_decompress(np.frombuffer(b"", dtype=np.uint8), 10)) == 0
    assert len(orc.lz4_decompress(np.frombuffer(b"\x00", dtype=np.uint8), 10)) == 0
    ok = b"\x11a\x01\x00" + b"\x10b"      # 'a', match(off 1, len 5), 'b'
    assert orc.lz4_decompress(np.frombuffer(ok, dtype=np.uint8), 64).tobytes() == b"aaaaaab"


def test_batch_drivers_match_single_frame(orc):
    frames = [dg.smooth_f32(4096, 1), dg.random_bytes(5000, 2), dg.lowent_i16(3000, 3), dg.ramp(100003)]
    src = np.concatenate(frames)
    lens = np.array([len(f) for f in frames], dtype=np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.uint64)
    for fast in (0, 1):
        rc, dst, dst_off, dst_len = orc.compress_batch_mt(src, offs, lens, orc.SHUFFLE, 4, threads=3, fast=fast)
        assert rc == 0
        for f, data in enumerate(frames):
            fr = dst[int(dst_off[f]):int(dst_off[f]) + int(dst_len[f])]
            rc1, want = orc.compress(data, orc.LZ4, 5, orc.SHUFFLE, 4)
            assert rc1 == 0 and np.array_equal(fr, want)
        rc, out, out_len = orc.decompress_batch_mt(dst, dst_off, dst_len, offs, src.size, threads=3, fast=fast)
        assert rc == 0 and np.array_equal(out, src) and np.array_equal(out_len, lens)
    for mode in (orc.SHUFFLE, orc.BITSHUFFLE):
        for T in (2, 4, 8):
            sh = orc.shuffle_mt(mode, 0, T, frames[3], threads=4)
            want = orc.shuffle(frames[3], T) if mode == orc.SHUFFLE else orc.bitshuffle(frames[3], T)
            assert np.array_equal(sh, want)
            assert np.array_equal(orc.shuffle_mt(mode, 1, T, sh, threads=4), frames[3])


# ---- committed golden fixtures -------------------------------------------------------------------
def test_golden_fixtures(orc):
    with open(GOLDEN) as f:
        g = json.load(f)
    import make_golden
    for item in g["filters"]:
        data = make_golden.make_input(item["input"])
        fn = {"shuffle": orc.shuffle, "unshuffle": orc.unshuffle, "bitshuffle": orc.bitshuffle,
              "bitunshuffle": orc.bitunshuffle}[item["op"]]
        assert hashlib.sha256(fn(data, item["typesize"]).tobytes()).hexdigest() == item["sha256"], item
    for item in g["frames"]:
        data = make_golden.make_input(item["input"])
        rc, fr = orc.compress(data, orc.LZ4, 5, item["shuffle"], item["typesize"], item.get("policy", 0))
        assert rc == 0
        assert fr[:16].tobytes().hex() == item["header"], item
        assert hashlib.sha256(fr.tobytes()).hexdigest() == item["frame_sha256"], item
        if "frame_hex" in item:
            assert fr.tobytes().hex() == item["frame_hex"]
        rc, out = orc.decompress(np.frombuffer(bytes.fromhex(item["frame_hex"]), dtype=np.uint8)) if "frame_hex" in item else orc.decompress(fr)
        assert rc == item.get("decode_status", 0)
        if rc == 0:
            assert hashlib.sha256(out.tobytes()).hexdigest() == item["decoded_sha256"]


def test_frames_produced_by_the_cuda_encoder_are_valid_go_blosc_frames():
    """tests/golden/gpu_frames_v1.json holds frames the CUDA path (shuffle + K3 + pack) produced on a
    B200 (generator: tests/make_gpu_golden.py).  Without a GPU: the oracle -- the restated reference
    Decompress -- and liblz4 decode them to the formula inputs, header fields are the reference's."""
    import hashlib
    import json
    from make_golden import make_input
    import oracle as orc
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpu_frames_v1.json")) as f:
        gold = json.load(f)
    assert len(gold["frames"]) >= 8
    for item in gold["frames"]:
        data = make_input(item["input"])
        assert hashlib.sha256(data.tobytes()).hexdigest() == item["input_sha256"]
        fr = np.frombuffer(bytes.fromhex(item["frame_hex"]), dtype=np.uint8)
        assert fr.size == item["frame_len"]
        rc, ref = orc.compress(data, orc.LZ4, 5, item["shuffle"], item["typesize"])
        assert rc == 0 and fr[:2].tobytes() == ref[:2].tobytes() and fr[3:12].tobytes() == ref[3:12].tobytes()
        assert (fr[2] & 0x5) == (ref[2] & 0x5)
        assert int.from_bytes(fr[12:16].tobytes(), "little") == fr.size
        rc, back = orc.decompress(fr)
        assert rc == 0 and np.array_equal(back, data), item["input"]
        if not fr[2] & 0x2:
            assert fr.size <= ref.size * 1.4 + 64, (item["input"], fr.size, ref.size)   # loose: size parity on the BASELINE configs is test_gpu_parity's job
            if orc.liblz4() is not None:
                filt = {0: lambda d, t: d, 1: orc.shuffle, 2: orc.bitshuffle}[item["shuffle"]](data, item["typesize"])
                got = orc.liblz4_decompress(fr[16:], data.size)
                assert got is not None and np.array_equal(got, filt), item["input"]
