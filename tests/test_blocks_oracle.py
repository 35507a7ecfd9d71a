"""CPU: the oracle of the opt-in Blosc-1 multi-block frame (oracle/blosc1_blocks.c).

The reference has no such frame (it ignores Options.BlockSize, blosc.go:227-234), so there is nothing of
the reference's to pin against: PARITY UNPINNED.  These tests hold the restated layout to the published
Blosc-1 chunk format byte by byte (header, bstarts, per-stream int32 sizes, raw-stream rule, stored
frames, the nbytes + 16 bound) with the LZ4 streams referee'd by the system liblz4.
"""
import struct

import numpy as np
import pytest

import datagen as dg


def parse(fr):
    v, vlz, flags, T, n, bs, cb = struct.unpack("<BBBBIII", fr[:16].tobytes())
    return dict(version=v, versionlz=vlz, flags=flags, typesize=T, nbytes=n, blocksize=bs, cbytes=cb)


def walk_blocks(orc, fr, data, filt):
    """Re-derive every block from the wire bytes alone; returns the number of streams seen."""
    h = parse(fr)
    n, bs, T, flags = h["nbytes"], h["blocksize"], h["typesize"], h["flags"]
    assert h["version"] == 2 and h["versionlz"] == 1 and h["cbytes"] == fr.size and flags >> 5 == 1
    nblocks = -(-n // bs)
    bstarts = np.frombuffer(fr[16:16 + 4 * nblocks].tobytes(), dtype="<i4")
    assert bstarts[0] == 16 + 4 * nblocks and np.all(np.diff(bstarts) > 4)
    streams = 0
    for b in range(nblocks):
        bsize = min(bs, n - b * bs)
        partial = bsize != bs
        ns = 1 if (flags & 0x10) or partial or T > 16 or bs // T < 128 else T
        ne = bsize // ns
        pos = int(bstarts[b])
        block = np.empty(bsize, dtype=np.uint8)
        for k in range(ns):
            c = struct.unpack("<i", fr[pos:pos + 4].tobytes())[0]
            pos += 4
            assert 0 < c <= ne
            if c == ne:
                block[k * ne:(k + 1) * ne] = fr[pos:pos + c]
            else:
                out = orc.liblz4_decompress(fr[pos:pos + c], ne) if orc.liblz4() else orc.lz4_decompress(fr[pos:pos + c], ne)
                assert out is not None and out.size == ne
                block[k * ne:(k + 1) * ne] = out
            pos += c
            streams += 1
        assert pos == (int(bstarts[b + 1]) if b + 1 < nblocks else fr.size)
        assert np.array_equal(filt(block, T), data[b * bs:b * bs + bsize])
    return streams


@pytest.mark.parametrize("shuffle,T", [(0, 1), (1, 4), (1, 3), (2, 8), (1, 17)])
@pytest.mark.parametrize("split", [False, True])
def test_layout_matches_the_published_chunk_format(orc, shuffle, T, split):
    unfilt = {0: lambda b, t: b, 1: orc.unshuffle, 2: orc.bitunshuffle}[shuffle]
    for name, data in dg.corpus(200003).items():
        for bs in (0, 4096, 65536, 1 << 20):
            rc, fr = orc.blocks_compress(data, shuffle, T, bs, split)
            assert rc == 0
            h = parse(fr)
            assert h["nbytes"] == data.size and h["typesize"] == T and fr.size <= data.size + 16
            assert h["blocksize"] == orc.blocks_blocksize(data.size, T, bs)
            assert bool(h["flags"] & 0x10) == (not split)
            assert (h["flags"] & 5) == {0: 0, 1: 1, 2: 4}[shuffle]
            if h["flags"] & 2:
                assert fr.size == data.size + 16 and np.array_equal(fr[16:], data)
            else:
                ns = walk_blocks(orc, fr, data, unfilt if T > 1 else (lambda b, t: b))
                nblocks = -(-data.size // h["blocksize"])
                if not split or T > 16 or h["blocksize"] // T < 128:
                    assert ns == nblocks
                else:
                    assert ns == (nblocks - 1) * T + (T if data.size % h["blocksize"] == 0 else 1)
            rc, back = orc.blocks_decompress(fr)
            assert rc == 0 and np.array_equal(back, data), (name, bs)


@pytest.mark.parametrize("n", [1, 2, 5, 127, 128, 129, 255, 4096, 65535, 65536, 65537, 3 * 65536])
def test_sizes_and_small_buffers(orc, n):
    for data in (dg.ramp(n), dg.random_bytes(n, 1), np.zeros(n, dtype=np.uint8)):
        rc, fr = orc.blocks_compress(data, 1, 4, 0, False)
        assert rc == 0 and fr.size <= n + 16
        h = parse(fr)
        if n < 128:
            assert h["flags"] & 2                    # Blosc-1 stores buffers below 128 bytes
        rc, back = orc.blocks_decompress(fr)
        assert rc == 0 and np.array_equal(back, data)


def test_errors(orc):
    assert orc.blocks_compress(np.zeros(0, dtype=np.uint8))[0] == orc.EINVALID_DATA
    data = dg.smooth_f32(50000, 3)
    rc, fr = orc.blocks_compress(data, 1, 4, 16384, False)
    assert rc == 0 and not parse(fr)["flags"] & 2
    assert orc.blocks_decompress(fr[:10])[0] == orc.EINVALID_HEADER
    bad = fr.copy(); bad[0] = 3
    assert orc.blocks_decompress(bad)[0] == orc.EINVALID_VERSION
    assert orc.blocks_decompress(fr[:-1])[0] == orc.EINVALID_DATA            # cbytes > len
    bad = fr.copy(); bad[2] = (bad[2] & 0x1F) | (4 << 5)
    assert orc.blocks_decompress(bad)[0] == orc.EUNSUPPORTED                 # zstd format id
    bad = fr.copy(); bad[2] = (bad[2] & 0x1F) | (7 << 5)
    assert orc.blocks_decompress(bad)[0] == orc.EINVALID_CODEC
    bad = fr.copy(); bad[16:20] = np.frombuffer(struct.pack("<i", 8), dtype=np.uint8)
    assert orc.blocks_decompress(bad)[0] == orc.EDECOMPRESSION_FAILED        # bstart inside the table
    bad = fr.copy(); bad[40] ^= 0xFF
    rc, back = orc.blocks_decompress(bad)
    assert rc == orc.EDECOMPRESSION_FAILED or not np.array_equal(back, data)
    assert orc.blocks_decompress(fr, cap=100)[0] == orc.EDST_TOO_SMALL
    stored = orc.blocks_compress(dg.random_bytes(1000, 2), 1, 4, 0, False)[1]
    assert parse(stored)["flags"] & 2
    assert orc.blocks_decompress(np.concatenate([stored, stored[:4]]))[0] == 0   # trailing bytes are not the frame
    bad = stored.copy(); bad[12:16] = np.frombuffer(struct.pack("<I", 1000 + 15), dtype=np.uint8)
    assert orc.blocks_decompress(bad)[0] == orc.ESIZE_MISMATCH


def test_block_frames_produced_by_the_cuda_path_hold_to_the_format(orc):
    """tests/golden/gpu_block_frames_v1.json: multi-block frames that b2b_compress_blocks produced on a
    B200 (generator: tests/make_gpu_golden.py).  Without a GPU: every block is re-derived from the wire
    bytes (liblz4 on the streams), the oracle decodes the frames, header fields are the oracle's."""
    import hashlib
    import json
    import os
    from make_golden import make_input
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpu_block_frames_v1.json")) as f:
        gold = json.load(f)
    assert len(gold["frames"]) >= 8
    stored = 0
    for item in gold["frames"]:
        data = make_input(item["input"])
        assert hashlib.sha256(data.tobytes()).hexdigest() == item["input_sha256"]
        fr = np.frombuffer(bytes.fromhex(item["frame_hex"]), dtype=np.uint8)
        sh, T = item["shuffle"], item["typesize"]
        h = parse(fr)
        rc, ref = orc.blocks_compress(data, sh, T, item["blocksize"], False)
        hr = parse(ref)
        assert rc == 0 and fr.size == item["frame_len"] == h["cbytes"] and fr.size <= data.size + 16
        assert all(h[k] == hr[k] for k in ("version", "versionlz", "typesize", "nbytes", "blocksize"))
        assert (h["flags"] | 2) == (hr["flags"] | 2)
        rc, back = orc.blocks_decompress(fr)
        assert rc == 0 and np.array_equal(back, data), item["input"]
        if h["flags"] & 2:
            stored += 1
            assert np.array_equal(fr[16:], data)
        else:
            unfilt = {0: lambda b, t: b, 1: orc.unshuffle, 2: orc.bitunshuffle}[sh] if T > 1 else (lambda b, t: b)
            assert walk_blocks(orc, fr, data, unfilt) == -(-data.size // h["blocksize"])
            assert fr.size <= ref.size * 1.5 + 64
    assert stored >= 2          # the incompressible input and the 100-byte buffer
