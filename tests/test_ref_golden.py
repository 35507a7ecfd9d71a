"""Vectors produced by the REAL reference (tests/tools/make_ref_golden/main.go -> tests/golden/ref_v1.json).

There is no Go toolchain in this image, so the file does not exist yet and these tests skip; they are the loader
a maintainer with Go gets for free.  With the file present they pin what is otherwise unpinned: the
oracle's compressor must reproduce the reference's NBytesComp exactly and decode its frames, the CUDA path must
decode them bit-exactly and compress within 1 % of them with the same header fields."""
import hashlib
import json
import os

import numpy as np
import pytest

import ref_golden_inputs as gi

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_v1.json")


def load():
    if not os.path.exists(PATH):
        pytest.skip("tests/golden/ref_v1.json not generated yet (needs a Go toolchain: tests/tools/make_ref_golden/main.go)")
    with open(PATH) as f:
        return json.load(f)["frames"]


def test_generators_are_deterministic():
    """The formula inputs are what the Go generator computes (integer arithmetic, exact float conversion): pin
    their hashes here so that a change of numpy or of the formulas cannot silently detach them from the Go side."""
    got = hashlib.sha256(gi.make("ramp(100000)").tobytes()).hexdigest()
    assert got == "db8f1d69251d95e2c88268d3c540533cc5182e0e33065a6f3f322f606a574489"
    for call in ("smooth_f32(65536, 3)", "smooth_f64(32768, 4)", "lowent_i16(131072, 5)", "random_bytes(65536, 6)"):
        a, b = gi.make(call), gi.make(call)
        assert np.array_equal(a, b) and a.dtype == np.uint8
    f = gi.make("smooth_f32(65536, 3)").view(np.float32)
    assert np.all(f * 1024 == np.round(f * 1024))          # exact integers / 1024
    assert gi.make("lowent_i16(131072, 5)").view(np.uint16).max() == 7


def test_oracle_reproduces_the_reference(orc):
    for e in load():
        data = gi.make(e["input"])
        assert hashlib.sha256(data.tobytes()).hexdigest() == e["input_sha256"], e["name"]
        fr = np.frombuffer(bytes.fromhex(e["frame_hex"]), dtype=np.uint8)
        rc, back = orc.decompress(fr)
        assert rc == 0
        if e["reference_round_trip"]:
            assert np.array_equal(back, data), e["name"]
        rc, mine = orc.compress(data, orc.LZ4, 5, e["shuffle"], e["typesize"], 1)     # policy 1 = the reference's memcpy quirk
        assert mine[:12].tobytes() == fr[:12].tobytes(), e["name"]
        assert mine.size == fr.size, (e["name"], mine.size, fr.size, "oracle compressor differs from pierrec/lz4")
        assert mine.tobytes() == fr.tobytes(), e["name"]


@pytest.mark.gpu
def test_cuda_path_against_the_reference(ctx):
    for e in load():
        data = gi.make(e["input"])
        fr = bytes.fromhex(e["frame_hex"])
        if e["reference_round_trip"]:
            assert ctx.decompress(fr) == data.tobytes(), e["name"]
        mine = ctx.compress(data, 1, 5, e["shuffle"], e["typesize"])
        assert mine[:12] == fr[:12], e["name"]
        if not (fr[2] & 2):
            assert len(mine) <= len(fr) * 1.01 + 16, (e["name"], len(mine), len(fr))
