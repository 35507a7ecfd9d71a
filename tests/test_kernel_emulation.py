"""Kernel LOGIC on the CPU: the device code of go-blosc_b200/csrc/*.cuh compiled by g++ against the
coroutine shim in tests/emu (one coroutine per CUDA thread, warp collectives as rendezvous) and run
through the same launch sequences as csrc/b2b.cu, one frame at a time.

This is test infrastructure, not a product path (the library has no CPU path and fails without a
device).  It lets the CPU suite -- which has no GPU -- hold the K3 match finder / emitter, finalize,
pack, the K4 token parser and copy stage, the filters and the scan to the oracle, and it is what the
encoder's candidate policies were tuned on (tests/tools/emu_sizes.py).  Lanes of a warp run one
after the other between two collectives instead of in lockstep, so compressed SIZES can differ from
the GPU's by a few hundredths of a percent (hash-table races resolve differently); validity,
decoded bytes, header fields and status words do not.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import datagen as dg
import reference_suite as rs

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "_build", "libemu_codec.so")


def build_emu(force=False):
    src = [os.path.join(HERE, "emu", f) for f in ("emu_codec.cpp", "emu_runtime.cpp", "cuda_runtime.h")]
    csrc = os.path.join(ROOT, "go-blosc_b200", "csrc")
    deps = src + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(HERE, "emu"), "-shared", "-fPIC",
                               "-o", LIB, src[0], src[1]])
    lib = ctypes.CDLL(LIB)
    lib.emu_compress_frame.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_uint64, ctypes.c_void_p]
    lib.emu_decompress_frame.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_uint32, ctypes.c_void_p]
    lib.emu_filter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32, ctypes.c_int]
    return lib


class Emu:
    def __init__(self):
        self.lib = build_emu()

    def compress(self, data, shuffle, typesize, hash_log=0, hash_bytes=0, quirk=0, independent=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(data.size + 64, dtype=np.uint8)
        n = ctypes.c_uint32(0)
        st = self.lib.emu_compress_frame(data.ctypes.data, data.size, shuffle, typesize, hash_log, hash_bytes, None,
                                         independent, quirk, out.ctypes.data, out.size, ctypes.byref(n))
        return st, out[:n.value].copy()

    def decompress(self, frame, cap, split=1, override=0):
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = ctypes.c_uint32(0)
        st = self.lib.emu_decompress_frame(frame.ctypes.data, frame.size, override, split, out.ctypes.data, cap, ctypes.byref(n))
        return st, out[:n.value].copy()

    def filter(self, data, mode, typesize, inverse=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty_like(data)
        self.lib.emu_filter(data.ctypes.data, out.ctypes.data, data.size, mode, typesize, inverse)
        return out


@pytest.fixture(scope="module")
def emu():
    return Emu()


@pytest.mark.parametrize("T", [2, 3, 4, 7, 8, 16])
def test_emulated_filters_match_the_oracle(emu, orc, T):
    for n in (1, 15, 64, 1003, 16384, 40000):
        data = dg.random_bytes(n, n + T)
        assert np.array_equal(emu.filter(data, 1, T), orc.shuffle(data, T)), (n, T)
        assert np.array_equal(emu.filter(data, 2, T), orc.bitshuffle(data, T)), (n, T)
        assert np.array_equal(emu.filter(orc.shuffle(data, T), 1, T, 1), data), (n, T)
        assert np.array_equal(emu.filter(orc.bitshuffle(data, T), 2, T, 1), data), (n, T)


CASES = [
    ("C3 smooth f32", lambda: dg.smooth_f32(65536, 1), 1, 4),
    ("C4 smooth f64", lambda: dg.smooth_f64(32768, 2), 2, 8),
    ("C5 lowent int16", lambda: dg.lowent_i16(131072, 3), 1, 2),
    ("C1 ramp", lambda: dg.ramp(100000), 1, 4),
    ("text", lambda: dg.text_like(150000, 9), 0, 1),
    ("random", lambda: dg.random_bytes(70000, 5), 1, 4),
    ("zeros", lambda: np.zeros(200001, dtype=np.uint8), 0, 1),
    ("f32 bitshuffle", lambda: dg.smooth_f32(40000, 4), 2, 4),
    ("tiny", lambda: dg.ramp(13), 1, 4),
]


@pytest.mark.parametrize("name,make,sh,T", CASES, ids=[c[0] for c in CASES])
def test_emulated_encoder_frames_decode_through_the_oracle(emu, orc, name, make, sh, T):
    """K1/K2 + K3 + finalize + K5 + pack on the CPU shim: the frame is a valid go-blosc frame (the oracle =
    reference Decompress semantics decodes it), the header fields are the oracle's, and the emulated K4
    (both decoders) returns the input."""
    data = make()
    st, fr = emu.compress(data, sh, T)
    assert st == 0
    rc, back = orc.decompress(fr)
    assert rc == 0 and np.array_equal(back, data)
    rc, ref = orc.compress(data, orc.LZ4, 5, sh, T)
    assert fr[:12].tobytes() == ref[:12].tobytes()
    assert bool(fr[2] & 2) == bool(ref[2] & 2), "memcpy decision differs from the oracle's"
    for split in (0, 1, 2, 3, 4, 5):
        taken = emu.lib.emu_jump_taken()
        st, out = emu.decompress(fr, data.size, split)
        assert st == 0 and np.array_equal(out, data)
        st, out = emu.decompress(ref, data.size, split)
        assert st == 0 and np.array_equal(out, data)
        if split in (4, 5) and not (fr[2] & 2):      # the pointer-jumping engine decoded both itself (no hand-over to the tile engine)
            assert emu.lib.emu_jump_taken() == taken + 2


def test_emulated_encoder_size_against_the_oracle(emu, orc):
    """Size of the strip-parallel encoder against the restated pierrec compressor on 256 KiB frames (parity
    unpinned: see oracle/blosc_oracle.h), at the automatic table / hash policy of launch_encode."""
    n = 262144
    cases = {
        "C3": (dg.smooth_f32(n // 4, 1), 1, 4, 1.01),
        "C4": (dg.smooth_f64(n // 8, 2), 2, 8, 1.025),
        "C5": (dg.lowent_i16(n // 2, 3), 1, 2, 1.01),
        "C1": (dg.ramp(100000), 1, 4, 1.01),
        "text": (dg.text_like(n, 9), 0, 1, 1.10),
        "lowent int16 unshuffled": (dg.lowent_i16(n // 2, 3), 0, 1, 1.06),
    }
    for name, (data, sh, T, bound) in cases.items():
        st, fr = emu.compress(data, sh, T)
        rc, ref = orc.compress(data, orc.LZ4, 5, sh, T)
        assert st == 0 and fr.size <= ref.size * bound + 16, (name, fr.size, ref.size)


def test_emulated_decoder_status_words_on_mutants(emu, orc):
    """The reference's fuzz contract on the CPU shim: mutated frames get the oracle's status and bytes from both
    K4 variants (fused and parse + copy)."""
    rng = np.random.default_rng(11)
    data = dg.smooth_f32(5000, 3)
    rc, fr = orc.compress(data, orc.LZ4, 5, 1, 4)
    mutants = [fr[:k].copy() for k in (0, 5, 15, 16, 17, fr.size // 2, fr.size - 1)]
    for _ in range(40):
        m = fr.copy()
        for _ in range(int(rng.integers(1, 4))):
            m[int(rng.integers(0, m.size))] = int(rng.integers(0, 256))
        mutants.append(m)
    m = fr.copy(); m[16:] ^= 0xFF; mutants.append(m)
    for m in mutants:
        rc, want = orc.decompress(m)
        norig = int.from_bytes(m[4:8].tobytes(), "little") if m.size >= 16 else 0
        cap = min(norig, 255 * m.size + 64)
        for split in (0, 1, 2, 3, 4, 5):
            st, out = emu.decompress(m, cap, split)
            assert st == rc, (st, rc, m[:16].tobytes().hex())
            if rc == 0:
                assert np.array_equal(out, want)


def test_emulated_chunk_repair_keeps_the_speculative_records(emu, orc):
    """300 000-byte frames of the oracle's compressor over the adversarial inputs (copies of far-back chunks, runs, strip
    periods): 30-odd parse chunks each, a third of them entered off the speculative chain.  The repair kernel repairs such
    chunks on an ASSUMED entry; when the assumption is wrong the stitch kernel adopts or searches the chunk's speculative
    records, which the repair must therefore have left alone (a first version wrote its tokens over them: status 0 and
    3 105 wrong bytes on `copies`).  Both users of the record table: tile copy engine (2) and pointer jumping (4)."""
    n = 300000
    adv = dg.strip_adversarial(n, seed=n)
    for name in ("copies", "runs", "period61", "period4096"):
        data = adv[name]
        for sh, T in ((0, 1), (1, 4)):
            rc, fr = orc.compress(data, orc.LZ4, 5, sh, T)
            for split in (2, 4, 5):
                st, out = emu.decompress(np.asarray(fr, dtype=np.uint8), n, split)
                assert st == 0 and np.array_equal(out, data), (name, sh, T, split)
