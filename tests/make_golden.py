"""Generates tests/golden/golden_v1.json from the CPU oracle.

There is no runnable reference here (Go, no toolchain), so these are ORACLE-produced vectors:
they freeze the restatement (any later change to the oracle or the kernels shows up as a diff)
and travel to the GPU box, where the CUDA path must reproduce the filter hashes bit for bit
and decode the frames to the recorded data.  Inputs are formulas (see make_input), so only
hashes and a few short frames are stored.  Run:  python tests/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import datagen as dg  # noqa: E402


def make_input(spec):
    kind, n = spec["kind"], spec["n"]
    if kind == "ramp":
        return dg.ramp(n)
    if kind == "lcg":
        return dg.lcg_bytes(n, spec.get("seed", 12345))
    if kind == "f32_ramp":
        return dg.f32_ramp(n // 4, spec.get("k", 0.1))
    if kind == "f64_ramp":
        return dg.f64_ramp(n // 8, spec.get("k", 0.1))
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "period3":
        return np.tile(np.array([1, 2, 3], dtype=np.uint8), n // 3 + 1)[:n].copy()
    if kind == "i16mod8":                      # integer-only low-entropy int16 (no RNG dependence)
        i = np.arange(n // 2, dtype=np.uint64)
        v = ((i * 2654435761) >> 7) & 7
        return v.astype(np.int16).view(np.uint8)
    raise ValueError(kind)


def main():
    import oracle as orc
    sha = lambda a: hashlib.sha256(a.tobytes()).hexdigest()
    filters, frames = [], []
    inputs = [{"kind": "ramp", "n": n} for n in (13, 127, 1000, 1003, 1024, 100003)]
    inputs += [{"kind": "lcg", "n": n} for n in (35, 4096, 65536)]
    inputs += [{"kind": "f32_ramp", "n": 40000, "k": 0.001}, {"kind": "f64_ramp", "n": 8000}]
    for spec in inputs:
        data = make_input(spec)
        for T in (2, 3, 4, 8, 16, 17):
            for op, fn in (("shuffle", orc.shuffle), ("unshuffle", orc.unshuffle),
                           ("bitshuffle", orc.bitshuffle), ("bitunshuffle", orc.bitunshuffle)):
                filters.append({"input": spec, "op": op, "typesize": T, "sha256": sha(fn(data, T))})
    fspecs = [({"kind": "ramp", "n": 1000}, 1, 4), ({"kind": "ramp", "n": 100000}, 1, 4),
              ({"kind": "ramp", "n": 1000}, 0, 1), ({"kind": "f32_ramp", "n": 4000}, 1, 4),
              ({"kind": "f64_ramp", "n": 8000}, 2, 8), ({"kind": "lcg", "n": 1000}, 0, 1),
              ({"kind": "lcg", "n": 1000}, 1, 2), ({"kind": "zeros", "n": 70000}, 1, 4),
              ({"kind": "period3", "n": 5000}, 0, 1), ({"kind": "i16mod8", "n": 131072}, 1, 2),
              ({"kind": "f32_ramp", "n": 40000, "k": 0.001}, 1, 4), ({"kind": "ramp", "n": 12}, 1, 4)]
    for spec, sh, T in fspecs:
        for policy in ((0, 1) if spec["kind"] == "lcg" and sh else (0,)):
            data = make_input(spec)
            rc, fr = orc.compress(data, orc.LZ4, 5, sh, T, policy)
            assert rc == 0
            rc, out = orc.decompress(fr)
            item = {"input": spec, "shuffle": sh, "typesize": T, "policy": policy,
                    "header": fr[:16].tobytes().hex(), "frame_sha256": sha(fr), "frame_len": int(fr.size),
                    "decode_status": rc, "decoded_sha256": sha(out) if rc == 0 else None,
                    "roundtrips": bool(rc == 0 and np.array_equal(out, data))}
            if fr.size <= 1100:
                item["frame_hex"] = fr.tobytes().hex()
            frames.append(item)
    out = {"note": "oracle-produced (parity unpinned for LZ4 bytes; see oracle/blosc_oracle.h)",
           "filters": filters, "frames": frames}
    os.makedirs(os.path.join(HERE, "golden"), exist_ok=True)
    with open(os.path.join(HERE, "golden", "golden_v1.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print(f"wrote {len(filters)} filter vectors, {len(frames)} frames")


if __name__ == "__main__":
    main()
