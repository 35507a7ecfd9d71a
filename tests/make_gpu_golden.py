"""Generates tests/golden/gpu_frames_v1.json ON A B200: frames produced by the CUDA encoder (K1/K2 +
K3 + pack through the C ABI) for formula inputs (tests/make_golden.make_input), stored as hex.

The CPU test suite then shows, without a GPU, that frames the CUDA path really produced are valid
go-blosc frames: the oracle (reference Decompress semantics) and liblz4 decode them to the inputs.
Re-run after an encoder change:  gpurun -- 'python tests/make_gpu_golden.py'  and copy the file
back from gpurun_out/.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from make_golden import make_input  # noqa: E402
import __graft_entry__ as entry  # noqa: E402


def main():
    pkg = entry.load_package()
    ctx = pkg.Context(0)
    specs = [({"kind": "ramp", "n": 100000}, 1, 4), ({"kind": "ramp", "n": 1003}, 1, 4),
             ({"kind": "f32_ramp", "n": 40000, "k": 0.001}, 1, 4), ({"kind": "f64_ramp", "n": 8000}, 2, 8),
             ({"kind": "i16mod8", "n": 70000}, 1, 2), ({"kind": "period3", "n": 4096}, 0, 1),
             ({"kind": "lcg", "n": 4096}, 1, 4), ({"kind": "zeros", "n": 200000}, 0, 1),
             ({"kind": "i16mod8", "n": 3000}, 2, 2), ({"kind": "f32_ramp", "n": 140000, "k": 0.1}, 1, 4)]
    frames = []
    for spec, sh, T in specs:
        data = make_input(spec)
        fr = ctx.compress(data, 1, 5, sh, T)
        assert ctx.decompress(fr) == data.tobytes()
        frames.append({"input": spec, "shuffle": sh, "typesize": T, "frame_len": len(fr),
                       "input_sha256": hashlib.sha256(data.tobytes()).hexdigest(), "frame_hex": bytes(fr).hex()})
    out = {"note": "frames produced by the CUDA encoder of this repository on a B200 (see tests/make_gpu_golden.py)",
           "library": pkg.lib().b2b_version().decode() if hasattr(pkg.lib(), "b2b_version") else "", "frames": frames}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_frames_v1.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(frames), "frames,", sum(fr["frame_len"] for fr in frames), "bytes")

    # the opt-in Blosc-1 multi-block frames (b2b_compress_blocks), same idea: gpu_block_frames_v1.json
    bspecs = [({"kind": "ramp", "n": 100000}, 1, 4, 16384), ({"kind": "f32_ramp", "n": 40000, "k": 0.001}, 1, 4, 0),
              ({"kind": "f64_ramp", "n": 8000}, 2, 8, 4096), ({"kind": "i16mod8", "n": 70000}, 1, 2, 32768),
              ({"kind": "lcg", "n": 4096}, 1, 4, 1024), ({"kind": "zeros", "n": 200000}, 0, 1, 0),
              ({"kind": "period3", "n": 4096}, 0, 1, 1000), ({"kind": "ramp", "n": 100}, 1, 4, 0),
              ({"kind": "f32_ramp", "n": 14000, "k": 0.001}, 1, 3, 5000)]
    bframes = []
    for spec, sh, T, bs in bspecs:
        data = make_input(spec)
        fr = ctx.compress_blocks(data, sh, T, bs)
        assert ctx.decompress_blocks(fr) == data.tobytes()
        bframes.append({"input": spec, "shuffle": sh, "typesize": T, "blocksize": bs, "frame_len": len(fr),
                        "input_sha256": hashlib.sha256(data.tobytes()).hexdigest(), "frame_hex": bytes(fr).hex()})
    out = {"note": "Blosc-1 multi-block frames produced by the CUDA path of this repository on a B200 "
                   "(b2b_compress_blocks; see tests/make_gpu_golden.py)", "frames": bframes}
    with open(os.path.join(ROOT, "gpurun_out", "gpu_block_frames_v1.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(bframes), "block frames,", sum(fr["frame_len"] for fr in bframes), "bytes")


if __name__ == "__main__":
    main()
