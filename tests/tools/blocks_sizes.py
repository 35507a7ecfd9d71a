"""Sizes of GPU block frames vs the oracle's block frames and vs the GPU's one-block frames (test corpus)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry
import datagen as dg

pkg = entry.load_package(); orc = entry.load_oracle(); ctx = pkg.Context(0)
for shuffle, T in ((0, 1), (1, 4), (2, 8), (1, 2)):
    for name, data in dg.corpus(1 << 20).items():
        row = []
        for bs in (4096, 32768, 0, 262144):
            g = len(ctx.compress_blocks(data, shuffle, T, bs))
            o = orc.blocks_compress(data, shuffle, T, bs, False)[1].size
            row.append(f"bs={bs or 65536}: {g}/{o}={g / o:.3f}")
        one = len(ctx.compress(data, pkg.Codec.LZ4, 5, shuffle, T))
        print(f"sh={shuffle} T={T} {name:14s} one-block {one:8d} | " + " | ".join(row))
