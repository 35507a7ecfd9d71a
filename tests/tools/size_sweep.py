"""Compressed size of the CUDA encoder vs the oracle over a corpus x size x filter grid: worst offenders first."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import __graft_entry__ as entry
import datagen as dg
pkg = entry.load_package(); orc = entry.load_oracle()
ctx = pkg.Context(0)
rows = []
for n in (1000, 4096, 20000, 70000, 200000, 600000):
    inputs = dict(dg.corpus(n)); inputs.update({"adv_" + k: v for k, v in dg.strip_adversarial(n, 3).items()})
    for name, data in inputs.items():
        for sh, T in ((0, 1), (1, 4), (2, 8), (1, 2)):
            g = len(ctx.compress(data, 1, 5, sh, T))
            rc, ref = orc.compress(data, orc.LZ4, 5, sh, T)
            rows.append((g / ref.size, g - ref.size, name, n, sh, T, g, ref.size))
rows.sort(reverse=True)
print("worst by ratio:")
for r in rows[:25]:
    print(f"  x{r[0]:.3f} (+{r[1]} B)  {r[2]:18s} n={r[3]:6d} sh={r[4]} T={r[5]}  gpu {r[6]} oracle {r[7]}")
print("worst by absolute excess / n:")
for r in sorted(rows, key=lambda r: -r[1] / r[3])[:15]:
    print(f"  +{100 * r[1] / r[3]:.2f}% of n  x{r[0]:.3f}  {r[2]:18s} n={r[3]:6d} sh={r[4]} T={r[5]}  gpu {r[6]} oracle {r[7]}")
tot_g = sum(r[6] for r in rows); tot_o = sum(r[7] for r in rows)
print(f"{len(rows)} cases, total gpu {tot_g} oracle {tot_o} ratio {tot_g / tot_o:.4f}; better-than-oracle cases: {sum(1 for r in rows if r[0] < 1)}")
