"""Every K1 / K2 shape once on one buffer (for ncu --set full), then a timing table that includes the shapes outside
the 16-byte fast paths (odd typesizes, sizes whose element count is not a multiple of 16, unaligned bases)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream
size = int(os.environ.get("PROBE_BYTES", 1 << 30))
src = gen_f32(size // 4 + 64)
dst = torch.empty_like(src)
shapes = [(mode, T, inv) for mode in (1, 2) for T in (2, 4, 8, 16) for inv in (0, 1)]
for mode, T, inv in shapes:                      # launch order = the order of the ncu report
    ctx.shuffle_dev(mode, bool(inv), T, src, dst, size, s)
torch.cuda.synchronize()
if os.environ.get("PROBE_TIMING", "1") != "0":
    def timeit(fn, iters=5):
        fn(); fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    cp = timeit(lambda: dst[:size].copy_(src[:size]))
    print(f"torch copy {2 * size / cp / 1e6:.0f} GB/s of traffic")
    rows = []
    for mode in (1, 2):
        for T in (2, 3, 4, 5, 6, 7, 8, 12, 16, 32):
            for label, n, so, do in (("aligned", size, 0, 0), ("E%16!=0", size - 16 * T + T * 5, 0, 0), ("base+4", size - 64, 4, 4)):
                for inv in (0, 1):
                    a, b = src[so:so + n], dst[do:do + n]
                    t = timeit(lambda: ctx.shuffle_dev(mode, bool(inv), T, a, b, n, s), iters=3)
                    rows.append((mode, T, label, inv, 2 * n / t / 1e6))
    for mode, T, label, inv, gbs in rows:
        print(f"{'shuffle' if mode == 1 else 'bitshuffle':10s} T={T:2d} {label:8s} inv={inv}: {gbs:7.0f} GB/s of traffic")
