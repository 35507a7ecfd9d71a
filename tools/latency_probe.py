"""Where the time of ONE small Compress / Decompress call goes (the reference's own benchmark shape, C1):
wall clock per call, device time per kernel, number of launches."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
ctx = pkg.Context(0)
for n in [int(x) for x in os.environ.get("PROBE_SIZES", "4096,100000,1048576").split(",")]:
    data = (np.arange(n) % 256).astype(np.uint8)
    for _ in range(5):
        fr = ctx.compress(data, 1, 5, 1, 4); back = ctx.decompress(fr)
    assert back == data.tobytes()
    reps = 50
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps): fr = ctx.compress(data, 1, 5, 1, 4)
    t1 = time.perf_counter()
    l1 = ctx.launch_count()
    for _ in range(reps): back = ctx.decompress(fr)
    t2 = time.perf_counter()
    l2 = ctx.launch_count()
    ctx.set_option(pkg.OPT_KERNEL_TIMING, 1); ctx.kernel_stats_reset()
    for _ in range(10): fr = ctx.compress(data, 1, 5, 1, 4)
    sc = ctx.kernel_stats(); ctx.kernel_stats_reset()
    for _ in range(10): back = ctx.decompress(fr)
    sd = ctx.kernel_stats(); ctx.set_option(pkg.OPT_KERNEL_TIMING, 0)
    fmt = lambda st: ", ".join(f"{k.replace('_kernel', '')} {1e3 * v[1] / 10:.1f}us x{v[0] // 10}" for k, v in st.items() if v[0])
    print(f"n={n}: compress {1e6 * (t1 - t0) / reps:.0f} us ({(l1 - l0) / reps:.0f} launches), decompress {1e6 * (t2 - t1) / reps:.0f} us ({(l2 - l1) / reps:.0f} launches), frame {len(fr)} B")
    print(f"   compress kernels:   {fmt(sc)}  (sum {1e3 * sum(v[1] for v in sc.values()) / 10:.0f} us)")
    print(f"   decompress kernels: {fmt(sd)}  (sum {1e3 * sum(v[1] for v in sd.values()) / 10:.0f} us)")
