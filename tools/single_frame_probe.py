"""ONE frame through b2b_decompress / b2b_compress (host pointers: what decompressBackend binds), sizes 16 KiB .. 64 MiB,
against the oracle on one CPU core; the decoder choice for a handful of frames is swept with option 107."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
ctx = pkg.Context(0)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import datagen as dg  # noqa: E402

for n in [int(x) for x in os.environ.get("PROBE_SIZES", "16384,65536,262144,1048576,4194304,16777216,67108864").split(",")]:
    data = dg.smooth_f32(n // 4, 7)
    fr = ctx.compress(data, 1, 5, 1, 4)
    line = f"n={n:9d} frame {len(fr):9d}:"
    for thr in [int(x) for x in os.environ.get("PROBE_THRESHOLDS", "0,8192").split(",")]:
        ctx.set_option(107, thr)
        for _ in range(3):
            back = ctx.decompress(fr)
        assert back == data.tobytes()
        reps = 20 if n <= (4 << 20) else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            back = ctx.decompress(fr)
        dt = (time.perf_counter() - t0) / reps
        line += f"  decompress[min={thr}] {1e6 * dt:9.0f} us ({n / dt / 1e9:6.2f} GB/s)"
    ctx.set_option(107, 0)
    ctx.set_option(pkg.OPT_KERNEL_TIMING, 1); ctx.kernel_stats_reset()
    for _ in range(5):
        back = ctx.decompress(fr)
    st = ctx.kernel_stats(); ctx.set_option(pkg.OPT_KERNEL_TIMING, 0)
    line += "\n      decompress kernels (us): " + ", ".join(f"{k.replace('_kernel', '').replace('lz4_', '')} {1e3 * v[1] / 5:.0f}" for k, v in st.items() if v[0])
    t0 = time.perf_counter()
    for _ in range(5):
        fr = ctx.compress(data, 1, 5, 1, 4)
    dt = (time.perf_counter() - t0) / 5
    line += f"  compress {1e6 * dt:9.0f} us ({n / dt / 1e9:6.2f} GB/s)"
    print(line, flush=True)
