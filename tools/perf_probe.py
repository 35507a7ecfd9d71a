"""Quick device-resident timing probe (not the bench): filters and the C3-shaped batch."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[0], ts[len(ts) // 2]


def gen_f32(n_elems, device="cuda"):
    i = torch.arange(n_elems, device=device, dtype=torch.float64)
    g = torch.Generator(device=device); g.manual_seed(0xB200)
    u = torch.rand(n_elems, device=device, generator=g, dtype=torch.float64) * 2 - 1
    x = torch.sin(2 * torch.pi * i / 4096) + 0.25 * torch.sin(2 * torch.pi * i / 333.3) + 1e-3 * u
    return x.to(torch.float32).view(torch.uint8)


size = int(os.environ.get("PROBE_BYTES", 4 << 30))
src = gen_f32(size // 4)
dst = torch.empty_like(src)
print(f"device {torch.cuda.get_device_name(0)}, buffer {size >> 20} MiB")
cp = timeit(lambda: dst.copy_(src))
print(f"torch copy: {2 * size / cp[0] / 1e6:.0f} GB/s (read+write)")
for cps in (0,):
    ctx.set_option(pkg.OPT_FILTER_CTAS_PER_SM, cps)
    for mode, name in ((1, "shuffle"), (2, "bitshuffle")):
        for T in (2, 4, 8, 16):
            for inv in (False, True):
                best, med = timeit(lambda: ctx.shuffle_dev(mode, inv, T, src, dst, size, s))
                print(f"ctas/sm={cps:2d} {name:10s} T={T:2d} inv={int(inv)}: best {size / best / 1e6:7.0f} GB/s uncompressed "
                      f"({2 * size / best / 1e6:7.0f} GB/s traffic), median {size / med / 1e6:7.0f}")
ctx.set_option(pkg.OPT_FILTER_CTAS_PER_SM, 0)

# C3-shaped batch
fl = 262144
nf = size // fl
d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
cap = size + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(nf, dtype=torch.int64, device="cuda")
d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_st = torch.empty(nf, dtype=torch.int32, device="cuda")
d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
ctx.reserve(size, nf)
d_out = torch.empty_like(src)
d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
from tools.perf_probe_lib import gen_f32 as gen2
src64 = gen2(size // 4, f64=True)
for label, data, sh, T in (("C3 f32 Shuffle T=4", src, 1, 4), ("C4 f64 BitShuffle T=8", src64, 2, 8)):
    for hl in (10, 11):
        ctx.set_option(pkg.OPT_HASH_LOG, hl)
        comp = lambda: ctx.compress_batch_dev(data, d_off, d_len, nf, size, fl, sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
        best, med = timeit(comp, iters=3, warm=1)
        total = int(d_tot.item())
        print(f"compress {label} hashlog={hl}: {nf} frames, ratio {total / size:.4f}, best {size / best / 1e6:.1f} GB/s, median {size / med / 1e6:.1f} GB/s, status ok={not bool(d_st.any())}")
        dec = lambda: ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, s)
        best, med = timeit(dec, iters=3, warm=1)
        print(f"decompress {label}: best {size / best / 1e6:.1f} GB/s, median {size / med / 1e6:.1f} GB/s, exact={torch.equal(d_out, data)}, status ok={not bool(d_st.any())}")
print("launches", ctx.launch_count())
