"""One LARGE frame (the reference's format allows up to 4 GiB per frame, blosc.go:363-365) through the device path:
compress once, then time Decompress with each K4 variant.  PROBE_BYTES = frame size (default 256 MiB)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
ctx = pkg.Context(0)
ctx.set_option(pkg.OPT_KERNEL_TIMING, 1)
if os.environ.get("PROBE_PARSE_CTAS"):
    ctx.set_option(109, int(os.environ["PROBE_PARSE_CTAS"]))
s = torch.cuda.current_stream().cuda_stream
for size in [int(x) for x in os.environ.get("PROBE_BYTES", str(256 << 20)).split(",")]:
    src = gen_f32(size // 4)
    d_off = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_len = torch.full((1,), size, dtype=torch.int32 if size < 2**31 else torch.int64, device="cuda").to(torch.int32) if size < 2**31 else None
    if d_len is None:
        d_len = torch.tensor([size - 2**32], dtype=torch.int32, device="cuda")        # u32 bit pattern
    cap = size + 96
    d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_foff = torch.empty(1, dtype=torch.int64, device="cuda"); d_flen = torch.empty(1, dtype=torch.int32, device="cuda")
    d_st = torch.empty(1, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
    d_out = torch.empty_like(src); d_olen = torch.empty(1, dtype=torch.int32, device="cuda")
    for sh, T, name in ((1, 4, "Shuffle1 T=4"), (0, 1, "NoShuffle")):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.compress_batch_dev(src, d_off, d_len, 1, size, min(size, 2**32 - 1), sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
        ctx.kernel_stats_reset()
        a.record()
        ctx.compress_batch_dev(src, d_off, d_len, 1, size, min(size, 2**32 - 1), sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
        b.record(); torch.cuda.synchronize()
        per = ", ".join(f"{k.replace('_kernel', '')} {v[1]:.2f}ms" for k, v in ctx.kernel_stats().items() if v[0] and v[1] > 0.05)
        print(f"{size >> 20} MiB frame, {name}: compress {a.elapsed_time(b):.2f} ms = {size / a.elapsed_time(b) / 1e6:.1f} GB/s, ratio {int(d_tot.item()) / size:.4f}, status {int(d_st.item())} | {per}", flush=True)
        for variant in [int(v) for v in os.environ.get('PROBE_VARIANTS', '4,-1,0').split(',')]:
            ctx.set_option(pkg.OPT_DECODER, variant)
            ctx.decompress_batch_dev(d_c, d_foff, d_flen, 1, 0, d_out, d_off, d_len, size, min(size, 2**32 - 1), d_olen, d_st, s)   # (grows the arena)
            d_out.zero_()
            torch.cuda.synchronize()
            ctx.kernel_stats_reset()
            a.record()
            ctx.decompress_batch_dev(d_c, d_foff, d_flen, 1, 0, d_out, d_off, d_len, size, min(size, 2**32 - 1), d_olen, d_st, s)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            st = ctx.kernel_stats()
            per = ", ".join(f"{k.replace('_kernel', '')} {v[1]:.2f}ms" for k, v in st.items() if v[0] and v[1] > 0.05)
            print(f"   decoder {variant}: {ms:.1f} ms = {size / ms / 1e6:.2f} GB/s, exact={torch.equal(d_out, src)}, status {int(d_st.item())} | {per}", flush=True)
        ctx.set_option(pkg.OPT_DECODER, -1)
    del src, d_c, d_out
