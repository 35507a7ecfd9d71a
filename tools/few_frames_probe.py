"""Where the pointer-jumping decoder (B2B_OPT_DECODER 4) beats the tile engine (0): nf frames of the same size in one
device-pointer call, C3-like data (float32 + Shuffle T=4).  Prints ms per call for both."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream
for fl in (1 << 20, 4 << 20, 16 << 20):
    for nf in [int(x) for x in os.environ.get("PROBE_NF", "8,24,48,64").split(",")]:
        size = nf * fl
        if size > (1 << 30):
            continue
        src = gen_f32(size // 4)
        d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
        d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
        cap = size + 32 * nf + 64
        d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
        d_foff = torch.empty(nf, dtype=torch.int64, device="cuda"); d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
        d_st = torch.empty(nf, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
        d_out = torch.empty_like(src); d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
        ctx.compress_batch_dev(src, d_off, d_len, nf, size, fl, 1, 4, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
        res = []
        for variant in (0, 4):
            ctx.set_option(pkg.OPT_DECODER, variant)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, s)
            d_out.zero_()
            a.record()
            ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, s)
            b.record(); torch.cuda.synchronize()
            res.append((a.elapsed_time(b), bool(torch.equal(d_out, src))))
        ctx.set_option(pkg.OPT_DECODER, -1)
        print(f"{nf:3d} frames x {fl >> 20:2d} MiB: tile engine {res[0][0]:7.2f} ms ({size / res[0][0] / 1e6:6.1f} GB/s) | pointer jumping {res[1][0]:7.2f} ms ({size / res[1][0] / 1e6:6.1f} GB/s) | exact {res[0][1]} {res[1][1]}", flush=True)
        del src, d_c, d_out
