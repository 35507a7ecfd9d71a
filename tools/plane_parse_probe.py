import os, sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32
pkg = entry.load_package(); ctx = pkg.Context(0)
ctx.set_option(pkg.OPT_KERNEL_TIMING, 1)
s = torch.cuda.current_stream().cuda_stream
full = gen_f32((256 << 20) // 4)
planes = full.view(-1, 4).t().contiguous()          # 4 planes of 64 MiB
for name, src in (("plane3 only (64 MiB)", planes[3].contiguous()), ("plane2 only (64 MiB)", planes[2].contiguous()), ("planes 2+3 (128 MiB)", planes[2:4].contiguous().view(-1))):
    size = src.numel()
    d_off = torch.zeros(1, dtype=torch.int64, device="cuda"); d_len = torch.tensor([size], dtype=torch.int32, device="cuda")
    cap = size + 96
    d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_foff = torch.empty(1, dtype=torch.int64, device="cuda"); d_flen = torch.empty(1, dtype=torch.int32, device="cuda")
    d_st = torch.empty(1, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
    d_out = torch.empty_like(src); d_olen = torch.empty(1, dtype=torch.int32, device="cuda")
    ctx.compress_batch_dev(src, d_off, d_len, 1, size, size, 0, 1, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
    ctx.set_option(108, 1)   # 8 KiB chunks
    for _ in range(2):
        ctx.kernel_stats_reset()
        ctx.decompress_batch_dev(d_c, d_foff, d_flen, 1, 0, d_out, d_off, d_len, size, size, d_olen, d_st, s)
        torch.cuda.synchronize()
    st = ctx.kernel_stats()
    per = ", ".join(f"{k.replace('_kernel', '')} {v[1]:.2f}ms" for k, v in st.items() if v[0] and v[1] > 0.02)
    print(name, "ratio", int(d_tot.item()) / size, "exact", torch.equal(d_out, src), "|", per, flush=True)
