"""Where K3 / K4 spend their time on config C3: the four byte planes of the float32 field, each as its
own data set (NoShuffle, 256 KiB frames), so that a plane's cost can be read off directly."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
ctx = pkg.Context(0)
ctx.set_option(pkg.OPT_KERNEL_TIMING, 1)
s = torch.cuda.current_stream().cuda_stream
size = int(os.environ.get("PROBE_BYTES", 8 << 30))
fl = 262144
data = gen_f32(size // 4)
# per-frame byte shuffle == what K1 hands to K3: plane k of frame f is 64 KiB
planes = data.view(-1, fl // 4, 4).permute(2, 0, 1).contiguous()        # (4, frames, 65536)
del data
n = planes[0].numel()
nf = n // fl
d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
cap = n + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(nf, dtype=torch.int64, device="cuda")
d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_st = torch.empty(nf, dtype=torch.int32, device="cuda")
d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
for k in [int(x) for x in os.environ.get('PROBE_PLANES', '0,1,2,3').split(',')]:
    src = planes[k].reshape(-1)
    for it in range(int(os.environ.get('PROBE_ITERS', 2))):
        ctx.kernel_stats_reset()
        ctx.compress_batch_dev(src, d_off, d_len, nf, n, fl, 0, 1, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
        ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, n, fl, d_olen, d_st, s)
        torch.cuda.synchronize()
    st = ctx.kernel_stats()
    per = ", ".join(f"{a.replace('_kernel', '')} {v[1]:.3f}ms" for a, v in st.items() if v[0] and v[1] > 0.05)
    print(f"plane {k}: {n >> 20} MiB, ratio {int(d_tot.item()) / n:.4f}, exact={torch.equal(d_out, src)} | {per}")
