"""One C3-shaped compress + decompress (device resident) for ncu / launch lists."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream
size = int(os.environ.get("PROBE_BYTES", 256 << 20))
shuffle = int(os.environ.get("PROBE_SHUFFLE", 1))
T = int(os.environ.get("PROBE_T", 4))
hl = int(os.environ.get("PROBE_HASHLOG", 0))
iters = int(os.environ.get("PROBE_ITERS", 2))
ctx.set_option(pkg.OPT_HASH_LOG, hl)
if os.environ.get("PROBE_DECODER"):
    ctx.set_option(pkg.OPT_DECODER, int(os.environ["PROBE_DECODER"]))
fl = 262144
nf = size // fl
src = gen_f32(size // 4) if T != 8 else gen_f32(size // 4, f64=True)
d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
cap = size + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(nf, dtype=torch.int64, device="cuda")
d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_st = torch.empty(nf, dtype=torch.int32, device="cuda")
d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
d_out = torch.empty_like(src)
d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
ctx.reserve(size, nf)
for it in range(iters):
    ctx.compress_batch_dev(src, d_off, d_len, nf, size, fl, shuffle, T, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
    ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, s)
torch.cuda.synchronize()
print("ratio", int(d_tot.item()) / size, "exact", torch.equal(d_out, src), "status_ok", not bool(d_st.any()))
