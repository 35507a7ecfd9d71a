"""ONE large frame (PROBE_BYTES, default 256 MiB of the C2/C3 field, float32 + Shuffle T=4) through compress and the
pointer-jumping decoder, for ncu: the pass between cudaProfilerStart / Stop is the one to capture
(ncu --profile-from-start off ...), after a warm-up pass that grows the arena."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream
size = int(os.environ.get("PROBE_BYTES", str(256 << 20)))
src = gen_f32(size // 4)
d_off = torch.zeros(1, dtype=torch.int64, device="cuda")
d_len = torch.tensor([size], dtype=torch.int32, device="cuda")
cap = size + 96
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(1, dtype=torch.int64, device="cuda"); d_flen = torch.empty(1, dtype=torch.int32, device="cuda")
d_st = torch.empty(1, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
d_out = torch.empty_like(src); d_olen = torch.empty(1, dtype=torch.int32, device="cuda")


def one_pass():
    ctx.compress_batch_dev(src, d_off, d_len, 1, size, size, 1, 4, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
    ctx.decompress_batch_dev(d_c, d_foff, d_flen, 1, 0, d_out, d_off, d_len, size, size, d_olen, d_st, s)
    torch.cuda.synchronize()


one_pass()
torch.cuda.profiler.start()
one_pass()
torch.cuda.profiler.stop()
print("exact", torch.equal(d_out, src), "status", int(d_st.item()), "ratio", int(d_tot.item()) / size)
