"""Compress a few frames of a probe workload on the GPU and dump them (for offline stream statistics)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32
pkg = entry.load_package()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream
fl = 262144; nf = 64; size = nf * fl
for name, data, sh, T in (("c3", gen_f32(size // 4), 1, 4), ("c4", gen_f32(size // 4, f64=True), 2, 8)):
    d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
    d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
    cap = size + 32 * nf + 64
    d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_foff = torch.empty(nf, dtype=torch.int64, device="cuda"); d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
    d_st = torch.empty(nf, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
    ctx.compress_batch_dev(data, d_off, d_len, nf, size, fl, sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
    torch.cuda.synchronize()
    foff, flen = d_foff.cpu().numpy(), d_flen.cpu().numpy()
    for f in (3, 40):
        d_c[int(foff[f]):int(foff[f]) + int(flen[f])].cpu().numpy().tofile(f"gpurun_out/frame_{name}_{f}.bin")
        data[f * fl:(f + 1) * fl].cpu().numpy().tofile(f"gpurun_out/raw_{name}_{f}.bin")
print("ok")
