import os, sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as entry
pkg = entry.load_package(); ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream
size = 512 << 20; fl = 2 << 20; nf = size // fl
g = torch.Generator(device="cuda"); g.manual_seed(7)
data = torch.randint(0, 8, (size // 2,), device="cuda", generator=g, dtype=torch.int16).view(torch.uint8)
d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
cap = size + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(nf, dtype=torch.int64, device="cuda"); d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_st = torch.empty(nf, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
d_out = torch.empty_like(data); d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.compress_batch_dev(data, d_off, d_len, nf, size, fl, 1, 2, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
    ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, s)
torch.cuda.synchronize()
print("exact", torch.equal(d_out, data))
