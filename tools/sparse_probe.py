import os, sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as entry
pkg = entry.load_package(); ctx = pkg.Context(0); ctx.set_option(pkg.OPT_KERNEL_TIMING, 1)
if os.environ.get("PROBE_DECODER"):
    ctx.set_option(pkg.OPT_DECODER, int(os.environ["PROBE_DECODER"]))
s = torch.cuda.current_stream().cuda_stream
size = 1 << 30; fl = int(os.environ.get("PROBE_FRAME", 4 << 20)); nf = size // fl
# sparse array: mostly zeros with a few nonzero int32 values per KiB
g = torch.Generator(device="cuda"); g.manual_seed(1)
data = torch.zeros(size // 4, dtype=torch.int32, device="cuda")
idx = torch.randint(0, size // 4, (size // 4096,), device="cuda", generator=g)
data[idx] = torch.randint(1, 1000, (idx.numel(),), device="cuda", generator=g, dtype=torch.int32)
data = data.view(torch.uint8)
d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
cap = size + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(nf, dtype=torch.int64, device="cuda"); d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_st = torch.empty(nf, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
d_out = torch.empty_like(data); d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
for it in range(3):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    ctx.compress_batch_dev(data, d_off, d_len, nf, size, fl, 1, 4, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
    b.record()
    ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, s)
    c.record(); torch.cuda.synchronize()
st = ctx.kernel_stats()
print("   " + ", ".join(f"{k.replace('_kernel', '')} {v[1] / max(v[0], 1):.3f}ms" for k, v in st.items() if v[0]))
print(f"sparse int32, {nf} frames of {fl >> 10} KiB: ratio {int(d_tot.item())/size:.4f}, compress {size/a.elapsed_time(b)/1e6:.1f} GB/s, decompress {size/b.elapsed_time(c)/1e6:.1f} GB/s, exact {torch.equal(d_out, data)}")
