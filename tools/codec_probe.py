"""Device-resident timing probe of the codec kernels over several data shapes (not the bench).
Per workload: compress / decompress GB/s (uncompressed bytes), per-kernel device ms, exactness."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
ctx = pkg.Context(0)
ctx.set_option(pkg.OPT_KERNEL_TIMING, 1)
if os.environ.get("PROBE_DECODER"):
    ctx.set_option(pkg.OPT_DECODER, int(os.environ["PROBE_DECODER"]))      # -1 automatic, 0 chunk-parallel, 1 fused, 2 parse + copy
if os.environ.get("PROBE_ONESTREAM"):
    ctx.set_option(pkg.OPT_DECODE_STREAMS, int(os.environ["PROBE_ONESTREAM"]))
if os.environ.get("PROBE_PERSIST"):
    ctx.set_option(105, int(os.environ["PROBE_PERSIST"]))
if os.environ.get("PROBE_FUSE"):
    ctx.set_option(pkg.OPT_FUSE_UNSHUFFLE, int(os.environ["PROBE_FUSE"]))
s = torch.cuda.current_stream().cuda_stream
size = int(os.environ.get("PROBE_BYTES", 2 << 30))
which = os.environ.get("PROBE_CASES", "c3,c4,c5,text").split(",")
iters = int(os.environ.get("PROBE_ITERS", 3))


def lowent_i16(n_bytes):
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    return torch.randint(0, 8, (n_bytes // 2,), device="cuda", generator=g, dtype=torch.int16).view(torch.uint8)


def text_like(n_bytes):
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    # words of 2..8 letters from a 200-word dictionary, space separated (approximately)
    words = torch.randint(97, 123, (200, 9), device="cuda", generator=g, dtype=torch.uint8)
    wl = torch.randint(2, 9, (200,), device="cuda", generator=g)
    words[torch.arange(9, device="cuda")[None, :] >= wl[:, None]] = 32
    pick = torch.randint(0, 200, (n_bytes // 6 + 8,), device="cuda", generator=g)
    rows = words[pick]                                   # (k, 9) padded with spaces
    keep = torch.arange(9, device="cuda")[None, :] <= wl[pick][:, None]
    out = rows[keep]
    reps = (n_bytes + out.numel() - 1) // out.numel()
    return out.repeat(reps)[:n_bytes].contiguous() if reps > 1 else out[:n_bytes].contiguous()


def mixed_c5(n_bytes):
    """alternating random / low-entropy int16 frames of 256 KiB (frame sizes fixed here)."""
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    a = lowent_i16(n_bytes).view(-1, 262144)
    r = torch.randint(0, 256, a.shape, device="cuda", generator=g, dtype=torch.uint8)
    a[1::2] = r[1::2]
    return a.reshape(-1)


cases = {
    "c3": ("C3 f32 Shuffle T=4", lambda: gen_f32(size // 4), 1, 4),
    "c4": ("C4 f64 BitShuffle T=8", lambda: gen_f32(size // 4, f64=True), 2, 8),
    "c5": ("C5 lowent i16 Shuffle T=2", lambda: lowent_i16(size), 1, 2),
    "c5mix": ("C5 random/lowent i16 Shuffle T=2", lambda: mixed_c5(size), 1, 2),
    "text": ("text NoShuffle", lambda: text_like(size), 0, 1),
}

fl = int(os.environ.get("PROBE_FRAME", 262144))
nf = size // fl
d_off = torch.arange(nf, dtype=torch.int64, device="cuda") * fl
d_len = torch.full((nf,), fl, dtype=torch.int32, device="cuda")
cap = size + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(nf, dtype=torch.int64, device="cuda")
d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_st = torch.empty(nf, dtype=torch.int32, device="cuda")
d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
ctx.reserve(size, nf)
print(f"device {torch.cuda.get_device_name(0)}, {size >> 20} MiB per case, {nf} frames of {fl} B")
hashlogs = [int(x) for x in os.environ.get("PROBE_HASHLOGS", "0").split(",")]
tunes = [tuple(int(v) for v in t.split(":")) for t in os.environ.get("PROBE_TUNES", "0:0:0:0").split(",")]
for key, hl, tune in [(k, h, t) for k in which for h in hashlogs for t in tunes]:
    ctx.set_option(pkg.OPT_HASH_LOG, hl)
    for i, v in enumerate(tune):
        ctx.set_option(100 + i, v)
    label, gen, sh, T = cases[key]
    label = f"{label} hl={hl} tune={tune}"
    data = gen()
    d_out = torch.empty_like(data)
    if os.environ.get("PROBE_INDEXED"):
        spf = ctx.index_segments(fl)
        d_idx = torch.empty(nf * spf, dtype=torch.int64, device="cuda")
        label += " indexed"
        comp = lambda: ctx.compress_batch_dev_indexed(data, d_off, d_len, nf, size, fl, sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, d_idx, spf, s)
        dec = lambda: ctx.decompress_batch_dev_indexed(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, d_idx, spf, s)
    elif os.environ.get("PROBE_BLOCKS"):                 # Blosc-1 multi-block frames, PROBE_BLOCKS = block size (1 = default)
        bsz = int(os.environ["PROBE_BLOCKS"]); bsz = 0 if bsz == 1 else bsz
        label += f" blocks({bsz or 65536})"
        comp = lambda: ctx.compress_blocks_batch_dev(data, d_off, d_len, nf, size, fl, sh, T, bsz, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
        dec = lambda: ctx.decompress_blocks_batch_dev(d_c, d_foff, d_flen, nf, d_out, d_off, d_len, size, fl, bsz, d_olen, d_st, s)
    else:
        comp = lambda: ctx.compress_batch_dev(data, d_off, d_len, nf, size, fl, sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
        dec = lambda: ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, size, fl, d_olen, d_st, s)
    comp(); dec(); torch.cuda.synchronize()
    ctx.kernel_stats_reset()
    tc, td = [], []
    for _ in range(iters):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); comp(); e[1].record(); dec(); e[2].record()
        torch.cuda.synchronize()
        tc.append(e[0].elapsed_time(e[1])); td.append(e[1].elapsed_time(e[2]))
    ok_c = not bool(d_st.any())
    total = int(d_tot.item())
    exact = torch.equal(d_out, data)
    st = ctx.kernel_stats()
    per = ", ".join(f"{k.replace('_kernel', '')} {v[1] / max(v[0], 1):.3f}ms" for k, v in st.items() if v[0])
    print(f"{label}: ratio {total / size:.4f} | compress {size / min(tc) / 1e6:.1f} GB/s | decompress {size / min(td) / 1e6:.1f} GB/s "
          f"| exact={exact} status_ok={ok_c and not bool(d_st.any())}\n    {per}")
    del data, d_out
