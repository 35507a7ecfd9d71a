import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import __graft_entry__ as entry
import datagen as dg
pkg = entry.load_package(); orc = entry.load_oracle()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream
def run(lens, gen, sh, T, indexed=False, label=""):
    lens = np.asarray(lens, dtype=np.uint32); nf = len(lens)
    offs = np.concatenate([[0], np.cumsum(lens[:-1].astype(np.uint64))]).astype(np.uint64)
    total = int(offs[-1] + lens[-1]); mx = int(lens.max())
    src = gen(total)
    d_off = torch.from_numpy(offs.astype(np.int64)).cuda(); d_len = torch.from_numpy(lens.astype(np.int32)).cuda()
    cap = total + 32 * nf + 64
    d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_foff = torch.empty(nf, dtype=torch.int64, device="cuda"); d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
    d_st = torch.empty(nf, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
    d_out = torch.zeros(total, dtype=torch.uint8, device="cuda"); d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
    if indexed:
        spf = ctx.index_segments(mx); d_idx = torch.empty(nf * spf, dtype=torch.int64, device="cuda")
        ctx.compress_batch_dev_indexed(src, d_off, d_len, nf, total, mx, sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, d_idx, spf, s)
    else:
        ctx.compress_batch_dev(src, d_off, d_len, nf, total, mx, sh, T, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
    torch.cuda.synchronize(); okc = not bool(d_st.any())
    if indexed:
        ctx.decompress_batch_dev_indexed(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, total, mx, d_olen, d_st, d_idx, spf, s)
    else:
        ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, total, mx, d_olen, d_st, s)
    torch.cuda.synchronize()
    ok = okc and not bool(d_st.any()) and torch.equal(d_out, src)
    # oracle decode of a few frames
    foff, flen = d_foff.cpu().numpy(), d_flen.cpu().numpy()
    for f in sorted(set([0, nf // 2, nf - 1])):
        fr = d_c[int(foff[f]):int(foff[f]) + int(flen[f])].cpu().numpy()
        rc, back = orc.decompress(fr)
        ok = ok and rc == 0 and np.array_equal(back, src[int(offs[f]):int(offs[f]) + int(lens[f])].cpu().numpy())
    print(f"{label}: nf={nf} total={total >> 20} MiB max={mx} ratio={int(d_tot.item()) / total:.4f} indexed={indexed} ok={ok}", flush=True)
    del src, d_c, d_out
    return ok
g = torch.Generator(device="cuda"); g.manual_seed(3)
def mixed(n):
    a = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.uint8)
    r = torch.randint(0, 256, (n,), device="cuda", generator=g, dtype=torch.uint8)
    sel = (torch.arange(n, device="cuda") // 300000) % 3 == 0
    a[sel] = r[sel]
    return a
allok = True
allok &= run([300 << 20], mixed, 0, 1, label="one 300 MiB frame")
allok &= run([300 << 20], mixed, 1, 4, indexed=True, label="one 300 MiB frame, indexed, shuffle 4")
allok &= run([(1 << 30) + 12345], mixed, 0, 1, label="one 1 GiB+ frame")
rng = np.random.default_rng(1)
allok &= run(rng.integers(1, 200, 300000), mixed, 1, 2, label="300k tiny frames")
allok &= run(rng.integers(1, 200, 300000), mixed, 0, 1, indexed=True, label="300k tiny frames indexed")
allok &= run(rng.integers(60000, 70000, 20000), mixed, 2, 8, label="20k frames around 64 KiB, bitshuffle 8")
print("ALL OK" if allok else "FAILURES")
