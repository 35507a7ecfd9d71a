"""BASELINE.json configs other than the bench workload: C1 (README example latency), C5 (mixed
random / low-entropy int16, frame sizes 32 KiB..2 MiB).  C2 is tools/perf_probe.py, C3/C4 tools/codec_probe.py."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry

pkg = entry.load_package()
orc = entry.load_oracle()
ctx = pkg.Context(0)
s = torch.cuda.current_stream().cuda_stream

# ---- C1: 100 KB ramp, LZ4 level 5, Shuffle1 typesize 4, one frame through the host API
data = (np.arange(100000) % 256).astype(np.uint8)
fr = ctx.compress(data, 1, 5, 1, 4)
back = ctx.decompress(fr)
rc, ref = orc.compress(data, orc.LZ4, 5, orc.SHUFFLE, 4)
tc, td = [], []
for _ in range(20):
    t0 = time.perf_counter(); fr = ctx.compress(data, 1, 5, 1, 4); t1 = time.perf_counter()
    back = ctx.decompress(fr); t2 = time.perf_counter()
    tc.append(t1 - t0); td.append(t2 - t1)
print(f"C1 README example (100 000 B ramp, one frame, host API): frame {len(fr)} B (oracle {ref.size} B), header {bytes(fr[:16]).hex()}, "
      f"round trip ok={back == data.tobytes()}, oracle decodes it={orc.decompress(np.frombuffer(fr, dtype=np.uint8))[0] == 0}, "
      f"compress {1e6 * min(tc):.0f} us, decompress {1e6 * min(td):.0f} us (median {1e6 * sorted(tc)[10]:.0f} / {1e6 * sorted(td)[10]:.0f})")

# ---- C5: mixed workload, T=2 Shuffle1, frame sizes cycling 32 KiB .. 2 MiB, alternating random / low-entropy int16
total = int(os.environ.get("PROBE_BYTES", 4 << 30))
sizes_k = [32, 64, 128, 256, 512, 1024, 2048]
lens = []
acc = 0
i = 0
while acc + (sizes_k[i % 7] << 10) <= total:
    lens.append(sizes_k[i % 7] << 10); acc += lens[-1]; i += 1
nf = len(lens)
lens_np = np.array(lens, dtype=np.uint32)
offs_np = np.concatenate([[0], np.cumsum(lens_np[:-1], dtype=np.uint64)]).astype(np.uint64)
g = torch.Generator(device="cuda"); g.manual_seed(5)
low = torch.randint(0, 8, (acc // 2,), device="cuda", generator=g, dtype=torch.int16).view(torch.uint8)
rnd = torch.randint(0, 256, (acc,), device="cuda", generator=g, dtype=torch.uint8)
src = low.clone()
is_rand = torch.zeros(acc, dtype=torch.bool, device="cuda")
for f in range(0, nf, 2):                       # even frames random, odd frames low entropy
    is_rand[int(offs_np[f]):int(offs_np[f]) + lens[f]] = True
src[is_rand] = rnd[is_rand]
del low, rnd, is_rand
d_off = torch.from_numpy(offs_np.astype(np.int64)).cuda()
d_len = torch.from_numpy(lens_np.astype(np.int32)).cuda()
cap = acc + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_foff = torch.empty(nf, dtype=torch.int64, device="cuda"); d_flen = torch.empty(nf, dtype=torch.int32, device="cuda")
d_st = torch.empty(nf, dtype=torch.int32, device="cuda"); d_tot = torch.empty(1, dtype=torch.int64, device="cuda")
d_out = torch.empty_like(src); d_olen = torch.empty(nf, dtype=torch.int32, device="cuda")
ctx.reserve(acc, nf)
mx = max(lens)
comp = lambda: ctx.compress_batch_dev(src, d_off, d_len, nf, acc, mx, 1, 2, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
dec = lambda: ctx.decompress_batch_dev(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, acc, mx, d_olen, d_st, s)
comp(); dec(); torch.cuda.synchronize()
tc, td = [], []
for _ in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); comp(); e[1].record(); dec(); e[2].record(); torch.cuda.synchronize()
    tc.append(e[0].elapsed_time(e[1])); td.append(e[1].elapsed_time(e[2]))
flags = torch.stack([d_c[int(o) + 2] for o in d_foff[:14].cpu()]).cpu().numpy()
print(f"C5 mixed (T=2 Shuffle1, {nf} frames of 32 KiB..2 MiB, {acc >> 20} MiB, random/low-entropy alternating): ratio {int(d_tot.item()) / acc:.4f}, "
      f"compress {acc / min(tc) / 1e6:.1f} GB/s, decompress {acc / min(td) / 1e6:.1f} GB/s, exact={torch.equal(d_out, src)}, "
      f"status ok={not bool(d_st.any())}, memcpy flag of the first 14 frames={[int(x) >> 1 & 1 for x in flags]}")

# ---- the same C5 batch with the side-car decode index (independent 64 KiB segments, one warp per segment)
spf = ctx.index_segments(mx)
d_idx = torch.empty(nf * spf, dtype=torch.int64, device="cuda")
compi = lambda: ctx.compress_batch_dev_indexed(src, d_off, d_len, nf, acc, mx, 1, 2, d_c, cap, d_foff, d_flen, d_st, d_tot, d_idx, spf, s)
deci = lambda: ctx.decompress_batch_dev_indexed(d_c, d_foff, d_flen, nf, 0, d_out, d_off, d_len, acc, mx, d_olen, d_st, d_idx, spf, s)
compi(); d_out.zero_(); deci(); torch.cuda.synchronize()
tc, td = [], []
for _ in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); compi(); e[1].record(); deci(); e[2].record(); torch.cuda.synchronize()
    tc.append(e[0].elapsed_time(e[1])); td.append(e[1].elapsed_time(e[2]))
print(f"C5 mixed with decode index ({spf} entries per frame): ratio {int(d_tot.item()) / acc:.4f}, compress {acc / min(tc) / 1e6:.1f} GB/s, "
      f"decompress {acc / min(td) / 1e6:.1f} GB/s, exact={torch.equal(d_out, src)}, status ok={not bool(d_st.any())}")

# ---- the same C5 batch as Blosc-1 multi-block frames (64 KiB blocks: every block an independent stream, no side-car)
compb = lambda: ctx.compress_blocks_batch_dev(src, d_off, d_len, nf, acc, mx, 1, 2, 0, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
decb = lambda: ctx.decompress_blocks_batch_dev(d_c, d_foff, d_flen, nf, d_out, d_off, d_len, acc, mx, 0, d_olen, d_st, s)
compb(); d_out.zero_(); decb(); torch.cuda.synchronize()
tc, td = [], []
for _ in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); compb(); e[1].record(); decb(); e[2].record(); torch.cuda.synchronize()
    tc.append(e[0].elapsed_time(e[1])); td.append(e[1].elapsed_time(e[2]))
print(f"C5 mixed as multi-block frames (64 KiB blocks): ratio {int(d_tot.item()) / acc:.4f}, compress {acc / min(tc) / 1e6:.1f} GB/s, "
      f"decompress {acc / min(td) / 1e6:.1f} GB/s, exact={torch.equal(d_out, src)}, status ok={not bool(d_st.any())}")
