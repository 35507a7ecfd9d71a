"""Host-pointer path timing: compress_batch and decompress_batch separately; PROBE_MEM=pinned|pageable|both,
PROBE_THREADS=comma list of B2B_OPT_HOST_THREADS values, PROBE_NOSTAGE=1 adds the driver-staged path for comparison."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32
pkg = entry.load_package()
ctx = pkg.Context(0)
FRAME = 262144
total = int(os.environ.get("PROBE_BYTES", 2 << 30))
nf = total // FRAME
src = gen_f32(total // 4)
offs = np.arange(nf, dtype=np.uint64) * FRAME
lens = np.full(nf, FRAME, dtype=np.uint32)
modes = {"pinned": ["pinned"], "pageable": ["pageable"], "both": ["pinned", "pageable"]}[os.environ.get("PROBE_MEM", "pinned")]
for mem in modes:
    pin = mem == "pinned"
    h_src = torch.empty(total, dtype=torch.uint8, pin_memory=pin); h_src.copy_(src)
    h_comp = torch.empty(total + 32 * nf + 64, dtype=torch.uint8, pin_memory=pin)
    h_out = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
    if not pin: h_comp.zero_(); h_out.zero_()          # touch the pages once
    a_src, a_comp, a_out = h_src.numpy(), h_comp.numpy(), h_out.numpy()
    blocks = os.environ.get("PROBE_BLOCKS")          # Blosc-1 multi-block frames of this block size (1 = default)
    if blocks:
        bsz = 0 if int(blocks) == 1 else int(blocks)
        comp = lambda: ctx.compress_blocks_batch(a_src, offs, lens, pkg.Shuffle.Shuffle1, 4, bsz, dst=a_comp)
        dec = lambda foff, flen: ctx.decompress_blocks_batch(a_comp, foff, flen, offs, total, bsz, dst=a_out)
    else:
        comp = lambda: ctx.compress_batch(a_src, offs, lens, pkg.Shuffle.Shuffle1, 4, dst=a_comp)
        dec = lambda foff, flen: ctx.decompress_batch(a_comp, foff, flen, offs, total, dst=a_out)
    configs = []
    for st in [int(x) for x in os.environ.get("PROBE_STAGES", "128").split(",")]:
        for th in [int(x) for x in os.environ.get("PROBE_THREADS", "0").split(",")]:
            configs.append((st, th, 0))
    if os.environ.get("PROBE_NOSTAGE") and mem == "pageable":
        configs.append((128, 0, 1))
    for st, th, nostage in configs:
        ctx.set_option(pkg.OPT_HOST_STAGE_BYTES, st << 20)
        ctx.set_option(pkg.OPT_HOST_THREADS, th)
        ctx.set_option(pkg.OPT_NO_HOST_STAGING, nostage)
        tc, td = [], []
        for it in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _, foff, flen, stt, tot = comp()
            t1 = time.perf_counter()
            _, olen, st2 = dec(foff, flen)
            t2 = time.perf_counter()
            if it: tc.append(t1 - t0); td.append(t2 - t1)
        print(f"{mem} stage {st} MiB threads {th} nostage {nostage}: compress {1e3 * min(tc):.1f} ms (H2D {total / 1e6:.0f} MB, D2H {tot / 1e6:.0f} MB -> {total / min(tc) / 1e9:.1f} GB/s in), "
              f"decompress {1e3 * min(td):.1f} ms ({total / min(td) / 1e9:.1f} GB/s out), ok={bool(torch.equal(h_out, h_src))}")
