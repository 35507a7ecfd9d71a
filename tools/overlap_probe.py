"""Does running two half-batches on two streams (two contexts, two arenas) beat one batch on one stream?
(HBM-bound filter / pack / unshuffle kernels of one half under the issue-bound LZ4 kernels of the other.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

pkg = entry.load_package()
size = int(os.environ.get("PROBE_BYTES", 4 << 30))
fl = 262144
nf = size // fl
src = gen_f32(size // 4)
parts = int(os.environ.get("PROBE_PARTS", 2))


class Half:
    def __init__(self, lo, hi):
        self.ctx = pkg.Context(0)
        if os.environ.get("PROBE_ENC_CTAS"):
            self.ctx.set_option(106, int(os.environ["PROBE_ENC_CTAS"]))
        self.stream = torch.cuda.Stream()
        self.s = self.stream.cuda_stream
        self.n = hi - lo
        self.bytes = self.n * fl
        self.src = src[lo * fl:hi * fl]
        self.off = torch.arange(self.n, dtype=torch.int64, device="cuda") * fl
        self.len = torch.full((self.n,), fl, dtype=torch.int32, device="cuda")
        self.cap = self.bytes + 32 * self.n + 64
        self.c = torch.empty(self.cap, dtype=torch.uint8, device="cuda")
        self.foff = torch.empty(self.n, dtype=torch.int64, device="cuda"); self.flen = torch.empty(self.n, dtype=torch.int32, device="cuda")
        self.st = torch.empty(self.n, dtype=torch.int32, device="cuda"); self.tot = torch.empty(1, dtype=torch.int64, device="cuda")
        self.out = torch.empty_like(self.src); self.olen = torch.empty(self.n, dtype=torch.int32, device="cuda")
        self.ctx.reserve(self.bytes, self.n)

    def comp(self):
        self.ctx.compress_batch_dev(self.src, self.off, self.len, self.n, self.bytes, fl, 1, 4, self.c, self.cap, self.foff, self.flen, self.st, self.tot, self.s)

    def dec(self):
        self.ctx.decompress_batch_dev(self.c, self.foff, self.flen, self.n, 0, self.out, self.off, self.len, self.bytes, fl, self.olen, self.st, self.s)


def run(halves, label):
    for h in halves: h.comp(); h.dec()
    torch.cuda.synchronize()
    best_c = best_d = 1e9
    for _ in range(3):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        for h in halves:
            h.stream.wait_event(a); h.comp()
        evs = [h.stream.record_event() for h in halves]
        for e in evs: torch.cuda.current_stream().wait_event(e)
        b.record()
        for h in halves:
            h.stream.wait_event(b); h.dec()
        evs = [h.stream.record_event() for h in halves]
        for e in evs: torch.cuda.current_stream().wait_event(e)
        c.record(); torch.cuda.synchronize()
        best_c = min(best_c, a.elapsed_time(b)); best_d = min(best_d, b.elapsed_time(c))
    ok = all(torch.equal(h.out, h.src) for h in halves)
    print(f"{label}: compress {best_c:.2f} ms ({size / best_c / 1e6:.1f} GB/s), decompress {best_d:.2f} ms ({size / best_d / 1e6:.1f} GB/s), exact={ok}", flush=True)


run([Half(0, nf)], "one batch, one stream")
cut = [nf * i // parts for i in range(parts + 1)]
run([Half(cut[i], cut[i + 1]) for i in range(parts)], f"{parts} part-batches on {parts} streams")
