"""N ranks (torchrun, NCCL): each compresses its shard on its GPU, the packed buffers are gathered on rank 0
with go_blosc_b200.parallel.gather_packed_frames (NVLink send/recv), rank 0 decompresses ALL frames from the
gathered buffer with the global offsets table and compares with the gathered sources."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from tools.perf_probe_lib import gen_f32

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pkg = entry.load_package()
import go_blosc_b200.parallel as par
ctx = pkg.Context(local)
s = torch.cuda.current_stream().cuda_stream
fl = 262144
size = int(os.environ.get("PROBE_BYTES", 256 << 20))
nf = size // fl
torch.manual_seed(rank)
src = gen_f32(size // 4) if rank % 2 == 0 else torch.randint(0, 8, (size // 2,), device=dev, dtype=torch.int16).view(torch.uint8)
d_off = torch.arange(nf, dtype=torch.int64, device=dev) * fl
d_len = torch.full((nf,), fl, dtype=torch.int32, device=dev)
cap = size + 32 * nf + 64
d_c = torch.empty(cap, dtype=torch.uint8, device=dev)
d_foff = torch.empty(nf, dtype=torch.int64, device=dev)
d_flen = torch.empty(nf, dtype=torch.int32, device=dev)
d_st = torch.empty(nf, dtype=torch.int32, device=dev)
d_tot = torch.zeros(1, dtype=torch.int64, device=dev)
ctx.compress_batch_dev(src, d_off, d_len, nf, size, fl, 1, 4 if rank % 2 == 0 else 2, d_c, cap, d_foff, d_flen, d_st, d_tot, s)
torch.cuda.synchronize()
assert not d_st.any()
total = int(d_tot.item())
for it in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    buf, g_off, g_len = par.gather_packed_frames(d_c, total, d_foff, d_flen, dst=0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
srcs = [torch.empty_like(src) for _ in range(world)]
dist.all_gather(srcs, src)
if rank == 0:
    n_all = nf * world
    want = torch.cat(srcs)
    out = torch.empty(size * world, dtype=torch.uint8, device=dev)
    o_off = torch.arange(n_all, dtype=torch.int64, device=dev) * fl
    o_len = torch.full((n_all,), fl, dtype=torch.int32, device=dev)
    o_olen = torch.empty(n_all, dtype=torch.int32, device=dev)
    o_st = torch.empty(n_all, dtype=torch.int32, device=dev)
    ctx.decompress_batch_dev(buf, g_off.contiguous(), g_len.contiguous(), n_all, 0, out, o_off, o_len, size * world, fl, o_olen, o_st, s)
    torch.cuda.synchronize()
    ok = (not o_st.any()) and torch.equal(out, want)
    print(f"gather_packed_frames on {world} ranks: {buf.numel() / 1e6:.1f} MB on rank 0 in {dt * 1e3:.2f} ms "
          f"({(buf.numel() - total) / dt / 1e9:.1f} GB/s received), all {n_all} frames decode from the gathered buffer: {ok}")
    assert ok
dist.barrier()
dist.destroy_process_group()
