"""Multi-GPU plumbing: one process per GPU, independent frames sharded with no collective on the
data path.  The only exchange is the all-gather of per-frame compressed sizes
that is needed when a global packed-offsets table over all ranks' frames is assembled (K6);
it rides on torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference has no counterpart (no goroutines, no chunk API; blosc.go:232-233).
"""
from __future__ import annotations


def shard_range(nframes: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous frame range [lo, hi) owned by `rank`: frame f -> rank floor(f * world / nframes)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (nframes * rank + world - 1) // world
    hi = (nframes * (rank + 1) + world - 1) // world
    return lo, hi


def owner_of(frame: int, nframes: int, world: int) -> int:
    return frame * world // nframes


def allgather_frame_sizes(local_sizes, group=None):
    """All-gather of the per-frame sizes of every rank (int32/int64 1-D tensor, ragged counts
    allowed).  Returns (all_sizes, counts): the concatenation in rank order and the number of
    frames each rank contributed."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n_local = torch.tensor([local_sizes.numel()], dtype=torch.int64, device=local_sizes.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    width = max(counts) if counts else 0
    padded = torch.zeros(width, dtype=local_sizes.dtype, device=local_sizes.device)
    padded[:local_sizes.numel()] = local_sizes
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    all_sizes = torch.cat([g[:c] for g, c in zip(gathered, counts)]) if world else padded
    return all_sizes, counts


def global_frame_table(ctx, local_frame_len, group=None, stream=0):
    """Global packed-offsets table over every rank's frames.

    local_frame_len: int32 CUDA tensor of this rank's frame lengths (as written by
    compress_batch_dev).  Returns (all_len, all_off, total, my_first): the gathered lengths, their
    exclusive scan with 16-byte aligned frame starts is NOT applied here -- this is the plain
    offsets table (K5 scan, run on the device through the C ABI), the total byte count and the
    index of this rank's first frame in the global table."""
    import torch
    import torch.distributed as dist

    all_len, counts = allgather_frame_sizes(local_frame_len, group)
    n = all_len.numel()
    all_off = torch.empty(n, dtype=torch.int64, device=all_len.device)
    total = torch.zeros(1, dtype=torch.int64, device=all_len.device)
    if not all_len.is_cuda:
        raise RuntimeError("global_frame_table needs device tensors (the scan runs on the GPU; "
                           "there is no CPU fallback)")
    ctx.scan_offsets_dev(all_len.contiguous(), n, all_off, total, stream)
    rank = dist.get_rank(group)
    return all_len, all_off, total, sum(counts[:rank])


def gather_packed_frames(local_buf, local_total: int, local_frame_off, local_frame_len, dst: int = 0, group=None):
    """Materialise ONE packed buffer with every rank's frames on rank `dst` (optional peer
    copies).  Each rank holds its own packed output of compress_batch_dev: `local_buf[:local_total]`
    with frames at `local_frame_off` (16-byte aligned) of `local_frame_len` bytes.  The ranks' buffers
    are laid end to end in rank order (every local total is a multiple of 16, so frames stay aligned);
    the bytes travel point to point (NCCL send/recv over NVLink on the GPU box, gloo in the CPU tests),
    the tables with two all-gathers.  Returns (buf, frame_off, frame_len): the global offsets / lengths
    of all frames on every rank, and the packed bytes on `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = local_buf.device
    tot = torch.tensor([int(local_total)], dtype=torch.int64, device=dev)
    tots = [torch.zeros_like(tot) for _ in range(world)]
    dist.all_gather(tots, tot, group=group)
    tots = [int(t.item()) for t in tots]
    bases = [sum(tots[:r]) for r in range(world)]
    all_len, counts = allgather_frame_sizes(local_frame_len, group)
    all_off_local, _ = allgather_frame_sizes(local_frame_off.to(torch.int64), group)
    base_per_frame = torch.cat([torch.full((c,), b, dtype=torch.int64, device=dev) for c, b in zip(counts, bases)]) \
        if sum(counts) else torch.zeros(0, dtype=torch.int64, device=dev)
    frame_off = all_off_local + base_per_frame
    buf, ops = None, []
    if rank == dst:
        buf = torch.empty(max(sum(tots), 1), dtype=torch.uint8, device=dev)
        buf[bases[rank]:bases[rank] + tots[rank]] = local_buf[:tots[rank]]
        ops = [dist.P2POp(dist.irecv, buf[bases[r]:bases[r] + tots[r]], r, group) for r in range(world)
               if r != dst and tots[r]]
    elif tots[rank]:
        ops = [dist.P2POp(dist.isend, local_buf[:tots[rank]].contiguous(), dst, group)]
    if ops:                                       # one batch: the receives of all peers run concurrently
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return buf, frame_off, all_len


class NcclComm:
    """A raw ncclComm_t for the C ABI (b2b_allgather_sizes takes one; torch does not hand its own out).  Created
    through ctypes on the NCCL library the process already uses; the unique id travels over torch.distributed
    (any backend) when world > 1."""

    def __init__(self, rank: int, world: int, device: int, group=None):
        import ctypes
        import torch
        self._nccl = ctypes.CDLL("libnccl.so.2")

        class UniqueId(ctypes.Structure):          # ncclUniqueId: passed BY VALUE to ncclCommInitRank
            _fields_ = [("internal", ctypes.c_byte * 128)]

        uid = UniqueId()
        if rank == 0:
            self._check(self._nccl.ncclGetUniqueId(ctypes.byref(uid)))
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                t = t.cuda(device)
            dist.broadcast(t, 0, group=group)
            ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().tolist()), 128)
        torch.cuda.set_device(device)
        self.comm = ctypes.c_void_p()
        self._nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
        self._check(self._nccl.ncclCommInitRank(ctypes.byref(self.comm), world, uid, rank))
        self.world, self.rank = world, rank

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise RuntimeError(f"NCCL call failed with ncclResult_t {rc}")

    @property
    def ptr(self) -> int:
        return self.comm.value

    def close(self):
        import ctypes
        if self.comm:
            self._nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = None


def global_frame_table_native(ctx, comm: NcclComm, local_frame_len, align16=False, stream=0):
    """Global packed-offsets table through the C ABI alone: one ncclAllGather + one scan kernel on `stream`, no
    host synchronisation (equal frame counts per rank).  Returns (all_len, all_off, total)."""
    import torch
    n = local_frame_len.numel()
    dev = local_frame_len.device
    all_len = torch.empty(n * comm.world, dtype=torch.int32, device=dev)
    all_off = torch.empty(n * comm.world, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    ctx.allgather_sizes(comm.ptr, local_frame_len, n, comm.world, all_len, all_off, total, align16, stream)
    return all_len, all_off, total
