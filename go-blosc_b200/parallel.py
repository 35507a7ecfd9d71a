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
