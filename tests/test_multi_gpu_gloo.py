"""CPU, world_size 2 over gloo: the N>1 host logic -- contiguous frame sharding with no
data-path collective, and the all-gather of per-frame compressed sizes that builds the global
packed-offsets table (K6).  Frame bytes here come from the oracle (the checker), because only
sizes and ordering are under test; the CUDA path itself is covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frames(nframes):
    import datagen as dg
    out = []
    for f in range(nframes):
        n = [32768, 65536, 1000, 131072, 13][f % 5]
        out.append(dg.random_bytes(n, f) if f % 3 == 0 else dg.lowent_i16((n + 1) // 2, f)[:n].copy())
    return out


def _worker(rank, world, port, nframes, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    import go_blosc_b200.parallel as par
    orc = entry.load_oracle()
    frames = _frames(nframes)
    lo, hi = par.shard_range(nframes, rank, world)
    sizes = []
    for f in range(lo, hi):                                # each rank works on its own shard only
        rc, fr = orc.compress(frames[f], orc.LZ4, 5, orc.SHUFFLE, 2)
        assert rc == 0
        sizes.append(fr.size)
    local = torch.tensor(sizes, dtype=torch.int32)
    all_sizes, counts = par.allgather_frame_sizes(local)
    offs = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int64), all_sizes.to(torch.int64)]), 0)
    q.put((rank, lo, hi, all_sizes.tolist(), counts, offs.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nframes", [37, 2, 1])
def test_sharding_and_size_allgather_world2(nframes):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import __graft_entry__ as entry
    orc = entry.load_oracle()
    want = []
    for fr in _frames(nframes):
        rc, c = orc.compress(fr, orc.LZ4, 5, orc.SHUFFLE, 2)
        want.append(int(c.size))
    results.sort()
    covered = []
    for rank, lo, hi, all_sizes, counts, offs in results:
        assert all_sizes == want                             # every rank sees the global table, in frame order
        assert counts == [r[2] - r[1] for r in results]
        assert offs == np.concatenate([[0], np.cumsum(want)]).tolist()
        covered += list(range(lo, hi))
    assert covered == list(range(nframes))                   # contiguous, disjoint, complete


def test_shard_range_properties():
    import __graft_entry__ as entry
    entry.load_package()
    import go_blosc_b200.parallel as par
    for n in (0, 1, 7, 8, 9, 32768, 100003):
        for world in (1, 2, 4, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        par.shard_range(10, 2, 2)


def _gather_worker(rank, world, port, nframes, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as entry
    entry.load_package()
    import go_blosc_b200.parallel as par
    orc = entry.load_oracle()
    frames = _frames(nframes)
    lo, hi = par.shard_range(nframes, rank, world)
    # this rank's packed output, laid out like compress_batch_dev's: frames on 16-byte boundaries
    parts, offs, lens, pos = [], [], [], 0
    for f in range(lo, hi):
        rc, fr = orc.compress(frames[f], orc.LZ4, 5, orc.SHUFFLE, 2)
        assert rc == 0
        offs.append(pos); lens.append(fr.size)
        pad = (-fr.size) % 16
        parts.append(np.concatenate([fr, np.zeros(pad, dtype=np.uint8)]))
        pos += fr.size + pad
    local = torch.from_numpy(np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8))
    buf, g_off, g_len = par.gather_packed_frames(local, pos, torch.tensor(offs, dtype=torch.int64),
                                                 torch.tensor(lens, dtype=torch.int32), dst=0)
    ok = True
    if rank == 0:
        blob = buf.numpy()
        assert g_off.numel() == nframes and g_len.numel() == nframes
        for f in range(nframes):                              # every rank's frames, in global frame order
            o, n = int(g_off[f]), int(g_len[f])
            assert o % 16 == 0
            rc, back = orc.decompress(blob[o:o + n])
            ok = ok and rc == 0 and np.array_equal(back, frames[f])
    else:
        assert buf is None
    q.put((rank, ok, g_off.tolist(), g_len.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nframes", [37, 3, 1])
def test_gather_packed_frames_world2(nframes):
    """K6 + peer copies: the ranks' packed buffers end to end on rank 0, global offsets on every rank."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in results)
    assert results[0][2] == results[1][2] and results[0][3] == results[1][3]     # the same table everywhere
