// blocks.cuh -- opt-in Blosc-1 MULTI-BLOCK frames.
//
// The reference writes and reads one block per frame and never reads Options.BlockSize
// (blosc.go:227-234, 320-374), so this container has no reference counterpart and stays
// behind its own entry points (b2b_*_blocks_*).  Layout = the published Blosc-1 chunk:
//
//   0..15   version 2 | versionlz 1 | flags | typesize | nbytes | blocksize | cbytes
//           flags: 0x1 byte shuffle, 0x2 stored ("memcpyed"), 0x4 bit shuffle, 0x10 blocks not split,
//           bits 5..7 compressor format (1 = LZ4)
//   16..    bstarts: int32[nblocks], offset of each block from the start of the frame
//   block   nsplits streams, each int32 csize + csize bytes; csize == stream length: stored raw,
//           otherwise one LZ4 block.  This encoder never splits (flag 0x10); the decoder reads both.
//   stored  header + the nbytes original bytes
//
// The filter runs per block.  A block is a frame of its own to K1/K2/K3/K4, so compression is
// those kernels over a table of block slots plus the small kernels here: geometry, slot -> frame
// map, frame sizes, and a pack kernel that writes header, bstarts and streams at their packed
// place.  Decompression: header pass, slot map, a stream descriptor per block, then K4's two halves
// (parse kernel -> sequence records -> copy kernel, lz4_kernels.cuh) over the block streams and
// K1/K2 inverse over the block slots.
#pragma once
#include "common.cuh"
#include "lz4_encode.cuh"
#include "lz4_kernels.cuh"

namespace b2b {

constexpr uint32_t kB1MinBuffer = 128;            // smaller buffers are stored; shorter streams are never split
constexpr uint32_t kB1MaxSplits = 16;
constexpr uint32_t kB1MaxBuffer = 0x7FFFFFFFu - 16u;
constexpr uint32_t kB1DontSplit = 0x10, kB1Lz4Format = 1, kB1DefaultBlock = 65536;
constexpr uint32_t kNoOwner = 0xFFFFFFFFu;

// block size actually used for a buffer of n bytes: whole elements, at most the buffer
__host__ __device__ __forceinline__ uint32_t b1_blocksize(uint64_t n, uint32_t T, uint32_t req) {
    uint64_t b = req ? req : kB1DefaultBlock;
    if (b > n) b = n;
    if (b > T) b = b / T * T;
    if (b < T) b = T;
    return (uint32_t)b;
}
__host__ __device__ __forceinline__ uint32_t b1_typesize(int64_t typesize) {
    return (typesize <= 0 || typesize > 255) ? 1u : (uint32_t)typesize;
}
__device__ __forceinline__ uint32_t b1_splits(uint32_t flags, uint32_t T, uint32_t bs, bool partial) {
    if ((flags & kB1DontSplit) || partial) return 1;
    return (T <= kB1MaxSplits && bs / T >= kB1MinBuffer) ? T : 1u;
}

// ---- compress side --------------------------------------------------------------------------------
struct BlocksGeomArgs {
    const uint32_t *src_len;
    uint32_t nframes, typesize, blocksize_req;
    uint32_t *bs, *nblk;        // out, per frame (0 blocks: empty or too large, see blocks_frame_kernel)
};
__global__ void blocks_geom_kernel(BlocksGeomArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    const uint32_t n = a.src_len[f];
    uint32_t b = 0, k = 0;
    if (n != 0 && n <= kB1MaxBuffer) {
        b = b1_blocksize(n, a.typesize, a.blocksize_req);
        k = (uint32_t)(((uint64_t)n + b - 1) / b);
    }
    a.bs[f] = b; a.nblk[f] = k;
}

// slot t of the block table belongs to the frame whose [blk_base, blk_base + nblk) holds it
__device__ __forceinline__ uint32_t slot_owner(const uint64_t *blk_base, const uint32_t *nblk, uint32_t nframes,
                                               uint32_t t) {
    uint32_t lo = 0, hi = nframes;               // blk_base[0] == 0 <= t
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (blk_base[mid] <= t) lo = mid; else hi = mid;
    }
    return (uint64_t)t - blk_base[lo] < nblk[lo] ? lo : kNoOwner;
}

struct BlocksExpandArgs {
    const uint64_t *blk_base;   // exclusive scan of nblk
    const uint32_t *nblk, *bs;
    const uint64_t *src_off;
    const uint32_t *src_len;
    uint32_t nframes, nslots;
    uint32_t *owner;            // out, per slot (kNoOwner: unused)
    uint64_t *blk_off;          // out: the block as a frame of its own
    uint32_t *blk_len;
};
__global__ void blocks_expand_kernel(BlocksExpandArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nslots) return;
    const uint32_t f = slot_owner(a.blk_base, a.nblk, a.nframes, t);
    uint64_t off = 0; uint32_t len = 0;
    if (f != kNoOwner) {
        const uint64_t o = (uint64_t)(t - a.blk_base[f]) * a.bs[f];
        off = a.src_off[f] + o;
        len = (uint32_t)min((uint64_t)a.bs[f], (uint64_t)a.src_len[f] - o);
    }
    a.owner[t] = f; a.blk_off[t] = off; a.blk_len[t] = len;
}

struct BlocksFrameArgs {
    const uint32_t *src_len, *nblk;
    const uint64_t *blk_base;
    const uint64_t *blk_pos;    // exclusive scan of (4 + stored bytes) over nslots + 1 slots
    uint32_t nframes, base_flags;
    uint32_t nslots;            // slots the table holds (sized from the caller's total_src_bytes)
    const uint32_t *blk_status; // per slot: kEDstTooSmall when the block did not fit the scratch
    uint32_t *frame_len, *frame_flags, *status;
};
__global__ void blocks_frame_kernel(BlocksFrameArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    const uint32_t n = a.src_len[f];
    uint32_t st = kOk, len = 0, flags = a.base_flags;
    if (n == 0) st = kEInvalidData;                       // as blosc.go:269-271 for the one-block frame
    else if (n > kB1MaxBuffer) st = kEDataTooLarge;
    else if (a.blk_base[f] + a.nblk[f] > a.nslots) st = kEDstTooSmall;   // total_src_bytes was not a bound
    else {
        const uint64_t b0 = a.blk_base[f], k = a.nblk[f];
        for (uint64_t i = 0; i < k; i++) if (a.blk_status[b0 + i] == kEDstTooSmall) st = kEDstTooSmall;
        if (st == kOk) {
            const uint64_t total = 16ull + 4ull * k + (a.blk_pos[b0 + k] - a.blk_pos[b0]);
            const bool stored = n < kB1MinBuffer || total > (uint64_t)n + 16ull;   // Blosc-1's bound: nbytes + 16
            if (stored) { flags |= 0x2u; len = n + 16u; } else len = (uint32_t)total;
        }
    }
    a.frame_len[f] = len; a.frame_flags[f] = flags; a.status[f] = st;
}

struct BlocksPackArgs {
    PackArgs p;                 // block-level tables: "frame" = block slot
    const uint8_t *orig;        // the caller's bytes: what a stored frame holds
    const uint32_t *owner;
    const uint64_t *blk_base;
    const uint32_t *nblk, *bs;
    const uint64_t *blk_pos;
    const uint64_t *src_off;    // frame level
    const uint32_t *src_len;
    const uint64_t *frame_off;
    const uint32_t *frame_len, *frame_flags, *frame_status;
    uint8_t *dst;
    uint32_t typesize;
};
__device__ __forceinline__ void wr32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
// one CTA per block slot
__global__ void __launch_bounds__(kFilterThreads, 8) blocks_pack_kernel(BlocksPackArgs a) {
    const uint32_t t = blockIdx.x;
    const uint32_t f = a.owner[t];
    if (f == kNoOwner || a.frame_status[f] != 0) return;
    const uint32_t local = (uint32_t)(t - a.blk_base[f]);
    const uint32_t n = a.src_len[f], bs = a.bs[f], flags = a.frame_flags[f], k = a.nblk[f];
    uint8_t *fo = a.dst + a.frame_off[f];                 // 16-byte aligned
    if (local == 0 && threadIdx.x == 0) {
        uint4 h;
        h.x = 2u | (1u << 8) | (flags << 16) | (a.typesize << 24);
        h.y = n; h.z = bs; h.w = a.frame_len[f];
        *reinterpret_cast<uint4 *>(fo) = h;
    }
    const uint32_t len = a.p.src_len[t];
    if (flags & 0x2u) {                                   // stored frame: the original bytes
        const uint64_t o = (uint64_t)local * bs;
        cta_copy(fo + 16 + o, a.orig + a.src_off[f] + o, len);
        return;
    }
    const uint64_t pos = 16ull + 4ull * k + (a.blk_pos[t] - a.blk_pos[a.blk_base[f]]);
    const uint32_t c = a.p.comp_len[t];
    if (threadIdx.x == 0) {
        reinterpret_cast<uint32_t *>(fo + 16)[local] = (uint32_t)pos;     // bstarts
        wr32(fo + pos, c);
    }
    uint8_t *out = fo + pos + 4;
    if (a.p.flags[t] & 0x2u) {                            // incompressible stream: the filtered bytes
        cta_copy(out, a.p.in + a.p.src_off[t], len);
        return;
    }
    const uint32_t nseg = seg_count(len);
    const uint64_t base = a.p.seg_base[t];
    for (uint32_t s = 0; s < nseg; s++) cta_pack_lz4_segment(a.p, t, s, len, nseg, base, out);
}

// ---- decompress side ------------------------------------------------------------------------------
struct BlocksInfoArgs {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const uint32_t *frame_len, *dst_cap;
    uint32_t nframes, slice;    // slice: the caller's block size (stored frames are copied in slices of it)
    uint32_t slice_min;         // smallest block size a frame written with `slice` can have (whole elements): sizes the slot table
    uint32_t *nblk, *bs;        // out: slots of frame f, bytes per slot
    uint32_t *out_len, *status;
};
__global__ void blocks_info_kernel(BlocksInfoArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    const uint8_t *fr = a.frames + a.frame_off[f];
    uint32_t flags = 0, vlz = 0, T = 0, n = 0, ncomp = 0;
    uint32_t st = check_header(fr, a.frame_len[f], flags, vlz, T, n, ncomp);
    uint32_t k = 0, b = 0;
    if (st == kOk && n > kB1MaxBuffer) st = kEInvalidData;
    if (st == kOk && n != 0) {
        const uint32_t bsz = rd32(fr + 8);
        if (T == 0 || bsz == 0) st = kEInvalidData;
        else if (flags & 0x2u) {
            if (ncomp != n + 16u) st = kESizeMismatch;
            else { b = a.slice; k = (uint32_t)(((uint64_t)n + b - 1) / b); }
        } else {
            const uint32_t fmt = flags >> 5;
            if (fmt > 4) st = kEInvalidCodec;
            else if (fmt != kB1Lz4Format) st = kEUnsupported;           // blosclz, snappy, zlib, zstd
            else if (vlz != 1) st = kEInvalidCodec;
            else {
                b = bsz; k = (uint32_t)(((uint64_t)n + b - 1) / b);
                if (16ull + 4ull * k > ncomp) st = kEInvalidData;
                else if (k > n / a.slice_min + 2u) st = kEUnsupported;  // blocks smaller than the table was sized for
            }
        }
        if (st == kOk && a.dst_cap[f] < n) st = kEDstTooSmall;
    }
    if (st != kOk) { k = 0; b = 0; }
    a.nblk[f] = k; a.bs[f] = b;
    a.out_len[f] = st == kOk ? n : 0u;
    a.status[f] = st;
}

struct BlocksOwnerArgs {
    const uint64_t *blk_base;
    const uint32_t *nblk;
    uint32_t nframes, nslots;
    uint32_t *owner;
};
__global__ void blocks_owner_kernel(BlocksOwnerArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.nslots) a.owner[t] = slot_owner(a.blk_base, a.nblk, a.nframes, t);
}

struct BlocksDecodeArgs {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const uint32_t *owner;
    const uint64_t *blk_base;
    const uint32_t *bs;
    uint8_t *dst, *scratch;     // blocks that still need their inverse filter are decoded into scratch
    const uint64_t *dst_off;
    uint32_t *status;           // per frame; a failing block raises it
    uint32_t nslots;
    // per slot, out of blocks_streams_kernel: the block's stream for the parse / copy kernels of
    // lz4_kernels.cuh (StreamArgs) ...
    uint64_t *strm_src;         // offset of the stream in `frames`
    uint32_t *strm_clen, *strm_kind;
    // ... and the block as a frame of its own for the inverse filter (slot_len 0: nothing to un-filter)
    uint64_t *slot_off;
    uint32_t *slot_len;         // also the stream's output bytes
    FrameMeta *slot_meta;
};

// One thread per block slot: where its stream is and what to do with it.  kind: 0 nothing, 1 stored,
// 2 one LZ4 block, 3 a block split into typesize streams (blocks_decode_kernel walks those); | 4 when
// the block goes to scratch because it still needs its inverse filter.
__global__ void blocks_streams_kernel(BlocksDecodeArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nslots) return;
    const uint32_t f = a.owner[t];
    uint64_t src = 0, d_off = 0;
    uint32_t clen = 0, kind = 0, cap = 0;
    FrameMeta m; m.mode = 0; m.typesize = 0;
    if (f != kNoOwner && a.status[f] == kOk) {
        const uint8_t *fr = a.frames + a.frame_off[f];
        const uint32_t flags = fr[2], T = fr[3], n = rd32(fr + 4), ncomp = rd32(fr + 12);
        const uint32_t bs = a.bs[f];
        const uint32_t local = (uint32_t)(t - a.blk_base[f]);
        const uint64_t o = (uint64_t)local * bs;
        const uint32_t bsize = (uint32_t)min((uint64_t)bs, (uint64_t)n - o);
        d_off = a.dst_off[f] + o; cap = bsize;
        if (flags & 0x2u) {                                 // a slice of a stored frame
            src = a.frame_off[f] + 16 + o; clen = bsize; kind = 1;
        } else {
            const uint32_t k = (uint32_t)(((uint64_t)n + bs - 1) / bs);
            const uint32_t mode = (flags & 0x4u) ? 2u : ((flags & 0x1u) ? 1u : 0u);
            const bool active = mode != 0 && T > 1 && bsize >= T;
            if (active) { m.mode = mode; m.typesize = T; }
            if (b1_splits(flags, T, bs, bsize != bs) > 1) kind = 3;
            else {
                const uint64_t pos = rd32(fr + 16 + 4ull * local);
                bool ok = pos >= 16ull + 4ull * k && pos + 4 <= ncomp;
                if (ok) {
                    clen = rd32(fr + pos);
                    ok = clen != 0 && clen <= 0x7FFFFFFFu && pos + 4 + clen <= ncomp;
                }
                if (ok) { src = a.frame_off[f] + pos + 4; kind = clen == bsize ? 1u : 2u; }
                else atomicMax(a.status + f, (uint32_t)kEDecompressionFailed);
            }
            if (kind != 0 && active) kind |= 4u;
        }
    }
    a.strm_src[t] = src; a.strm_clen[t] = clen; a.strm_kind[t] = kind;
    a.slot_off[t] = d_off; a.slot_len[t] = cap; a.slot_meta[t] = m;
}

// split blocks only (kind 3): one warp walks the block's typesize streams with the fused decoder
__global__ void __launch_bounds__(kCodecThreads, 6) blocks_decode_kernel(BlocksDecodeArgs a) {
    __shared__ SeqTable seq_tables[kCodecWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * kCodecWarps + warp;
    if (t >= a.nslots || (a.strm_kind[t] & 3u) != 3u) return;
    const uint32_t f = a.owner[t];
    const uint8_t *fr = a.frames + a.frame_off[f];
    const uint32_t T = fr[3], n = rd32(fr + 4), ncomp = rd32(fr + 12);
    const uint32_t bs = a.bs[f];
    const uint32_t local = (uint32_t)(t - a.blk_base[f]);
    const uint32_t k = (uint32_t)(((uint64_t)n + bs - 1) / bs);
    uint8_t *out = ((a.strm_kind[t] & 4u) ? a.scratch : a.dst) + a.slot_off[t];
    const uint32_t ns = T, ne = bs / T;                     // only whole blocks are split
    uint64_t pos = rd32(fr + 16 + 4ull * local);
    bool ok = pos >= 16ull + 4ull * k;
    for (uint32_t s = 0; ok && s < ns; s++) {
        if (pos + 4 > ncomp) { ok = false; break; }
        const uint32_t c = rd32(fr + pos);
        pos += 4;
        if (c == 0 || c > 0x7FFFFFFFu || pos + c > ncomp) { ok = false; break; }
        if (c == ne) warp_copy(out + (uint64_t)s * ne, fr + pos, ne, lane);
        else ok = warp_lz4_decode(fr + pos, c, out + (uint64_t)s * ne, ne, &seq_tables[warp], lane) == (int64_t)ne;
        pos += c;
        __syncwarp();
    }
    if (!ok && lane == 0) atomicMax(a.status + f, (uint32_t)kEDecompressionFailed);
}

// frames that failed in a block produce nothing; a frame whose blocks did not all get a slot (the
// caller's total_dst_bytes was not a bound) must not pass for decoded
__global__ void blocks_finish_kernel(uint32_t *status, uint32_t *out_len, const uint64_t *blk_base,
                                     const uint32_t *nblk, uint32_t nslots, uint32_t nframes) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    if (status[f] == kOk && blk_base[f] + nblk[f] > nslots) status[f] = kEDstTooSmall;
    if (status[f] != kOk) out_len[f] = 0;
}

}  // namespace b2b
