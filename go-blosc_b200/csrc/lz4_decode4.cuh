// lz4_decode4.cuh -- K4, fourth arrangement: a FEW LARGE frames, every output byte its own thread (pointer jumping).
//
// Replaces lz4Codec.Decompress (codec.go:77-84 -> pierrec UncompressBlock) for the call the reference's API
// actually makes: Decompress(one frame) (blosc.go:291-303), where the frame may hold up to 4 GiB.  One LZ4 block
// is one serial token chain and every match may read what the match before it wrote, so the engines that walk a
// frame in stream order (one warp per frame, lz4_kernels.cuh; one CTA per frame, lz4_decode2.cuh) decode ONE
// large frame at 0.15-0.35 GB/s however many SMs idle next to them.  This engine removes the order:
//
//   parse / stitch   (lz4_decode2.cuh) one thread per 8 KiB of stream: afterwards every sequence knows its place in
//                    the stream AND in the output.
//   map              one CTA per chunk, its records 256 at a time: literals go from the stream to the output, and
//                    every output byte gets a SOURCE INDEX S[x]: itself for a literal, x - offset for a match byte
//                    (a match that overlaps itself -- a run, a short period -- points into its first period, which
//                    removes the longest chains at the source).  Bytes are dealt out to the threads of the CTA through a
//                    prefix sum of the records' lengths, so a thread's cost does not depend on sequence lengths; literal
//                    runs and matches of 4 KiB or more go to a queue that the whole grid works off in 16 KiB slices.
//   jump rounds      S[x] <- S[S[x]] in place until S[x] is a literal for every x: ceil(log2(depth)) rounds for chains
//                    of depth `depth` (a match of a match of a match ...).  Blocks of 4 KiB whose bytes all point at
//                    literals are flagged and skipped; a round whose predecessor changed nothing returns at once
//                    (the launches are stream-ordered, the host never waits).
//   gather           out[x] = out[S[x]] for the match bytes.
//
// Everything is validated in the map kernel with the rules of the other engines; a frame with ANY malformed or
// over-capacity record is handed to the tile engine of lz4_decode2.cuh untouched (state 2), which reports the
// reference's status for it (blosc.go:410-413, 429-431).  Nothing here writes outside a frame's [0, dcap).
//
// Cost: 4 bytes of scratch per output byte and ~12 bytes of traffic per byte and round, which only pays when there
// are too few frames to fill the device in stream order: the host picks this engine for batches whose frames are
// few and large (b2b.cu), e.g. the single-frame b2b_decompress call.
#pragma once
#include "lz4_decode2.cuh"

namespace b2b {

constexpr int kJumpThreads = 256;
constexpr uint32_t kJumpBlock = 4096;      // output bytes of one work item of the rounds / the gather
constexpr uint32_t kJumpLong = 4096;       // literal runs / matches from here on go through the queue
constexpr uint32_t kJumpSlice = 16384;     // bytes of one queue slice
constexpr uint32_t kJumpRounds = 32;       // chains are shorter than 2^32

struct JumpLong {
    uint32_t frame, kind;                  // kind 0: literals (src = stream position), 1: match (src = offset)
    uint32_t pos, len, src, pad;           // pos: output position of the first byte
};

struct JumpArgs {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const FrameDec *fd;
    uint32_t nframes;
    uint8_t *dst, *scratch;
    const uint64_t *dst_off;
    const uint64_t *chunk_base;
    const uint64_t *total_chunks;
    uint64_t table_chunks;
    const ChunkDesc *desc;
    const uint32_t *last_chunk;
    const uint2 *table;
    const uint32_t *fallback;
    uint32_t *S;                // one entry per byte of dst: S[dst_off[f] + i] (values are such indices, too)
    uint32_t *state;            // per frame: 0 not this engine's, 1 this engine's, 2 malformed / over capacity: the tile engine's
    uint32_t *total;            // per frame: output bytes (from the closing token), ~0 until it was seen
    JumpLong *longq;
    uint32_t *nlong;
    uint32_t long_cap;
    uint8_t *blockdone;         // [nframes][blocks_per_frame]: 1 every byte points at a literal, 2 every byte IS a literal
    uint32_t blocks_per_frame;
    uint32_t blocks_grid;       // CTAs per frame of the rounds / the gather (they stride over the frame's blocks)
    uint32_t *changed;          // [kJumpRounds]: round r found something to do
    uint32_t *out_len, *status;
    FrameMeta *meta;
    uint32_t chunk_shift = kChunkShift;
};

// ---- select: which frames this engine takes ------------------------------------------------------------------
__global__ void lz4_jump_select_kernel(JumpArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f == 0) {
        *a.nlong = 0;
        for (uint32_t r = 0; r < kJumpRounds; r++) a.changed[r] = 0;
    }
    if (f >= a.nframes) return;
    a.state[f] = (a.fd[f].kind == 2 && !a.fallback[f] && a.last_chunk[f] != 0xFFFFFFFFu) ? 1u : 0u;
    a.total[f] = 0xFFFFFFFFu;
}

// length bytes behind a token whose nibble is 15 for the length v >= 15
__device__ __forceinline__ uint32_t len_ext_bytes_dec(uint32_t v) { return (v - 15u) / 255u + 1u; }

// source index of byte m of a match that starts at output position oM: a match that overlaps itself repeats the
// `off` bytes in front of it
__device__ __forceinline__ uint32_t jump_match_src(uint32_t oM, uint32_t off, bool ovl, uint32_t m) {
    return ovl ? oM - off + m % off : oM + m - off;
}

// ---- map: literals out, one source index per output byte ---------------------------------------------------------
__global__ void __launch_bounds__(kJumpThreads) lz4_jump_map_kernel(JumpArgs a) {
    __shared__ uint32_t s_cs[kJumpThreads + 1];       // exclusive prefix sum of the bytes the CTA handles itself
    __shared__ uint32_t s_oL[kJumpThreads], s_ll[kJumpThreads], s_lls[kJumpThreads], s_lit[kJumpThreads], s_off[kJumpThreads];
    __shared__ uint32_t s_wsum[kJumpThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint64_t nchunks = *a.total_chunks;
    if (nchunks > a.table_chunks) nchunks = a.table_chunks;
    for (uint64_t g = blockIdx.x; g < nchunks; g += gridDim.x) {
        uint32_t lo = 0, hi = a.nframes;                  // last frame whose first chunk is <= g
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (a.chunk_base[mid] <= g) lo = mid; else hi = mid;
        }
        const uint32_t f = lo;
        if (a.state[f] == 0) continue;
        const uint32_t k = (uint32_t)(g - a.chunk_base[f]);
        if (k > a.last_chunk[f]) continue;
        const ChunkDesc D = a.desc[g];
        if (D.count == 0) continue;
        const FrameDec d = a.fd[f];
        const uint2 *rec = a.table + g * chunk_slot_records(a.chunk_shift) + D.start;
        const uint32_t X = (uint32_t)a.dst_off[f];
        uint8_t *out = (d.mode ? a.scratch : a.dst) + a.dst_off[f];
        const uint8_t *__restrict__ src = a.frames + a.frame_off[f] + 16;
        for (uint32_t r = 0; r < D.count; r += kJumpThreads) {
            const uint32_t i = r + tid;
            uint32_t oL = 0, ll = 0, lit = 0, off = 0, lls = 0, mls = 0;
            bool ovl = false;
            if (i < D.count) {
                // the record's place and its token: the checks of lz4_copy2_kernel, in the same order
                const uint2 rc = rec[i], nx = rec[i + 1];
                const long long o = (i < D.split ? D.base_a : D.base_b) + (long long)rc.y;
                const long long on = (i + 1 < D.split ? D.base_a : D.base_b) + (long long)nx.y;
                const uint32_t kind = i + 1 == D.count ? D.end : (uint32_t)kEndCont;
                bool bad = o < 0 || o > (long long)d.dcap || on < o || (kind != kEndCont && kind != kEndFinal);
                bool lit_ok = false, match_ok = false;
                uint64_t ml = 0;
                if (!bad) {
                    uint32_t p = rc.x;
                    const uint32_t tok = src[p++];
                    ll = tok >> 4;
                    const uint64_t span = (uint64_t)(on - o);         // output bytes of the sequence: ll + ml
                    if (ll == 15u && span >= 4096u && (kind == kEndFinal || (tok & 15u) != 15u) &&
                        span >= (kind == kEndFinal ? 15u : (tok & 15u) + 4u + 15u)) {
                        // A long run: its length follows from the record table (the parse walked and checked every length
                        // byte: 255, ..., 255, last) -- the closing token is all literals, and a match nibble under 15 is
                        // the whole match length -- so a million length bytes are not walked a second time by one thread
                        ll = (uint32_t)(span - (kind == kEndFinal ? 0u : (tok & 15u) + 4u));
                        p += len_ext_bytes_dec(ll);
                    } else if (ll == 15u) {
                        uint32_t b;
                        do {
                            b = src[p++]; ll += b;
                            if (b == 255u) { const uint32_t sk = skip_ff_blocks(src, d.plen, p); p += sk; ll += 255u * sk; }
                        } while (b == 255u);
                    }
                    lit = p;
                    oL = (uint32_t)o;
                    if ((uint64_t)o + ll > d.dcap || (uint64_t)(on - o) < ll) bad = true;
                    else if (kind == kEndFinal) { lit_ok = true; a.total[f] = oL + ll; }
                    else {
                        ml = (uint64_t)(on - o) - ll;
                        off = (uint32_t)src[lit + ll] | ((uint32_t)src[lit + ll + 1] << 8);
                        if (off == 0 || (uint64_t)off > (uint64_t)o + ll || (uint64_t)o + ll + ml > d.dcap) bad = true;
                        else { lit_ok = true; match_ok = true; ovl = (uint64_t)off < ml; }
                    }
                }
                if (bad) { a.state[f] = 2; ll = 0; lit_ok = match_ok = false; }
                if (lit_ok) {
                    if (ll < kJumpLong) lls = ll;
                    else if (ll) {
                        const uint32_t q = atomicAdd(a.nlong, 1u);
                        if (q < a.long_cap) { JumpLong e; e.frame = f; e.kind = 0; e.pos = oL; e.len = ll; e.src = lit; e.pad = 0; a.longq[q] = e; }
                        else a.state[f] = 2;
                    }
                }
                if (match_ok) {
                    if (ml < kJumpLong) mls = (uint32_t)ml;
                    else {
                        const uint32_t q = atomicAdd(a.nlong, 1u);
                        if (q < a.long_cap) { JumpLong e; e.frame = f; e.kind = 1; e.pos = oL + ll; e.len = (uint32_t)ml; e.src = off; e.pad = 0; a.longq[q] = e; }
                        else a.state[f] = 2;
                    }
                }
            }
            // exclusive prefix sum of lls + mls over the CTA
            const uint32_t mine = lls + mls;
            uint32_t incl = mine;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, dd);
                if ((int)lane >= dd) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kJumpThreads / 32; w++) {
                const uint32_t v = s_wsum[w];
                if ((uint32_t)w < warp) wbase += v;
                total += v;
            }
            s_cs[tid] = wbase + incl - mine;
            if (tid == 0) s_cs[kJumpThreads] = total;
            s_oL[tid] = oL; s_ll[tid] = ll; s_lls[tid] = lls; s_lit[tid] = lit; s_off[tid] = off | (ovl ? 0x80000000u : 0u);
            __syncthreads();
            for (uint32_t c = tid; c < total; c += kJumpThreads) {
                uint32_t ra = 0, rb = kJumpThreads;           // last record whose first byte is <= c
                while (rb - ra > 1) {
                    const uint32_t mid = (ra + rb) >> 1;
                    if (s_cs[mid] <= c) ra = mid; else rb = mid;
                }
                const uint32_t q = c - s_cs[ra];
                const uint32_t nl = s_lls[ra];
                if (q < nl) {
                    const uint32_t p = s_oL[ra] + q;
                    out[p] = src[s_lit[ra] + q];
                    a.S[X + p] = X + p;
                } else {
                    const uint32_t m = q - nl, oM = s_oL[ra] + s_ll[ra], of = s_off[ra];
                    a.S[X + oM + m] = X + jump_match_src(oM, of & 0xFFFFu, (of >> 31) != 0, m);
                }
            }
            __syncthreads();
        }
    }
}

// ---- long literal runs and matches: the whole grid, 16 KiB slices ---------------------------------------------------
__global__ void __launch_bounds__(kJumpThreads) lz4_jump_long_kernel(JumpArgs a) {
    uint32_t n = *a.nlong;
    if (n > a.long_cap) n = a.long_cap;
    const uint32_t G = gridDim.x;
    for (uint32_t e = 0; e < n; e++) {
        const JumpLong it = a.longq[e];
        if (a.state[it.frame] == 0) continue;
        const FrameDec d = a.fd[it.frame];
        const uint32_t X = (uint32_t)a.dst_off[it.frame];
        uint8_t *out = (d.mode ? a.scratch : a.dst) + a.dst_off[it.frame];
        const uint8_t *src = a.frames + a.frame_off[it.frame] + 16;
        const uint32_t nsl = (it.len + kJumpSlice - 1) / kJumpSlice;
        for (uint32_t s = (blockIdx.x + G - e % G) % G; s < nsl; s += G) {
            const uint32_t m0 = s * kJumpSlice;
            const uint32_t cnt = it.len - m0 < kJumpSlice ? it.len - m0 : kJumpSlice;
            if (it.kind == 0) {
                cta_copy(out + it.pos + m0, src + it.src + m0, cnt);
                for (uint32_t i = threadIdx.x; i < cnt; i += kJumpThreads) a.S[X + it.pos + m0 + i] = X + it.pos + m0 + i;
                // blocks that lie wholly inside this slice hold literals only: nothing to resolve, nothing to gather (the
                // incompressible byte planes of a shuffled frame are half of it)
                const uint32_t b_lo = (it.pos + m0 + kJumpBlock - 1) / kJumpBlock, b_hi = (it.pos + m0 + cnt) / kJumpBlock;
                for (uint32_t b = b_lo + threadIdx.x; b < b_hi; b += kJumpThreads) a.blockdone[(uint64_t)it.frame * a.blocks_per_frame + b] = 2;
            } else {
                const bool ovl = it.src < it.len;
                for (uint32_t i = threadIdx.x; i < cnt; i += kJumpThreads)
                    a.S[X + it.pos + m0 + i] = X + jump_match_src(it.pos, it.src, ovl, m0 + i);
            }
        }
    }
}

// ---- check: a frame whose chain did not end in a closing token is the tile engine's; the others get their status -----
__global__ void lz4_jump_check_kernel(JumpArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes || a.state[f] != 1) return;
    const uint32_t total = a.total[f];
    if (total == 0xFFFFFFFFu) { a.state[f] = 2; return; }
    const FrameDec d = a.fd[f];
    const uint32_t st = total == d.norig ? (uint32_t)kOk : (uint32_t)kESizeMismatch;      // blosc.go:429-431
    FrameMeta m; m.mode = 0; m.typesize = 0;
    if (st == kOk) { m.mode = d.mode; m.typesize = d.typesize; }
    a.status[f] = st; a.out_len[f] = total; a.meta[f] = m;
}

// ---- one round of pointer jumping -------------------------------------------------------------------------------
// In place and unordered: whatever a thread reads in S[j] is an ancestor of j (entries only ever move towards the
// literal they end in), so a stale or a fresh value are both right and the rounds need no double buffer.
__global__ void __launch_bounds__(kJumpThreads) lz4_jump_round_kernel(JumpArgs a, uint32_t round) {
    if (round > 0 && a.changed[round - 1] == 0) return;
    const uint32_t f = blockIdx.x / a.blocks_grid, b0 = blockIdx.x % a.blocks_grid;
    if (a.state[f] != 1) return;
    const uint32_t total = a.total[f];
    const uint32_t X = (uint32_t)a.dst_off[f];
    uint8_t *done = a.blockdone + (uint64_t)f * a.blocks_per_frame;
    const uint32_t nb = (total + kJumpBlock - 1) / kJumpBlock;
    for (uint32_t b = b0; b < nb; b += a.blocks_grid) {
        if (done[b]) continue;
        const uint32_t base = b * kJumpBlock;
        const uint32_t n = total - base < kJumpBlock ? total - base : kJumpBlock;
        uint32_t *Sb = a.S + X + base;
        bool open = false;
        for (uint32_t i0 = threadIdx.x; i0 < n; i0 += 4 * kJumpThreads) {
            uint32_t j[4], kk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * kJumpThreads;
                j[u] = i < n ? Sb[i] : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * kJumpThreads;
                kk[u] = j[u];
                if (i < n && j[u] != X + base + i) kk[u] = __ldcg(a.S + j[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * kJumpThreads;
                if (i < n && kk[u] != j[u]) { Sb[i] = kk[u]; open = true; }
            }
        }
        const int any = __syncthreads_or(open ? 1 : 0);
        if (threadIdx.x == 0) {
            if (any) a.changed[round] = 1; else done[b] = 1;
        }
    }
}

// ---- gather: every match byte takes its literal ---------------------------------------------------------------------
__global__ void __launch_bounds__(kJumpThreads) lz4_jump_gather_kernel(JumpArgs a) {
    const uint32_t f = blockIdx.x / a.blocks_grid, b0 = blockIdx.x % a.blocks_grid;
    if (a.state[f] != 1) return;
    const uint32_t total = a.total[f];
    const uint32_t X = (uint32_t)a.dst_off[f];
    uint8_t *out = (a.fd[f].mode ? a.scratch : a.dst) + a.dst_off[f];
    const uint32_t nb = (total + kJumpBlock - 1) / kJumpBlock;
    const uint8_t *done = a.blockdone + (uint64_t)f * a.blocks_per_frame;
    for (uint32_t b = b0; b < nb; b += a.blocks_grid) {
        if (done[b] == 2) continue;                          // literals only (lz4_jump_long_kernel)
        const uint32_t base = b * kJumpBlock;
        const uint32_t n = total - base < kJumpBlock ? total - base : kJumpBlock;
        const uint32_t *Sb = a.S + X + base;
        for (uint32_t i0 = threadIdx.x; i0 < n; i0 += 4 * kJumpThreads) {
            uint32_t j[4];
            uint8_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * kJumpThreads;
                j[u] = i < n ? Sb[i] : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * kJumpThreads;
                v[u] = 0;
                if (i < n && j[u] != X + base + i) v[u] = out[j[u] - X];
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * kJumpThreads;
                if (i < n && j[u] != X + base + i) out[base + i] = v[u];
            }
        }
    }
}

}  // namespace b2b
