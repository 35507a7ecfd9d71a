// lz4_decode2.cuh -- K4, second design: chunk-parallel token parse, per-frame stitch, tile copy engine.
//
// Replaces lz4Codec.Decompress (codec.go:77-84 -> pierrec UncompressBlock) and the frame checks of
// decompressBackend (blosc.go:377-434) for batches of frames.  The wire format is the reference's: one
// raw LZ4 block per frame, so a frame is ONE serial token chain.  The first design (lz4_kernels.cuh)
// walked that chain with one warp per frame; here the work is cut so that the unit of SIMD work is a
// thread, not a warp:
//
//   prep     one thread per frame: header checks in the reference's order, FrameDec, payload length
//   K5       exclusive scan of ceil(payload / kChunkBytes): chunk slots of every frame
//   parse    one THREAD per 8 KiB chunk of the stream.  It starts kChunkWarm bytes before its chunk,
//            walks tokens without recording until it enters the chunk (LZ4 chains started at a wrong
//            byte usually merge with the true chain within a few tokens), then writes one 8-byte
//            record (token position, output position relative to its start) per token that starts in
//            the chunk, plus where it entered and where it left.
//   stitch   one thread per frame walks the chunks in order with the TRUE position: a chunk whose
//            speculative entry equals the true one is adopted as it is; otherwise the thread walks from
//            the true position until it meets the speculative chain (its records are sorted by
//            position) and writes the few missing records in front of it -- or, when the chains never
//            meet, re-parses the chunk.  Worst case (every chunk wrong) = one thread per frame walking
//            the whole stream, which is what a scalar decoder does.  Output: one descriptor per chunk
//            (record range, absolute output base), i.e. every record now knows its input AND output
//            position: the copy stage is position independent.
//   copy     one CTA (4 warps) per frame, one THREAD per sequence.  Output is staged in shared memory
//            (two 4 KiB tiles used as a ring: the tile being filled and the one before it as history)
//            and leaves as 16-byte vectors.  Per group of 128 records: every thread decodes its token
//            and copies its literals (runs over 64 bytes: the whole CTA, word-wise); then matches run
//            in WAVES: a match whose source bytes are all written is copied by its thread (over 32
//            bytes: by its warp), the others wait for the next wave.  "Written" is tracked exactly, one
//            bit per pending match byte of the tile, so only true dependencies serialise.
//
// Error rules are the first design's (warp_decode_one): every token is checked in stream order --
// literal length beyond the stream (-1), beyond the capacity (-2), closing token with a match nibble
// (-1), offset missing / zero / beyond the output so far (-1), match beyond the capacity (-2) -- and
// the first failing record decides the status.
#pragma once
#include "common.cuh"
#include "lz4_kernels.cuh"

#ifndef B2B_STAT
#define B2B_STAT(i, v) ((void)0)
#endif
#ifdef B2B_EMU
#define B2B_NANOSLEEP(ns) ((void)0)
#else
#define B2B_NANOSLEEP(ns) __nanosleep(ns)
#endif

namespace b2b {

constexpr uint32_t kChunkBytes = 8192;                  // stream bytes per parse chunk
constexpr uint32_t kChunkWarm = 1024;                   // bytes walked before the chunk to find the chain
constexpr uint32_t kChunkHead = 64;                     // free records in front of a chunk's records (stitch prefix)
constexpr uint32_t kChunkCap = kChunkBytes / 3 + 3;     // a token that opens a match takes >= 3 bytes
constexpr uint32_t kChunkSlot = kChunkHead + kChunkCap + 1;   // records per chunk slot (+ the closing record)
// The chunk size is a launch parameter (chunk_shift in the argument structs, 13 = the 8 KiB above unless the host says
// otherwise): a handful of frames of a few MiB are latency, not throughput, and there the host asks for 1 KiB chunks --
// eight times as many threads, each with an eighth of the serial walk (b2b.cu).
constexpr uint32_t kChunkShift = 13, kChunkShiftSmall = 10;
__host__ __device__ __forceinline__ uint32_t chunk_slot_records(uint32_t shift) { return kChunkHead + ((1u << shift) / 3u + 3u) + 1u; }

// how a chain of tokens ended
enum : uint32_t {
    kEndCont = 0,        // ran into the next chunk
    kEndFinal = 1,       // closing token (literals only, ends the stream)
    kEndBadA = 2,        // malformed before its literals (length bytes / literals beyond the stream, or no token at all)
    kEndBadB = 3,        // malformed after its literals (offset missing, closing token with a match nibble, length bytes beyond the stream)
    kEndOverrun = 4,     // match longer than any capacity (after its literals and offset)
    kEndOverrunLit = 5,  // literals beyond any capacity
    kEndDead = 6         // the speculative chain never reached the chunk
};

struct FrameDec {
    uint32_t kind;       // 0: nothing to decode (status final), 1: stored (memcpy flag), 2: LZ4 block, 3: LZ4 block left to the
                         // one-warp-per-frame kernel (small against its output)
    uint32_t plen;       // payload bytes
    uint32_t norig;      // NBytesOrig
    uint32_t dcap;       // min(capacity, NBytesOrig)
    uint32_t mode;       // filter still to run (0 none, 1 byte shuffle, 2 bit shuffle)
    uint32_t typesize;
};

struct ChunkMeta {       // written by the parse kernel
    uint32_t entry;      // first token position >= chunk start on the speculative chain (~0: none)
    uint32_t exit;       // position the chain left the chunk at (or of the token that ended it)
    uint32_t count;      // records (without the closing one)
    uint32_t end;        // kEnd*
    uint32_t out;        // output bytes of those records
    // written by the repair kernel (zero after the parse): the chain that enters at pad[0] -- the exit of the chunk before --
    // meets the speculative one at its record j after mc tokens of its own, which now stand at the end of the head-room;
    // pad[1] = 0x80000000 | mc << 12 | j, pad[2] = output bytes of the chunk on that chain
    uint32_t pad[3];
};

struct ChunkDesc {       // written by the stitch kernel
    long long base_a;    // absolute output position of relative position 0, records [0, split)
    long long base_b;    // the same for records [split, count]
    uint32_t start;      // first record, relative to the chunk slot
    uint32_t count;      // 0: no token of the true chain starts in this chunk
    uint32_t split;
    uint32_t end;
};

// ---- one token, one thread ----------------------------------------------------------------------------
// kEndCont: a sequence (ll literals at lit, ml match bytes, next token at next); kEndFinal: closing token;
// kEndBadA / kEndBadB: malformed (B: ll / lit are valid and inside the stream).  Written without early
// returns: the callers keep the lanes of a warp converged across tokens (one vote per token), and only the
// rare length-extension loops diverge.
// Length bytes of a very long run (a 100 MB literal run has 400 000 of them): whole aligned 16-byte blocks of 0xFF
// are taken at once.  Returns how many bytes were skipped (a multiple of 16); each adds 255.
// A run of R literals has R / 255 length bytes (the incompressible byte plane of ONE 1 GiB frame: a million of them) and
// one thread walks them, so the walk must not pay a memory round trip per 16 bytes (one 256 MiB frame: chunk parse 8.0 ->
// 4.1 ms).  Not inlined: its registers are only paid for where a run is that long.
__device__ __noinline__ uint32_t skip_ff_blocks(const uint8_t *__restrict__ s, uint32_t clen, uint32_t p) {
    uint32_t n = 0;
    if (((uintptr_t)(s + p) & 15u) == 0) {
        while (p + n + 512u <= clen) {                        // four cache lines per step: as many loads in flight as registers allow
            const uint4 *q = reinterpret_cast<const uint4 *>(s + p + n);
            uint32_t acc = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < 32; i++) { const uint4 v = q[i]; acc &= v.x & v.y & v.z & v.w; }
            if (acc != 0xFFFFFFFFu) break;
            n += 512u;
        }
        while (p + n + 128u <= clen) {                        // a cache line per step, its eight loads in flight together
            const uint4 *q = reinterpret_cast<const uint4 *>(s + p + n);
            const uint4 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3], v4 = q[4], v5 = q[5], v6 = q[6], v7 = q[7];
            const uint32_t lo = (v0.x & v0.y & v0.z & v0.w) & (v1.x & v1.y & v1.z & v1.w) & (v2.x & v2.y & v2.z & v2.w) &
                                (v3.x & v3.y & v3.z & v3.w);
            const uint32_t hi = (v4.x & v4.y & v4.z & v4.w) & (v5.x & v5.y & v5.z & v5.w) & (v6.x & v6.y & v6.z & v6.w) &
                                (v7.x & v7.y & v7.z & v7.w);
            if ((lo & hi) != 0xFFFFFFFFu) break;
            n += 128u;
        }
        while (p + n + 16u <= clen) {
            const uint4 v = *reinterpret_cast<const uint4 *>(s + p + n);
            if ((v.x & v.y & v.z & v.w) != 0xFFFFFFFFu) break;
            n += 16u;
        }
    }
    return n;
}

__device__ __forceinline__ uint32_t tok_step(const uint8_t *__restrict__ s, uint32_t clen, uint32_t p, uint32_t &ll,
                                             uint32_t &lit, uint64_t &ml, uint32_t &next) {
    ll = 0; lit = p; ml = 0; next = p;
    uint32_t kind = kEndCont;
    if (p >= clen) {
        kind = kEndBadA;
    } else {
        const uint32_t tok = s[p++];
        uint64_t l = tok >> 4;
        if (l == 15) {
            bool more = true;
            while (more) {
                if (p >= clen) { kind = kEndBadA; more = false; }
                else {
                    const uint32_t b = s[p++]; l += b; more = b == 255u;
                    if (more) { const uint32_t k = skip_ff_blocks(s, clen, p); p += k; l += 255ull * k; }
                }
            }
        }
        if (kind == kEndCont && l > (uint64_t)(clen - p)) kind = kEndBadA;
        if (kind == kEndCont) {
            ll = (uint32_t)l; lit = p;
            p += ll;
            uint64_t m = tok & 15u;
            if (p == clen) kind = m != 0 ? kEndBadB : kEndFinal;
            else if (clen - p < 2) kind = kEndBadB;
            else {
                p += 2;
                if (m == 15) {
                    bool more = true;
                    while (more) {
                        if (p >= clen) { kind = kEndBadB; more = false; }
                        else {
                            const uint32_t b = s[p++]; m += b; more = b == 255u;
                            if (more) { const uint32_t k = skip_ff_blocks(s, clen, p); p += k; m += 255ull * k; }
                        }
                    }
                    if (m > 0xFFFFFFFFull) kind = kEndBadB;
                }
                if (kind == kEndCont) { ml = m + 4; next = p; }
            }
        }
    }
    return kind;
}

struct Walk { uint32_t pos, n, end; uint64_t rel; };

// Eight stream bytes from position p as one little-endian word: two ALIGNED 8-byte loads and a funnel shift.  The lanes of
// a warp walk 32 different chunks, so every byte load is 32 cache-line lookups; a token, its length bytes and the next
// field's length bytes taken byte by byte were ~17 such loads per turn, and the turn time was the LSU working through
// them.  Reads the two aligned words that cover [p, p + 8): up to 7 bytes before p and 15 behind it.
__device__ __forceinline__ uint64_t walk_window8(const uint8_t *__restrict__ s, uint32_t p) {
    const uintptr_t a = (uintptr_t)(s + p);
    const uint64_t *q = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    const uint32_t sh = 8u * (uint32_t)(a & 7u);
    const uint64_t lo = q[0], hi = q[1];
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
}
// A length field whose nibble was 15: its length bytes are bytes first .. 7 of the window w (255, ..., 255, last).  v grows
// by their sum; returns the bytes taken, 0 when the field does not end inside the window (the caller takes the general
// path).  The first byte that is not 0xFF is found with one find-first-set on the inverted word: a walk is a serial
// instruction stream, and a loop over the bytes was most of its instructions.
__device__ __forceinline__ uint32_t walk_len_ext(uint64_t w, uint32_t first, uint32_t &v) {
    const uint64_t t = (~w) >> (8u * first);              // byte i: zero iff length byte i is 0xFF (zeros come in at the top)
    if (t == 0) return 0u;
    const uint32_t idx = (uint32_t)(__ffsll((long long)t) - 1) >> 3;
    v += 255u * idx + ((uint32_t)(w >> (8u * (first + idx))) & 0xFFu);
    return idx + 1u;
}

// one step of a walk: the token at w.pos; record (kEmit) at out[w.n].  Returns false when the chain ended
// (w.end says how).  rel counts output bytes from the start of the walk and never passes 2^32 - 1.
//
// Fast path: a token whose two lengths take at most seven / eight length bytes each (lengths up to ~2 000) and whose
// sequence ends at least 24 bytes before the end of the stream -- practically every token -- is stepped over with two
// round trips to memory: an 8-byte window with the token and its length bytes, then one with the match's length bytes.  A walk is ONE dependent chain (a thread of
// the parse kernel among 31 others in lockstep, a warp of the stitch kernel) and the general tok_step pays a round trip per
// length byte; with 32 lanes per warp some lane had a match of several hundred bytes nearly every turn (the sign /
// exponent plane of a smooth field: 1.7 us per token).  Anything else -- longer lengths, the last sequences of a stream,
// anything malformed, a position counter near 2^32 -- takes the general path, so the result is the same by construction.
template <bool kEmit>
__device__ __forceinline__ bool walk_step(const uint8_t *__restrict__ s, uint32_t clen, Walk &w, uint2 *out) {
    {
        const uint32_t p = w.pos;
        if (p + 24u <= clen && w.rel < 0xFFFF0000ull) {           // the token's window lies inside the stream
            const uint64_t w0 = walk_window8(s, p);               // token and up to seven length bytes
            const uint32_t tok = (uint32_t)w0 & 0xFFu;
            uint32_t l = tok >> 4, q = p + 1u;
            bool ok = true;
            if (l == 15u) { const uint32_t k = walk_len_ext(w0, 1u, l); ok = k != 0u; q += k; }
            const uint32_t e = q + l;                             // literals [q, e), offset [e, e + 2), length bytes from e + 2
            if (ok && e + 26u <= clen) {
                uint32_t m = tok & 15u, nx = e + 2u;
                if (m == 15u) { const uint32_t k = walk_len_ext(walk_window8(s, nx), 0u, m); ok = k != 0u; nx += k; }
                if (ok) {
                    if (kEmit) out[w.n] = make_uint2(p, (uint32_t)w.rel);
                    w.n++;
                    w.rel += l + m + 4u;                          // <= 2 * (15 + 8 * 255) + 4: far from 2^32 (checked above)
                    w.pos = nx;
                    return true;
                }
            }
        }
    }
    uint32_t ll, lit, next; uint64_t ml;
    const uint32_t kind = tok_step(s, clen, w.pos, ll, lit, ml, next);
    if (kEmit) out[w.n] = make_uint2(w.pos, (uint32_t)w.rel);
    w.n++;
    bool go = false;
    if (kind == kEndBadA) w.end = kEndBadA;
    else if (w.rel + ll > 0xFFFFFFFFull) w.end = kEndOverrunLit;
    else if (kind != kEndCont) { w.rel += ll; w.end = kind; }
    else if (w.rel + ll + ml > 0xFFFFFFFFull) { w.rel += ll; w.end = kEndOverrun; }
    else { w.rel += ll + ml; w.pos = next; go = true; }
    return go;
}

// Walks the chain from pos while pos < cend, at most max_n tokens (single thread: the stitch kernel)
template <bool kEmit>
__device__ __forceinline__ Walk chain_walk(const uint8_t *__restrict__ s, uint32_t clen, uint32_t pos, uint32_t cend,
                                           uint32_t max_n, uint2 *out) {
    Walk w; w.pos = pos; w.n = 0; w.end = kEndCont; w.rel = 0;
    bool go = w.pos < cend && w.n < max_n;
    while (go) go = walk_step<kEmit>(s, clen, w, out) && w.pos < cend && w.n < max_n;
    return w;
}

// ---- prep: header checks (blosc.go:296-303, 165-185, 385-390, 393-407, 417-431) ---------------------------
struct Prep2Args {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const uint32_t *frame_len;
    const uint32_t *dst_cap;
    uint32_t nframes;
    int64_t typesize_override;
    FrameDec *fd;
    uint32_t *plen_eff;     // payload bytes of the frames that hold an LZ4 block, else 0 (input of the chunk scan)
    uint32_t *out_len, *status;
    FrameMeta *meta;
    uint32_t keep_sparse = 0;   // 1: frames that are small against their output stay with this decoder (kind 2), see below
};

__global__ void frame_prep_kernel(Prep2Args a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    const uint8_t *fr = a.frames + a.frame_off[f];
    uint32_t flags = 0, codec = 0, tsz = 0, norig = 0, ncomp = 0;
    uint32_t st = check_header(fr, a.frame_len[f], flags, codec, tsz, norig, ncomp);
    FrameDec d; d.kind = 0; d.plen = 0; d.norig = norig; d.dcap = 0; d.mode = 0; d.typesize = 0;
    if (st == kOk) {
        const bool is_memcpy = (flags & 0x2u) != 0;
        if (!is_memcpy) {
            if (codec < 1 || codec > 5) st = kEInvalidCodec;          // blosc.go:403-407
            else if (codec != 1 && codec != 2) st = kEUnsupported;    // Snappy/ZLIB/ZSTD: host side
        }
        if (st == kOk) {
            d.plen = ncomp - 16;
            const uint64_t T = a.typesize_override > 0 ? (uint64_t)a.typesize_override : (uint64_t)tsz;
            const uint32_t mode = (flags & 0x4u) ? 2u : ((flags & 0x1u) ? 1u : 0u);
            const bool active = mode != 0 && T > 1 && (uint64_t)norig >= T;
            d.mode = active ? mode : 0u;
            d.typesize = active ? (uint32_t)T : 0u;
            const uint32_t cap = a.dst_cap[f];
            d.dcap = cap < norig ? cap : norig;
            if (is_memcpy) {
                if (d.plen != norig) st = kESizeMismatch;             // blosc.go:398-400, 429-431
                else if (cap < norig) st = kEDstTooSmall;
                else d.kind = 1;
            } else {
                // a block that decodes to n bytes is at most n + n / 255 + 16 bytes long: a longer one
                // runs over the capacity or is malformed, whichever comes first
                const uint64_t bound = (uint64_t)d.dcap + d.dcap / 255u + 16u;
                if ((uint64_t)d.plen > bound) st = d.dcap == norig ? kEDecompressionFailed : kEDstTooSmall;
                // A frame whose block is small against its output (long runs, sparse arrays: ratio under 0.3) is a few
                // thousand sequences however large it is: one warp walks that faster than a CTA walks the frame's tiles
                // (4 MiB frames of a sparse int32 array: 35 -> 197 GB/s), so it goes to the one-warp-per-frame kernel
                // that runs behind the copy engine (kind 3; the stitch kernel flags it).
                // (not when the pointer-jumping engine of lz4_decode4.cuh follows: it does not care about sequence lengths)
                else d.kind = (!a.keep_sparse && 10ull * d.plen < 3ull * norig) ? 3u : 2u;
            }
        }
    }
    a.fd[f] = d;
    a.plen_eff[f] = d.kind == 2 ? d.plen : 0u;
    a.status[f] = st;
    a.out_len[f] = 0;
    FrameMeta m; m.mode = 0; m.typesize = 0;
    a.meta[f] = m;
}

// ---- parse: one thread per chunk ---------------------------------------------------------------------------
struct Parse2Args {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const FrameDec *fd;
    uint32_t nframes;
    const uint64_t *chunk_base;     // exclusive scan of the chunks per frame
    const uint64_t *total_chunks;
    uint2 *table;                   // kChunkSlot records per chunk
    ChunkMeta *meta;
    uint64_t table_chunks;          // chunks the table has room for
    uint32_t *dead;                 // one word per chunk, zero before launch: "inside a long literal run, do not walk"
    unsigned long long *ticket;     // zero before launch: next chunk to hand out
    uint32_t chunk_shift = kChunkShift;
};

constexpr int kParse2Threads = 128;
constexpr int kWalkBurst = 8;           // tokens a parse lane walks between two votes of its warp
#ifndef B2B_PARSE2_CTAS
#define B2B_PARSE2_CTAS 8       // 64 registers (at 12 CTAs / 40 registers the walk spilled 528 bytes into its token loop)
#endif

// Persistent lanes: every lane walks one chunk at a time, one token per turn, in lockstep with the other lanes of
// its warp (a vote per turn keeps them converged whatever the tokens are); a lane that is done takes the next
// chunk from a ticket as soon as a quarter of the warp is idle.  A lane whose token runs over whole following
// chunks (a long literal run: the incompressible byte planes of a shuffled frame) flags those chunks, and whoever
// holds them stops walking what can only be the inside of that run.  The flag may come from a speculative chain
// that is itself wrong; the stitch kernel re-parses such a chunk, so it costs time, never correctness.
__global__ void __launch_bounds__(kParse2Threads, B2B_PARSE2_CTAS) lz4_chunk_parse_kernel(Parse2Args a) {
    uint64_t total = *a.total_chunks;
    if (total > a.table_chunks) total = a.table_chunks;
    const uint32_t CSH = a.chunk_shift, CB = 1u << CSH, CS = chunk_slot_records(CSH);
    const int lane = (int)(threadIdx.x & 31u);
    ChunkMeta m; m.pad[0] = m.pad[1] = m.pad[2] = 0;
    m.entry = 0xFFFFFFFFu; m.exit = 0; m.count = 0; m.end = kEndDead; m.out = 0;
    const uint8_t *s = nullptr;
    uint32_t plen = 0, cbeg = 0, cend = 0, k = 0, nch = 0;
    uint64_t g = 0;
    uint2 *rec = nullptr;
    int phase = 2;                                            // 0 warm-up, 1 recording, 2 idle, 3 no chunks left
    Walk w; w.pos = 0; w.n = 0; w.end = kEndCont; w.rel = 0;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, phase == 2);
        const uint32_t busy = __ballot_sync(0xffffffffu, phase == 0 || phase == 1);
        if (idle == 0 && busy == 0) break;
        if (idle && (__popc(idle) >= 8 || busy == 0)) {
            unsigned long long base = 0;
            if (lane == __ffs(idle) - 1) base = atomicAdd(a.ticket, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, __ffs(idle) - 1);
            if (phase == 2) {
                g = base + __popc(idle & ((1u << lane) - 1u));
                if (g >= total) phase = 3;
                else {
                    uint32_t lo = 0, hi = a.nframes;          // last frame whose first chunk is <= g
                    while (hi - lo > 1) {
                        const uint32_t mid = lo + (hi - lo) / 2;
                        if (a.chunk_base[mid] <= g) lo = mid; else hi = mid;
                    }
                    const uint32_t f = lo;
                    k = (uint32_t)(g - a.chunk_base[f]);
                    plen = a.fd[f].plen;
                    nch = (uint32_t)(((uint64_t)plen + CB - 1) >> CSH);
                    m.entry = 0xFFFFFFFFu; m.exit = 0; m.count = 0; m.end = kEndDead; m.out = 0;
                    w.n = 0; w.end = kEndCont; w.rel = 0;
                    if (a.fd[f].kind == 2 && k < nch && !__ldcg(a.dead + g)) {
                        s = a.frames + a.frame_off[f] + 16;
                        cbeg = k << CSH;
                        cend = k + 1 == nch ? plen + 1 : cbeg + CB;   // the last chunk owns position plen
                        w.pos = cbeg > kChunkWarm ? cbeg - kChunkWarm : 0u;
                        rec = a.table + g * CS + kChunkHead;
                        phase = w.pos < cbeg ? 0 : 1;
                        if (phase == 1) m.entry = w.pos;
                    } else {
                        a.meta[g] = m;                        // nothing to walk
                    }
                }
            }
        }
        // Up to kWalkBurst tokens per turn: the vote, the ticket logic and the phase dispatch around a token cost as many
        // instructions as the token itself, and a walk is as fast as its instruction stream is short (ncu on a frame of the
        // sign / exponent plane alone: 130 instructions per token at 12 cycles each).  The lanes still meet every turn.
        if (phase == 0) {
            bool alive = true;
#pragma unroll 1
            for (int r = 0; r < kWalkBurst; r++) {
                alive = walk_step<false>(s, plen, w, nullptr);
                if (!alive || w.pos >= cbeg) break;
            }
            if (!alive) { phase = 2; a.meta[g] = m; }   // the speculative chain died before the chunk
            else if (w.pos >= cbeg) {
                // (a token that jumps over whole chunks from the warm-up zone says nothing certain: no flags)
                m.entry = w.pos; m.exit = w.pos; m.end = kEndCont;
                w.n = 0; w.rel = 0;
                if (w.pos < cend && !__ldcg(a.dead + g)) phase = 1;
                else { if (w.pos < cend) { m.entry = 0xFFFFFFFFu; m.end = kEndDead; } phase = 2; a.meta[g] = m; }
            }
        } else if (phase == 1) {
            bool go = true;
#pragma unroll 1
            for (int r = 0; r < kWalkBurst; r++) {
                go = walk_step<true>(s, plen, w, rec);
                if (!go || w.pos >= cend || (w.n & 31u) == 0u) break;     // (w.pos >= cbeg + 2 * CB implies w.pos >= cend)
            }
            if (go && w.pos >= cbeg + 2 * CB) {
                // this token covers the chunks up to the one that holds w.pos: nothing starts inside them
                const uint32_t k1 = (w.pos >> CSH) < nch ? (w.pos >> CSH) : nch;
                for (uint32_t j = k + 1; j < k1; j++) a.dead[g + (j - k)] = 1u;
            }
            if (!go || w.pos >= cend) {
                rec[w.n] = make_uint2(w.pos, (uint32_t)w.rel);
                m.exit = w.pos; m.count = w.n; m.end = w.end; m.out = (uint32_t)w.rel;
                phase = 2; a.meta[g] = m;
            } else if ((w.n & 31u) == 0u && __ldcg(a.dead + g)) {
                m.entry = 0xFFFFFFFFu; m.count = 0; m.end = kEndDead; m.out = 0;
                phase = 2; a.meta[g] = m;
            }
        }
    }
}

// ---- repair: one thread per chunk ---------------------------------------------------------------------------------
// A chunk whose speculative chain did not start where the chain of the chunk before it ends is repaired by walking
// from that exit until the two chains meet (the stitch kernel's own rule).  The stitch kernel does this where the TRUE
// chain arrives, one chunk after the other; here every such chunk is repaired at once, on the assumption that the
// chunk before it ends where its speculative chain ends -- true whenever that chain was right or was itself repaired
// by merging, which is the case the stitch kernel then only has to confirm (entry == pad[0]); where the assumption
// was wrong the stitch kernel ignores the result and repairs as before.  The repair's own tokens go into the head-room
// in front of the chunk's records and nowhere else: the speculative records must stay whole, they are what the stitch
// kernel adopts or searches when the assumption does not hold.
struct Repair2Args {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const FrameDec *fd;
    uint32_t nframes;
    const uint64_t *chunk_base;
    const uint64_t *total_chunks;
    uint2 *table;
    ChunkMeta *meta;
    uint64_t table_chunks;
    uint32_t chunk_shift = kChunkShift;
};

__global__ void __launch_bounds__(128) lz4_chunk_repair_kernel(Repair2Args a) {
    uint64_t total = *a.total_chunks;
    if (total > a.table_chunks) total = a.table_chunks;
    const uint32_t CSH = a.chunk_shift, CB = 1u << CSH, CS = chunk_slot_records(CSH);
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0 || g >= total) return;
    uint32_t lo = 0, hi = a.nframes;                          // last frame whose first chunk is <= g
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (a.chunk_base[mid] <= g) lo = mid; else hi = mid;
    }
    const uint32_t f = lo;
    const FrameDec d = a.fd[f];
    const uint32_t k = (uint32_t)(g - a.chunk_base[f]);
    const uint32_t plen = d.plen, nch = (uint32_t)(((uint64_t)plen + CB - 1) >> CSH);
    if (d.kind != 2 || k == 0 || k >= nch) return;
    const ChunkMeta prev = a.meta[g - 1], m = a.meta[g];
    if (prev.entry == 0xFFFFFFFFu || prev.end != kEndCont) return;           // no chain runs out of the chunk before
    const uint32_t e = prev.exit;
    if (((e >> CSH) < nch ? (e >> CSH) : nch - 1) != k || m.entry == e) return;
    const uint32_t sc = m.entry == 0xFFFFFFFFu ? 0u : m.count;
    if (sc == 0) return;                                                      // nothing to meet: the stitch kernel re-parses
    const uint8_t *s = a.frames + a.frame_off[f] + 16;
    uint2 *slot = a.table + g * CS;
    const uint32_t cend = k + 1 == nch ? plen + 1 : (k + 1) << CSH;
    uint32_t pos = e, j = 0, mc = 0, spec_tok = slot[kChunkHead].x;
    uint64_t rel = 0;
    for (;;) {
        if (pos >= cend) return;
        if (spec_tok < pos) {
            if (++j >= sc) return;
            spec_tok = slot[kChunkHead + j].x;
        } else if (spec_tok == pos) {
            if (mc > kChunkHead) return;                                      // no room in the head-room
            break;
        } else {
            Walk t; t.pos = pos; t.n = 0; t.end = kEndCont; t.rel = 0;
            const bool go = walk_step<false>(s, plen, t, nullptr);
            mc++;
            if (!go || rel + t.rel > 0xFFFFFFFFull) return;                   // the chain ends inside the chunk
            rel += t.rel; pos = t.pos;
        }
    }
    const uint64_t out_alt = rel + (uint64_t)(m.out - slot[kChunkHead + j].y);
    if (out_alt > 0xFFFFFFFFull || mc >= 4096u || j >= 4096u) return;
    Walk w; w.pos = e; w.n = 0; w.end = kEndCont; w.rel = 0;
    uint2 *dstrec = slot + kChunkHead - mc;       // the head-room only: the speculative records stay whole (they are the truth if the assumption is not)
    while (w.n < mc && w.pos < cend) { if (!walk_step<true>(s, plen, w, dstrec)) return; }
    ChunkMeta r = m;
    r.pad[0] = e; r.pad[1] = 0x80000000u | (mc << 12) | j; r.pad[2] = (uint32_t)out_alt;
    a.meta[g] = r;
}

// A chunk is adopted on its repaired chain (pad[] of its meta; output so far = op).  Only now is it certain that the
// speculative records in front of the meeting point are dead, so only now do the repair's tokens move from the head-room
// to their place in front of record j: the copy engines read a chunk's records as one run.
__device__ __forceinline__ ChunkDesc adopt_repaired(uint2 *slot, const ChunkMeta &m, long long op, bool move = true) {
    const uint32_t j = m.pad[1] & 0xFFFu, mc = (m.pad[1] >> 12) & 0xFFFu;
    if (move && j != 0) for (uint32_t q = mc; q-- > 0;) slot[kChunkHead + j - mc + q] = slot[kChunkHead - mc + q];   // (moves up: last first)
    ChunkDesc D;
    D.base_a = op; D.base_b = op + ((long long)m.pad[2] - (long long)m.out);
    D.start = kChunkHead + j - mc; D.count = mc + (m.count - j); D.split = mc; D.end = m.end;
    return D;
}

// ---- stitch: one thread per frame ----------------------------------------------------------------------------
struct Stitch2Args {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const FrameDec *fd;
    uint32_t nframes;
    const uint64_t *chunk_base;
    uint2 *table;
    const ChunkMeta *meta;
    ChunkDesc *desc;
    uint32_t *last_chunk;           // last chunk with a descriptor (~0: the frame has no chunk)
    uint32_t *fallback;             // 1: no room in the table, the frame is decoded by the first design's kernel
    uint64_t table_chunks;
    uint32_t chunk_shift = kChunkShift;
    uint2 *scratch = nullptr;       // warp per frame: one spare chunk slot per frame (records of a walk whose place is not known yet)
};

// A state machine per lane, one small step per turn, so that the lanes of a warp stay together whatever their
// frames need: 0 look at the next chunk, 1 walk from the true position towards the speculative chain, 2 write the
// records that walk found missing (or the whole chunk), 3 done.
// kWarpPerFrame: the same walk with a whole WARP per frame (few, large frames: one 1 GiB frame has 10^5 chunks and a
// thread that adopts them one dependent load at a time needs 40 ms).  All lanes run the state machine redundantly (same
// loads, same stores), and the common step -- adopt the next chunk -- is taken 32 chunks at a time: lane i looks at chunk
// k + i, the leading lanes whose entry equals the exit of the chunk before them (and whose chain simply runs into the
// next chunk) are adopted together, output bases from a warp prefix sum; the first chunk that is anything else
// (a wrong entry, a token that jumps over chunks, the end of the chain, the capacity) goes through the scalar steps.
template <bool kWarpPerFrame>
__device__ __forceinline__ void stitch_body(const Stitch2Args &a) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t f = kWarpPerFrame ? gt >> 5 : gt;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t CSH = a.chunk_shift, CB = 1u << CSH, CS = chunk_slot_records(CSH);
    int mode = 3;
    FrameDec d; d.kind = 0; d.plen = 0; d.dcap = 0;
    uint32_t plen = 0, nch = 0;
    uint64_t cb = 0;
    const uint8_t *s = nullptr;
    if (f < a.nframes) {
        a.fallback[f] = 0;
        a.last_chunk[f] = 0xFFFFFFFFu;
        d = a.fd[f];
        plen = d.plen;
        nch = (uint32_t)(((uint64_t)plen + CB - 1) >> CSH);
        if (d.kind == 3) a.fallback[f] = 1;
        if (d.kind == 2 && nch != 0) {
            cb = a.chunk_base[f];
            if (cb + nch > a.table_chunks) a.fallback[f] = 1;
            else { s = a.frames + a.frame_off[f] + 16; mode = 0; }
        }
    }
    uint32_t e = 0, knext = 0, k = 0, cend = 0;
    long long op = 0;
    ChunkMeta m; m.entry = 0; m.exit = 0; m.count = 0; m.end = 0; m.out = 0;
    uint2 *slot = nullptr;
    // mode 1
    uint32_t pos = 0, j = 0, sc = 0, spec_tok = 0, mc = 0;
    uint64_t rel = 0;
    // mode 2
    Walk w; w.pos = 0; w.n = 0; w.end = kEndCont; w.rel = 0;
    uint2 *dstrec = nullptr;
    uint32_t limit = 0;
    bool full = false, prev_full = false;
    while (__any_sync(0xffffffffu, mode != 3)) {
        ChunkDesc D; D.count = 0xFFFFFFFFu;                   // set: a descriptor is ready this turn
        uint32_t endk = kEndCont;
        if (mode == 0) {
            k = e >> CSH;
            if (k >= nch) k = nch - 1;
            if (kWarpPerFrame) { for (uint32_t q = knext + lane; q < k; q += 32) a.desc[cb + q].count = 0; }
            else { for (uint32_t q = knext; q < k; q++) a.desc[cb + q].count = 0; }   // chunks inside one long token
            knext = k + 1;
            bool took = false;
            if (kWarpPerFrame) {
                const uint32_t ki = k + lane;
                ChunkMeta mi; mi.entry = 0xFFFFFFFFu; mi.exit = 0; mi.count = 0; mi.end = kEndDead; mi.out = 0;
                if (ki < nch) mi = a.meta[cb + ki];
                uint32_t prev_exit = __shfl_up_sync(0xffffffffu, mi.exit, 1);
                if (lane == 0) prev_exit = e;
                const bool alt = mi.entry != prev_exit && (mi.pad[1] >> 31) != 0 && mi.pad[0] == prev_exit;   // repaired chain
                const uint32_t my_out = alt ? mi.pad[2] : mi.out;
                unsigned long long incl = my_out;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, dd);
                    if ((int)lane >= dd) incl += t;
                }
                const uint32_t nk = (mi.exit >> CSH) < nch ? (mi.exit >> CSH) : nch - 1;
                const bool reg = ki + 1 < nch && (mi.entry == prev_exit || alt) && mi.end == kEndCont && nk == ki + 1 &&
                                 op + (long long)incl <= (long long)d.dcap;
                const uint32_t irr = ~__ballot_sync(0xffffffffu, reg);
                const uint32_t r = irr ? (uint32_t)__ffs((int)irr) - 1u : 32u;
                if (lane == 0) { B2B_STAT(22, 1); B2B_STAT(23, r); }
                if (r > 0) {
                    if (lane < r) {
                        ChunkDesc Di;
                        Di.base_a = 0; Di.base_b = op + (long long)(incl - my_out); Di.start = kChunkHead; Di.count = mi.count;
                        Di.split = 0; Di.end = kEndCont;
                        if (alt) Di = adopt_repaired(a.table + (cb + ki) * CS, mi, op + (long long)(incl - my_out));
                        a.desc[cb + ki] = Di;
                    }
                    op += (long long)__shfl_sync(0xffffffffu, incl, (int)r - 1);
                    e = __shfl_sync(0xffffffffu, mi.exit, (int)r - 1);
                    knext = k + r;
                    took = true; prev_full = false;
                }
            }
            if (took) continue;
            m = a.meta[cb + k];
            slot = a.table + (cb + k) * CS;
            cend = k + 1 == nch ? plen + 1 : (k + 1) << CSH;
            if (m.entry == e) {
                D.base_a = 0; D.base_b = op; D.start = kChunkHead; D.count = m.count; D.split = 0; D.end = m.end;
                op += m.out; e = m.exit; endk = m.end; prev_full = false;
            } else if ((m.pad[1] >> 31) != 0 && m.pad[0] == e) {          // repaired for exactly this entry (repair kernel)
                D = adopt_repaired(slot, m, op, !kWarpPerFrame || lane == 0);   // (the move overlaps itself: one lane does it)
                if (kWarpPerFrame) __syncwarp();
                op += m.pad[2]; e = m.exit; endk = m.end; prev_full = false;
            } else if (kWarpPerFrame && a.scratch) {
                // One walk from the true entry does both jobs: it looks for the speculative chain AND keeps its tokens (in
                // the frame's spare slot), so that neither outcome walks the chunk a second time -- the records move to
                // their place, 32 lanes wide, when the walk knows where that is.  (One 1 GiB frame: 17 chunks of the
                // sign / exponent plane never meet their speculative chain, 2 000 tokens each; search + re-parse was 17 of
                // the kernel's 19 ms.)
                B2B_STAT(20, 1);
                sc = m.entry == 0xFFFFFFFFu ? 0u : m.count;
                j = 0; spec_tok = sc ? slot[kChunkHead].x : 0xFFFFFFFFu;
                uint2 *tmp = a.scratch + (uint64_t)f * CS;
                w.pos = e; w.n = 0; w.end = kEndCont; w.rel = 0;
                bool merged = false;
                for (;;) {
                    if (w.pos >= cend) break;
                    if (j < sc && spec_tok < w.pos) { j++; spec_tok = j < sc ? slot[kChunkHead + j].x : 0xFFFFFFFFu; continue; }
                    if (j < sc && spec_tok == w.pos && w.n <= kChunkHead + j) { merged = true; break; }
                    if (!walk_step<true>(s, plen, w, tmp)) break;          // the chain ends inside the chunk
                }
                __syncwarp();                                              // every lane is done reading the speculative records
                if (merged) {
                    const uint32_t nmine = w.n, spec_rel = slot[kChunkHead + j].y;
                    for (uint32_t q = lane; q < nmine; q += 32) slot[kChunkHead + j - nmine + q] = tmp[q];
                    D.base_a = op; D.base_b = op + (long long)w.rel - (long long)spec_rel;
                    D.start = kChunkHead + j - nmine; D.count = nmine + (m.count - j); D.split = nmine; D.end = m.end;
                    op += (long long)w.rel + (long long)(m.out - spec_rel);
                    e = m.exit; endk = m.end; prev_full = false;
                } else {
                    B2B_STAT(21, 1);
                    for (uint32_t q = lane; q < w.n; q += 32) slot[kChunkHead + q] = tmp[q];
                    slot[kChunkHead + w.n] = make_uint2(w.pos, (uint32_t)w.rel);
                    D.base_a = 0; D.base_b = op; D.start = kChunkHead; D.count = w.n; D.split = 0; D.end = w.end;
                    op += (long long)w.rel; e = w.pos; endk = w.end; prev_full = true;
                }
            } else if (prev_full) {
                // the chunk before this one had to be re-parsed as a whole and this one is wrong again: two chains that
                // run side by side without meeting (sequences of one fixed length, e.g. token + offset + one length byte
                // in the sign / exponent plane of a smooth field: whatever starts off phase stays off phase).  Looking for
                // the meeting point first would walk the chunk twice.
                B2B_STAT(20, 1); B2B_STAT(21, 1);
                full = true;
                w.pos = e; w.n = 0; w.end = kEndCont; w.rel = 0;
                limit = 0xFFFFFFFFu; dstrec = slot + kChunkHead;
                mode = 2;
            } else {
                B2B_STAT(20, 1);
                sc = m.entry == 0xFFFFFFFFu ? 0u : m.count;
                j = 0; spec_tok = sc ? slot[kChunkHead].x : 0xFFFFFFFFu;
                pos = e; mc = 0; rel = 0;
                mode = 1;
            }
        } else if (mode == 1) {
            bool to_full = false, to_prefix = false;
            if (pos >= cend) to_full = true;
            else if (j < sc && spec_tok < pos) { j++; spec_tok = j < sc ? slot[kChunkHead + j].x : 0xFFFFFFFFu; }
            else if (j < sc && spec_tok == pos && mc <= kChunkHead + j) to_prefix = true;
            else {
                Walk t; t.pos = pos; t.n = 0; t.end = kEndCont; t.rel = 0;
                const bool go = walk_step<false>(s, plen, t, nullptr);
                mc++;
                if (!go || rel + t.rel > 0xFFFFFFFFull) to_full = true;    // the chain ends inside the chunk
                else { rel += t.rel; pos = t.pos; }
            }
            if (to_full || to_prefix) {
                full = to_full;
                if (full) B2B_STAT(21, 1);
                w.pos = e; w.n = 0; w.end = kEndCont; w.rel = 0;
                limit = full ? 0xFFFFFFFFu : mc;
                dstrec = full ? slot + kChunkHead : slot + kChunkHead + j - mc;
                mode = 2;
            }
        } else if (mode == 2) {
            bool fin = !(w.pos < cend && w.n < limit);
            if (!fin) fin = !walk_step<true>(s, plen, w, dstrec);
            // (a warp per frame has nobody to stay in step with: the whole walk in one turn, without the state machine around it)
            if (kWarpPerFrame) while (!fin) { fin = !(w.pos < cend && w.n < limit); if (!fin) fin = !walk_step<true>(s, plen, w, dstrec); }
            if (fin) {
                if (full) {
                    dstrec[w.n] = make_uint2(w.pos, (uint32_t)w.rel);
                    D.base_a = 0; D.base_b = op; D.start = kChunkHead; D.count = w.n; D.split = 0; D.end = w.end;
                    op += (long long)w.rel; e = w.pos; endk = w.end;
                    prev_full = true;
                } else {
                    prev_full = false;
                    // mc records in front of spec[j]; they count from the true position (base_a)
                    const uint32_t spec_rel = slot[kChunkHead + j].y;
                    D.base_a = op; D.base_b = op + (long long)rel - (long long)spec_rel;
                    D.start = kChunkHead + j - mc; D.count = mc + (m.count - j); D.split = mc; D.end = m.end;
                    op += (long long)rel + (long long)(m.out - spec_rel);
                    e = m.exit; endk = m.end;
                }
                mode = 0;
            }
        }
        if (D.count != 0xFFFFFFFFu) {
            a.desc[cb + k] = D;
            if (endk != kEndCont || op > (long long)d.dcap) { a.last_chunk[f] = k; mode = 3; }
        }
    }
}

__global__ void __launch_bounds__(64) lz4_stitch_kernel(Stitch2Args a) { stitch_body<false>(a); }
__global__ void __launch_bounds__(64) lz4_stitch_warp_kernel(Stitch2Args a) { stitch_body<true>(a); }

// ---- copy: one CTA per frame, one thread per sequence ---------------------------------------------------------
#ifndef B2B_COPY2_THREADS
#define B2B_COPY2_THREADS 128
#endif
constexpr int kCopy2Threads = B2B_COPY2_THREADS;
constexpr uint32_t kTile2 = 4096;                 // output bytes of one tile
constexpr uint32_t kRing2 = 2 * kTile2;           // the tile being filled + the one before it
constexpr uint32_t kLit2 = 64;                    // longest literal run a thread copies by itself
constexpr uint32_t kMatch2 = 32;                  // longest match a thread copies by itself

struct Copy2Args {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const FrameDec *fd;
    uint32_t nframes;
    uint8_t *dst, *scratch;
    const uint64_t *dst_off;
    const uint64_t *chunk_base;
    const ChunkDesc *desc;
    const uint32_t *last_chunk;
    const uint2 *table;
    const uint32_t *fallback;
    uint32_t *out_len, *status;
    FrameMeta *meta;
    const uint32_t *jump_state = nullptr;   // frames with state 1 were decoded by the pointer-jumping engine (lz4_decode4.cuh)
    uint32_t chunk_shift = kChunkShift;
};

struct CoopLit { uint32_t dst, n; const uint8_t *src; };
// a record that does not end inside the tile it was decoded for: kept for the next tile (a literal run of
// 100 KB has hundreds of length bytes; they are read once)
struct Carry { uint32_t tag, chunk, idx, ll, lit, off, flags; uint64_t vL, ml; };

__device__ __forceinline__ uint32_t bits_mask(uint32_t w, uint32_t a, uint32_t b) {   // bits of [a, b) inside word w
    uint32_t mask = 0xFFFFFFFFu;
    if (w == (a >> 5)) mask &= 0xFFFFFFFFu << (a & 31u);
    if (w == ((b - 1) >> 5)) mask &= 0xFFFFFFFFu >> (31u - ((b - 1) & 31u));
    return mask;
}
// pending-match bits of the tile: [a, b) set / cleared / tested (a < b).  Ranges of up to 32 bits (every match a
// thread handles alone) touch two words, addressed without a loop; pend[] has one word of padding behind it.
__device__ __forceinline__ void pend_set(uint32_t *pend, uint32_t a, uint32_t b) {
    if (b - a <= 32u) {
        const uint32_t m = b - a == 32u ? 0xFFFFFFFFu : (1u << (b - a)) - 1u, sh = a & 31u;
        atomicOr(&pend[a >> 5], m << sh);
        if (sh && (m >> (32u - sh))) atomicOr(&pend[(a >> 5) + 1], m >> (32u - sh));
    } else {
        for (uint32_t w = a >> 5; w <= ((b - 1) >> 5); w++) atomicOr(&pend[w], bits_mask(w, a, b));
    }
}
__device__ __forceinline__ void pend_clear(uint32_t *pend, uint32_t a, uint32_t b) {
    if (b - a <= 32u) {
        const uint32_t m = b - a == 32u ? 0xFFFFFFFFu : (1u << (b - a)) - 1u, sh = a & 31u;
        atomicAnd(&pend[a >> 5], ~(m << sh));
        if (sh && (m >> (32u - sh))) atomicAnd(&pend[(a >> 5) + 1], ~(m >> (32u - sh)));
    } else {
        for (uint32_t w = a >> 5; w <= ((b - 1) >> 5); w++) atomicAnd(&pend[w], ~bits_mask(w, a, b));
    }
}
__device__ __forceinline__ bool pend_any(const volatile uint32_t *pend, uint32_t a, uint32_t b) {
    if (b - a <= 32u) {
        const uint32_t m = b - a == 32u ? 0xFFFFFFFFu : (1u << (b - a)) - 1u;
        return (__funnelshift_r(pend[a >> 5], pend[(a >> 5) + 1], a & 31u) & m) != 0;
    }
    bool any = false;
    for (uint32_t w = a >> 5; w <= ((b - 1) >> 5); w++) any = any || (pend[w] & bits_mask(w, a, b)) != 0;
    return any;
}

// all threads of the CTA: tile bytes [0, upto) of the tile at v-position T0 go to outv (16-byte aligned); bytes
// before v-position a0 do not exist
__device__ __forceinline__ void flush_tile(const uint8_t *ring, uint8_t *outv, uint64_t T0, uint32_t upto, uint32_t a0) {
    const uint32_t rb = (uint32_t)T0 & (kRing2 - 1u);
    const uint32_t nvec = upto >> 4;
    for (uint32_t i = threadIdx.x; i < nvec; i += kCopy2Threads) {
        const uint64_t v = T0 + 16ull * i;
        if (v < a0) {                                     // the frame's first vector, output not 16-byte aligned
            for (uint32_t b = (uint32_t)(a0 - v); b < 16u; b++) outv[v + b] = ring[rb + 16u * i + b];
        } else {
            stg128(outv + v, *reinterpret_cast<const uint4 *>(ring + rb + 16u * i));
        }
    }
    const uint32_t tail = upto & 15u;
    if (threadIdx.x < tail) {
        const uint64_t v = T0 + 16ull * nvec + threadIdx.x;
        if (v >= a0) outv[v] = ring[rb + 16u * nvec + threadIdx.x];
    }
}

__global__ void __launch_bounds__(kCopy2Threads, 1024 / kCopy2Threads) lz4_copy2_kernel(Copy2Args a) {
    __shared__ __align__(16) uint8_t ring[kRing2];
    __shared__ uint32_t pend[kTile2 / 32 + 1];
    __shared__ CoopLit s_coop[kCopy2Threads];
    __shared__ Carry s_carry[2];
    __shared__ uint32_t s_ncoop[3], s_bulk[3], s_err[3];   // per-group flags, three slots in rotation (a fast warp is at most one group ahead)
    __shared__ const uint8_t *s_bulk_src[3];
    __shared__ unsigned long long s_total;
    const uint32_t f = blockIdx.x;
    const uint32_t tid = threadIdx.x;
    const int lane = (int)(tid & 31u);
    const FrameDec d = a.fd[f];
    if (d.kind == 0 || a.fallback[f]) return;             // status is final (prep) / the fallback kernel's
    if (a.jump_state && a.jump_state[f] == 1) return;     // decoded by the pointer-jumping engine
    if (blockIdx.y != 0 && d.kind != 1) return;           // only stored frames are shared between CTAs
    uint8_t *out = (d.mode ? a.scratch : a.dst) + a.dst_off[f];
    const uint8_t *fr = a.frames + a.frame_off[f];
    if (d.kind == 1) {                                    // stored frame (memcpy flag, blosc.go:393-401): gridDim.y CTAs share it
        const uint64_t slice = (((uint64_t)d.plen + gridDim.y - 1) / gridDim.y + 15ull) & ~15ull;
        const uint64_t lo = (uint64_t)blockIdx.y * slice;
        if (lo < d.plen) cta_copy(out + lo, fr + 16 + lo, d.plen - lo < slice ? d.plen - lo : slice);
        if (tid == 0 && blockIdx.y == 0) {
            FrameMeta m; m.mode = d.mode; m.typesize = d.typesize;
            a.status[f] = kOk; a.out_len[f] = d.norig; a.meta[f] = m;
        }
        return;
    }
    const uint8_t *__restrict__ src = fr + 16;
    const uint32_t a0 = (uint32_t)((uintptr_t)out & 15u);
    uint8_t *outv = out - a0;                             // v-space: v = output position + a0; outv + v is 16-byte aligned at v % 16 == 0
    const uint64_t vlimit = (uint64_t)a0 + d.dcap;
    for (uint32_t i = tid; i < kTile2 / 32 + 1; i += kCopy2Threads) pend[i] = 0;
    if (tid == 0) {
        s_ncoop[0] = s_ncoop[1] = s_ncoop[2] = 0; s_bulk[0] = s_bulk[1] = s_bulk[2] = 0;
        s_err[0] = s_err[1] = s_err[2] = 0xFFFFFFFFu; s_total = ~0ull; s_carry[0].tag = s_carry[1].tag = 0xFFFFFFFFu;
    }
    __syncthreads();
    uint64_t T0 = 0;                                      // v-position of the tile being filled
    int32_t W0 = (int32_t)a0;                             // ring holds the positions [T0 + W0, T0 + kTile2) (W0 <= 0 after the first tile)
    const uint32_t last = a.last_chunk[f];
    const uint64_t cb = a.chunk_base[f];
    uint32_t code = 0;                                    // 0 running, 1 malformed, 2 over capacity, 3 complete
    uint32_t it = 0;                                      // group counter (tags the carry, picks the flag slots)
    if (last == 0xFFFFFFFFu) { code = 3; if (tid == 0) s_total = 0; }
    for (uint32_t k = 0; code == 0 && k <= last; k++) {
        const ChunkDesc D = a.desc[cb + k];
        if (D.count == 0) continue;
        const uint2 *rec = a.table + (cb + k) * chunk_slot_records(a.chunk_shift) + D.start;
        uint32_t r = 0;
        while (r < D.count) {
            it++;
            const uint32_t slot = it % 3u;
            // ---- my record: positions, token, checks in stream order
            const uint32_t i = r + tid;
            const bool have = i < D.count;
            uint64_t vL = 0, vM = 0, vE = 0;                  // literal start / match start / end, v-space
            uint32_t ll = 0, lit = 0, off = 0, err = 0;
            uint64_t ml = 0;
            bool lit_ok = false, match_ok = false;
            if (have) {
                const Carry &c = s_carry[(it - 1u) & 1u];
                if (tid == 0 && c.tag == it - 1u && c.chunk == k && c.idx == i) {
                    vL = c.vL; ll = c.ll; lit = c.lit; off = c.off; ml = c.ml;
                    lit_ok = (c.flags & 1u) != 0; match_ok = (c.flags & 2u) != 0; err = c.flags >> 2;
                    vM = vL + ll; vE = vM + ml;
                } else {
                    const uint2 rc = rec[i], nx = rec[i + 1];
                    const long long o = (i < D.split ? D.base_a : D.base_b) + (long long)rc.y;
                    const long long on = (i + 1 < D.split ? D.base_a : D.base_b) + (long long)nx.y;
                    const uint32_t kind = i + 1 == D.count ? D.end : (uint32_t)kEndCont;
                    vL = (uint64_t)o + a0;
                    if (kind == kEndBadA) { err = 1; vM = vE = vL; }
                    else if (kind == kEndOverrunLit) { err = 2; vM = vE = vL; }
                    else {
                        uint32_t p = rc.x;
                        const uint32_t tok = src[p++];
                        ll = tok >> 4;
                        if (ll == 15u) {
                            uint32_t b;
                            do {
                                b = src[p++]; ll += b;
                                if (b == 255u) { const uint32_t sk = skip_ff_blocks(src, d.plen, p); p += sk; ll += 255u * sk; }
                            } while (b == 255u);
                        }
                        lit = p;
                        vM = vL + ll;
                        ml = (uint64_t)(on - o) - ll;
                        vE = vM + ml;
                        if (vM > vlimit) err = 2;
                        else if (kind == kEndBadB) { err = 1; lit_ok = true; }
                        else if (kind == kEndFinal) { lit_ok = true; s_total = (uint64_t)o + ll; }
                        else {
                            lit_ok = true;
                            off = (uint32_t)src[lit + ll] | ((uint32_t)src[lit + ll + 1] << 8);
                            if (off == 0 || (uint64_t)off > (uint64_t)o + ll) err = 1;
                            else if (kind == kEndOverrun || vE > vlimit) err = 2;
                            else match_ok = true;
                        }
                    }
                }
                if (err) atomicMin(&s_err[slot], (tid << 2) | err);
            }
            // ---- clip to the tile [T0, T0 + kTile2): everything below is 32-bit and relative to T0
            const uint64_t T1 = T0 + kTile2;
            auto clip = [&](uint64_t v) -> uint32_t {
                if (v > vlimit) v = vlimit;
                if (v <= T0) return 0u;
                return v >= T1 ? kTile2 : (uint32_t)(v - T0);
            };
            const uint32_t rb = (uint32_t)T0 & (kRing2 - 1u);      // the tile's place in the ring
            const uint32_t la = have ? clip(vL) : kTile2, lb = have ? clip(vM) : kTile2, mb = have ? clip(vE) : kTile2;
            const bool fin = have && (err != 0 || vE <= T1);
            // a literal run that covers this tile and at least the next one: whole tiles go straight from the
            // stream to the output
            const bool bulk = lit_ok && vL <= T0 && vM >= T1 + kTile2;
            if (bulk) { s_bulk[slot] = (uint32_t)((vM - T0) / kTile2); s_bulk_src[slot] = src + lit + (uint32_t)(T0 - vL); }
            // the same for a periodic match (a run of zeros, a repeated 2 / 4 / 8 / 16-byte element) that began at least
            // 16 bytes before this tile and covers it and the next one: every aligned 16-byte vector of it equals the
            // 16 bytes in front of the tile
            const bool mbulk = match_ok && off <= 16u && (off & (off - 1u)) == 0u && vM + 16u <= T0 && vE >= T1 + kTile2;
            if (mbulk) { s_bulk[slot] = 0x80000000u | (uint32_t)((vE - T0) / kTile2); s_bulk_src[slot] = nullptr; }
            // ---- literals: no ordering, the source is the stream
            if (lit_ok && lb > la && !bulk) {
                const uint32_t n = lb - la;
                const uint8_t *sp = src + lit + (uint32_t)(T0 + la - vL);
                if (n <= kLit2) {
                    uint8_t *dp = ring + rb + la;
                    for (uint32_t q = 0; q < n; q++) dp[q] = sp[q];
                } else {
                    const uint32_t cs = atomicAdd(&s_ncoop[slot], 1u);
                    s_coop[cs].dst = rb + la; s_coop[cs].n = n; s_coop[cs].src = sp;
                }
            }
            const bool mine = match_ok && mb > lb && !mbulk;
            if (mine) pend_set(pend, lb, mb);
            if (have && !fin && vL < T1) {                        // at most one record starts in the tile and ends behind it
                Carry &c = s_carry[it & 1u];
                c.chunk = k; c.idx = i; c.ll = ll; c.lit = lit; c.off = off; c.vL = vL; c.ml = ml;
                c.flags = (lit_ok ? 1u : 0u) | (match_ok ? 2u : 0u) | (err << 2);
                c.tag = it;
            }
            const uint32_t nfin = (uint32_t)__syncthreads_count(fin ? 1 : 0);
            // ---- (barrier passed: pending bits, literals, flags of this group are visible)
            if (tid == 0) { s_ncoop[(it + 2u) % 3u] = 0; s_bulk[(it + 2u) % 3u] = 0; s_err[(it + 2u) % 3u] = 0xFFFFFFFFu; }
            const uint32_t gerr = s_err[slot];
            if (gerr != 0xFFFFFFFFu) { code = gerr & 3u; break; }
            const uint32_t nbulk = s_bulk[slot] & 0x7FFFFFFFu;
            if (nbulk) {
                if (s_bulk[slot] & 0x80000000u) {
                    const uint4 pat = W0 <= -16 ? *reinterpret_cast<const uint4 *>(ring + ((rb - 16u) & (kRing2 - 1u)))
                                                : __ldcg(reinterpret_cast<const uint4 *>(outv + (T0 - 16u)));
                    const uint64_t nvec = (uint64_t)nbulk * (kTile2 / 16u);
                    for (uint64_t q = tid; q < nvec; q += kCopy2Threads) stg128(outv + T0 + 16ull * q, pat);
                } else {
                    cta_copy(outv + T0, s_bulk_src[slot], (uint64_t)nbulk * kTile2);
                }
                T0 += (uint64_t)nbulk * kTile2;
                W0 = 0;                                           // the ring holds nothing of what was just written
                __syncthreads();                                  // the copy is visible to later match reads of the CTA
                continue;                                         // same record again (from the carry)
            }
            const uint32_t nc = s_ncoop[slot];
            if (nc) {   // literal runs over kLit2 bytes: the whole CTA, 4 bytes per thread and step
                for (uint32_t c = 0; c < nc; c++) {
                    uint32_t dd = s_coop[c].dst, n = s_coop[c].n;
                    const uint8_t *sp = s_coop[c].src;
                    uint32_t head = (4u - (dd & 3u)) & 3u;
                    if (head > n) head = n;
                    if (tid < head) ring[dd + tid] = sp[tid];
                    dd += head; sp += head; n -= head;
                    const uint32_t nw = n >> 2;
                    const uint32_t sh = (uint32_t)((uintptr_t)sp & 3u);
                    const uint32_t *al = reinterpret_cast<const uint32_t *>(sp - sh);
                    uint32_t *dw = reinterpret_cast<uint32_t *>(ring + dd);
                    if (sh == 0) {
                        for (uint32_t w = tid; w < nw; w += kCopy2Threads) dw[w] = al[w];
                    } else {
                        for (uint32_t w = tid; w < nw; w += kCopy2Threads)
                            dw[w] = __funnelshift_r(al[w], al[w + 1], 8u * sh);
                    }
                    if (tid < (n & 3u)) ring[dd + 4u * nw + tid] = sp[4u * nw + tid];
                }
                __syncthreads();
            }
            // ---- matches: a match runs when none of its source bytes is a pending match byte.  Each warp
            // spins over its own lanes; other warps' progress shows in the pending bits.
            {
                bool pending = mine;
                const uint32_t mn = mine ? mb - lb : 0u;          // bytes of my match inside the tile
                const bool ovl = (uint64_t)off < ml;
                // bytes of this match before the tile (a match that began in an earlier tile); for a periodic
                // match only their place inside the period matters
                uint32_t skip = 0;
                if (mine) {
                    const uint64_t sk = T0 + lb - vM;
                    skip = ovl ? (uint32_t)(sk % off) : 0u;
                }
                // source, relative to T0: [s0, s0 + mn) without overlap, else the period [s0, s0 + off)
                const int32_t s0 = (int32_t)lb - (int32_t)skip - (int32_t)off;
                uint32_t sa = 0, sb = 0;                          // its part inside the tile
                if (mine) {
                    const int32_t s1 = s0 + (int32_t)(ovl ? off : mn);
                    if (s1 > 0) { sa = s0 > 0 ? (uint32_t)s0 : 0u; sb = (uint32_t)s1; }
                }
                uint32_t idle = 0;                                // turns in a row in which no lane of the warp could go
                while (__any_sync(0xffffffffu, pending)) {
                    const bool was = pending;
                    if (pending && mn <= kMatch2 && (sb <= sa || !pend_any(pend, sa, sb))) {
                        uint8_t *dp = ring + rb + lb;
                        if (s0 >= W0) {                           // the source is in the ring
                            if (!ovl) {
                                for (uint32_t q = 0; q < mn; q++) dp[q] = ring[(rb + (uint32_t)s0 + q) & (kRing2 - 1u)];
                            } else {
                                uint32_t rem = skip;
                                for (uint32_t q = 0; q < mn; q++) {
                                    dp[q] = ring[(rb + (uint32_t)s0 + rem) & (kRing2 - 1u)];
                                    rem = rem + 1u == off ? 0u : rem + 1u;
                                }
                            }
                        } else {                                  // (partly) behind the ring: from the output itself
                            uint32_t rem = skip;
                            for (uint32_t q = 0; q < mn; q++) {
                                const int32_t sp = s0 + (int32_t)(ovl ? rem : q);
                                dp[q] = sp >= W0 ? ring[(rb + (uint32_t)sp) & (kRing2 - 1u)] : __ldcg(outv + (T0 + (int64_t)sp));
                                rem = rem + 1u == off ? 0u : rem + 1u;
                            }
                        }
                        __threadfence_block();
                        pend_clear(pend, lb, mb);
                        pending = false;
                    }
                    // matches over kMatch2 bytes: 32 lanes wide
                    uint32_t longs = __ballot_sync(0xffffffffu, pending && mn > kMatch2);
                    while (longs) {
                        const int j = __ffs(longs) - 1;
                        longs &= longs - 1;
                        const uint32_t jsa = __shfl_sync(0xffffffffu, sa, j), jsb = __shfl_sync(0xffffffffu, sb, j);
                        bool busy = false;
                        if (jsb > jsa) {
                            for (uint32_t w0 = jsa >> 5; w0 <= ((jsb - 1) >> 5); w0 += 32) {
                                const uint32_t w = w0 + lane;
                                const bool b = w <= ((jsb - 1) >> 5) &&
                                               (((volatile uint32_t *)pend)[w] & bits_mask(w, jsa, jsb)) != 0;
                                if (__any_sync(0xffffffffu, b)) { busy = true; break; }
                            }
                        }
                        if (busy) continue;
                        const uint32_t jlb = __shfl_sync(0xffffffffu, lb, j), jmn = __shfl_sync(0xffffffffu, mn, j);
                        const int32_t js0 = __shfl_sync(0xffffffffu, s0, j);
                        const uint32_t joff = __shfl_sync(0xffffffffu, off, j), jskip = __shfl_sync(0xffffffffu, skip, j);
                        const bool jov = __shfl_sync(0xffffffffu, (uint32_t)ovl, j) != 0;
                        uint8_t *dp = ring + rb + jlb;
                        uint32_t rem = jov ? (jskip + (uint32_t)lane) % joff : 0u;   // place of my byte inside the period
                        const uint32_t step = jov ? 32u % joff : 0u;
                        // a run / short period (1, 2, 4, 8, 16 bytes): once 16 bytes of it stand at an aligned place, every
                        // further aligned 16-byte vector of the match is that same vector
                        uint32_t bytewise = jmn;
                        const bool vec = jov && joff <= 16u && (joff & (joff - 1u)) == 0u && jmn >= 64u;
                        if (vec) bytewise = ((16u - (jlb & 15u)) & 15u) + 16u;
                        for (uint32_t q = lane; q < bytewise; q += 32) {
                            const int32_t sp = js0 + (int32_t)(jov ? rem : q);
                            dp[q] = sp >= W0 ? ring[(rb + (uint32_t)sp) & (kRing2 - 1u)] : __ldcg(outv + (T0 + (int64_t)sp));
                            rem += step;
                            if (rem >= joff) rem -= joff;
                        }
                        if (vec) {
                            __syncwarp();
                            const uint4 pat = *reinterpret_cast<const uint4 *>(dp + bytewise - 16u);
                            const uint32_t nvec = (jmn - bytewise) >> 4;
                            for (uint32_t q = lane; q < nvec; q += 32) *reinterpret_cast<uint4 *>(dp + bytewise + 16u * q) = pat;
                            const uint32_t done = bytewise + 16u * nvec;
                            if (done + (uint32_t)lane < jmn) dp[done + lane] = dp[bytewise - 16u + lane];   // < 16 bytes behind the last vector
                        }
                        __threadfence_block();
                        __syncwarp();
                        for (uint32_t w = (jlb >> 5) + lane; w <= ((jlb + jmn - 1) >> 5); w += 32)
                            atomicAnd(&pend[w], ~bits_mask(w, jlb, jlb + jmn));
                        if (lane == j) pending = false;
                    }
                    // what this warp waits for is another warp's match: leave the issue slots to it
                    if (__any_sync(0xffffffffu, was && !pending)) idle = 0;
                    else { idle++; B2B_NANOSLEEP(idle < 4u ? 32u * idle : 128u); }
                }
            }
            // ---- progress; a full tile leaves
            const uint32_t nhave = D.count - r < (uint32_t)kCopy2Threads ? D.count - r : (uint32_t)kCopy2Threads;
            r += nfin;
            if (nfin < nhave) {
                __syncthreads();                                  // every match of the tile is done
                flush_tile(ring, outv, T0, kTile2, a0);
                T0 += kTile2;
                W0 = T0 - kTile2 > a0 ? -(int32_t)kTile2 : (int32_t)a0 - (int32_t)kTile2;
                if (W0 < -(int32_t)kTile2) W0 = -(int32_t)kTile2;
            }
        }
        if (code == 0 && D.end != kEndCont) code = D.end == kEndFinal ? 3u : 1u;
    }
    __syncthreads();
    uint32_t st, produced = 0;
    if (code == 3 && s_total != ~0ull) {
        const uint64_t total = s_total;                               // <= dcap
        const uint64_t vend = a0 + total;
        if (vend > T0) flush_tile(ring, outv, T0, (uint32_t)(vend - T0), a0);
        produced = (uint32_t)total;
        st = total == d.norig ? (uint32_t)kOk : (uint32_t)kESizeMismatch;      // blosc.go:429-431
    } else if (code == 2) {
        st = d.dcap == d.norig ? (uint32_t)kEDecompressionFailed : (uint32_t)kEDstTooSmall;
    } else {
        st = kEDecompressionFailed;                                   // blosc.go:410-413
    }
    if (tid == 0) {
        FrameMeta m; m.mode = 0; m.typesize = 0;
        if (st == kOk) { m.mode = d.mode; m.typesize = d.typesize; }
        a.status[f] = st; a.out_len[f] = produced; a.meta[f] = m;
    }
}

}  // namespace b2b
