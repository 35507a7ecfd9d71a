// lz4_kernels.cuh -- K4 LZ4 block decompression and frame header parsing.  One warp owns one
// frame.  (K3, the compressor, is in lz4_encode.cuh.)
//
// Replaces the LZ4 branch of codec.go (lz4Codec.Compress/Decompress, codec.go:63-84, which
// call pierrec/lz4 v4.1.23 CompressBlock/UncompressBlock) and the frame assembly /
// validation of blosc.go:320-434.  The wire format is the reference's: 16-byte header and
// ONE raw LZ4 block (or the raw bytes when the memcpy flag is set).
#pragma once
#include "common.cuh"

namespace b2b {

constexpr int kCodecWarps = 4;                      // frames (warps) per CTA
constexpr int kCodecThreads = kCodecWarps * 32;

// status codes (mirror include/b2b.h)
enum : uint32_t {
    kOk = 0, kEInvalidData = 1, kEInvalidHeader = 2, kEInvalidVersion = 3, kEInvalidCodec = 4,
    kESizeMismatch = 5, kEDataTooLarge = 6, kECompressionFailed = 7, kEDecompressionFailed = 8,
    kEUnsupported = 10, kEDstTooSmall = 11
};

// unaligned little-endian 32-bit load; touches only the aligned words that hold p[0..3]
__device__ __forceinline__ uint32_t load32u(const uint8_t *p) {
    const uint32_t r = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t *q = reinterpret_cast<const uint32_t *>((uintptr_t)p - r);
    uint32_t lo = q[0];
    if (r == 0) return lo;
    return __funnelshift_r(lo, q[1], 8u * r);
}

// =========================================================================================
// K4: warp-cooperative LZ4 block decoder
// =========================================================================================
// Length extension bytes (0..255 each, ends at the first byte < 255), read 32 at a time.
__device__ __forceinline__ bool warp_read_len_ext(const uint8_t *src, uint32_t clen, uint32_t &ip,
                                                  uint64_t &len, int lane) {
    for (;;) {
        const uint32_t idx = ip + lane;
        const uint32_t b = idx < clen ? (uint32_t)src[idx] : 0x100u;
        const uint32_t stop = __ballot_sync(0xffffffffu, b != 255u);
        if (stop == 0) { len += 255u * 32u; ip += 32; if (len > 0xFFFFFFFFull) return false; continue; }
        const int first = __ffs(stop) - 1;
        const uint32_t bv = __shfl_sync(0xffffffffu, b, first);
        if (bv == 0x100u) return false;  // ran off the end of the stream
        len += 255u * (uint32_t)first + bv;
        ip += (uint32_t)first + 1;
        return true;
    }
}

// dst[0..ml) = periodic continuation of the `off` bytes before dst (LZ4 match semantics).
// All source bytes of one round are already written; rounds are separated by __syncwarp.
__device__ __forceinline__ void warp_match_copy(uint8_t *dst, uint32_t off, uint32_t ml, int lane) {
    const uint8_t *base = dst - off;
    if (off >= ml) { warp_copy(dst, base, ml, lane); return; }
    if (ml >= 64 && off <= 16 && (off & (off - 1)) == 0) {
        // period divides 16: every 16-byte aligned chunk of the output is the same vector
        const uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
        if ((uint32_t)lane < head) dst[lane] = base[lane & (off - 1)];
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) v |= (uint32_t)base[(head + 4 * k + b) & (off - 1)] << (8 * b);
            w[k] = v;
        }
        const uint4 pat = make_uint4(w[0], w[1], w[2], w[3]);
        uint8_t *d = dst + head;
        const uint32_t rem = ml - head, nvec = rem >> 4;
        for (uint32_t i = lane; i < nvec; i += kWarp) stg128(d + 16ull * i, pat);
        for (uint32_t i = (nvec << 4) + lane; i < rem; i += kWarp) d[i] = base[(head + i) & (off - 1)];
        return;
    }
    // general overlap: the written region doubles every round (always a multiple of off)
    uint32_t written = 0;
    while (written < ml) {
        uint32_t len = off + written;
        if (len > ml - written) len = ml - written;
        warp_copy(dst + written, base, len, lane);
        written += len;
        __syncwarp();
    }
}

// Lean variants for the batch loop (n <= 288: regular tokens): plain byte loops, no 16-byte realigning path,
// so that the loop's register footprint is not set by a copy routine it hardly ever needs at that size.
__device__ __forceinline__ void warp_copy_short(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    for (uint32_t i = lane; i < n; i += 2 * kWarp) {       // two loads in flight (the ranges never overlap)
        const bool p1 = i + kWarp < n;
        const uint8_t v0 = src[i];
        uint8_t v1 = 0;
        if (p1) v1 = src[i + kWarp];
        dst[i] = v0;
        if (p1) dst[i + kWarp] = v1;
    }
}
__device__ __forceinline__ void warp_match_copy_short(uint8_t *dst, uint32_t off, uint32_t ml, int lane) {
    const uint8_t *base = dst - off;
    if (off >= ml) { warp_copy_short(dst, base, ml, lane); return; }
    if (off == 1) {                                        // a run
        const uint8_t v = base[0];
        for (uint32_t i = lane; i < ml; i += kWarp) dst[i] = v;
        return;
    }
    uint32_t written = 0;                                  // general overlap: the written region doubles every round
    while (written < ml) {
        uint32_t len = off + written;
        if (len > ml - written) len = ml - written;
        warp_copy_short(dst + written, base, len, lane);
        written += len;
        __syncwarp();
    }
}

// ---- batched decode -----------------------------------------------------------------------
// A "regular" token carries at most one length-extension byte per field (literals <= 269,
// match <= 273).  The decoder parses such tokens 32 stream positions at a time: every lane
// decodes the token that WOULD start at its byte, the real chain of token starts is then walked
// with one shuffle per token, and the sequences found are appended to a per-warp table in shared
// memory.  A full table (up to 32 sequences, one per lane) is then copied out at once:
//   literals  every lane copies the first kLaneLit bytes of its own run, longer runs are finished
//             32 lanes wide;
//   matches   lanes whose source lies before the batch (or inside their own literals) copy their
//             own match, byte by byte, all at the same time; matches that read what another match
//             of the same batch writes, and long ones, follow in stream order, 32 lanes wide.
// Tokens with longer extensions, the closing token and anything malformed go through the
// sequence-at-a-time path below (warp_decode_one), which also reports every error.
#ifndef B2B_LANE_LIT
#define B2B_LANE_LIT 16
#endif
#ifndef B2B_LANE_MATCH
#define B2B_LANE_MATCH 16
#endif
constexpr uint32_t kLaneLit = B2B_LANE_LIT;       // literal bytes a lane copies by itself
constexpr uint32_t kLaneMatch = B2B_LANE_MATCH;   // longest match a lane copies by itself
constexpr uint32_t kBatchFill = 21;   // a window adds at most 11 tokens: parse while count <= 21

struct SeqTable {
    uint32_t a[32];   // offset | literals << 16
    uint32_t b[32];   // (literal position - batch start in the stream) | match length << 16
};

// Copies out `count` sequences, lane i holding sequence i as a = offset | literals << 16 and
// b = (literal position - batch start in the stream) | match length << 16 (0 for idle lanes).
// 0: ok, -1 malformed, -2 would overrun cap.
__device__ __forceinline__ int warp_copy_batch(const uint8_t *__restrict__ src, uint32_t batch_ip,
                                               uint8_t *dst, uint32_t &op, uint32_t cap, uint32_t count,
                                               uint32_t a, uint32_t b, int lane) {
    const bool act = (uint32_t)lane < count;
    const uint32_t off = a & 0xFFFFu, ll = a >> 16, ml = b >> 16, lrel = b & 0xFFFFu;
    const uint32_t tot = ll + ml;
    uint32_t incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31), excl = incl - tot;
    const bool ovf = act && incl > cap - op;
    const bool bad = act && (off == 0 || off > op + excl + ll);
    const uint32_t fail = __ballot_sync(0xffffffffu, ovf || bad);
    if (fail) {
        const uint32_t ob = __ballot_sync(0xffffffffu, ovf);
        return ((ob >> (__ffs(fail) - 1)) & 1u) ? -2 : -1;
    }
    // ---- literals (the source is the compressed stream: no ordering between lanes)
    uint8_t *d = dst + op + excl;
    const uint8_t *s = src + batch_ip + lrel;
    const uint32_t pl = ll < kLaneLit ? ll : kLaneLit;
    const uint32_t maxpl = __reduce_max_sync(0xffffffffu, pl);
    for (uint32_t i0 = 0; i0 < maxpl; i0 += 4) {
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k < pl) v[k] = s[i0 + k];
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k < pl) d[i0 + k] = (uint8_t)v[k];
    }
    uint32_t tails = __ballot_sync(0xffffffffu, ll > kLaneLit);
    while (tails) {
        const int j = __ffs(tails) - 1;
        tails &= tails - 1;
        const uint32_t dj = __shfl_sync(0xffffffffu, excl, j), sj = __shfl_sync(0xffffffffu, lrel, j);
        const uint32_t lj = __shfl_sync(0xffffffffu, ll, j);
        warp_copy_short(dst + op + dj + kLaneLit, src + batch_ip + sj + kLaneLit, lj - kLaneLit, lane);
    }
    __syncwarp();
    // ---- matches
    const uint32_t mpos = op + excl + ll;                 // where this lane's match starts
    const uint32_t first_match = op + __shfl_sync(0xffffffffu, ll, 0);
    const uint32_t s_end = mpos - off + (ml < off ? ml : off);
    const bool indep = act && ml <= kLaneMatch && (s_end <= first_match || off <= ll);
    const uint32_t pm = indep ? ml : 0u;
    const uint32_t maxpm = __reduce_max_sync(0xffffffffu, pm);
    {
        uint8_t *m0 = dst + mpos;
        const uint8_t *ms = m0 - off;                     // all reads stay inside [ms, m0)
        {
            uint32_t k = 0;
            for (uint32_t i0 = 0; i0 < maxpm; i0 += 4) {
                uint32_t v[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (i0 + q < pm) { v[q] = ms[k]; k = (k + 1 == off) ? 0u : k + 1; }
#pragma unroll
                for (int q = 0; q < 4; q++) if (i0 + q < pm) m0[i0 + q] = (uint8_t)v[q];
            }
        }
    }
    uint32_t rest = __ballot_sync(0xffffffffu, act && !indep);
    while (rest) {
        const int j = __ffs(rest) - 1;
        rest &= rest - 1;
        const uint32_t pj = __shfl_sync(0xffffffffu, mpos, j), oj = __shfl_sync(0xffffffffu, off, j);
        const uint32_t mj = __shfl_sync(0xffffffffu, ml, j);
        __syncwarp();
        warp_match_copy_short(dst + pj, oj, mj, lane);
    }
    op += total;
    return 0;
}

// the same, fed from the per-warp table in shared memory
__device__ __forceinline__ int warp_flush_batch(const uint8_t *__restrict__ src, uint32_t batch_ip,
                                                uint8_t *dst, uint32_t &op, uint32_t cap, uint32_t count,
                                                const SeqTable *tab, int lane) {
    __syncwarp();
    uint32_t a = 0, b = 0;
    if ((uint32_t)lane < count) { a = tab->a[lane]; b = tab->b[lane]; }
    return warp_copy_batch(src, batch_ip, dst, op, cap, count, a, b, lane);
}

// One sequence, 32 lanes wide.  1: sequence done, 0: that was the closing token (stream ends),
// -1 malformed, -2 would overrun cap.  Strictness follows the oracle: zero offsets, offsets
// beyond the output so far, reads past the stream, writes past cap and a final token with a
// non-zero match nibble are errors (pierrec UncompressBlock, used at codec.go:79).
__device__ __forceinline__ int warp_decode_one(const uint8_t *__restrict__ src, uint32_t clen, uint8_t *dst,
                                               uint32_t cap, uint32_t &ip, uint32_t &op, int lane,
                                               bool open_end = false) {
    // open_end: this is a sub-stream cut out of a block at a sequence boundary (decode index): it
    // ends after a complete sequence instead of with a literal-only token
    if (ip >= clen) return (open_end && ip == clen) ? 0 : -1;
    const uint32_t tok = src[ip++];
    uint64_t ll = tok >> 4;
    if (ll == 15 && !warp_read_len_ext(src, clen, ip, ll, lane)) return -1;
    if (ll > (uint64_t)(clen - ip)) return -1;
    if (ll > (uint64_t)(cap - op)) return -2;
    if (ll) warp_copy(dst + op, src + ip, (uint32_t)ll, lane);
    ip += (uint32_t)ll; op += (uint32_t)ll;
    uint64_t ml = tok & 15u;
    if (ip == clen) return ml != 0 ? -1 : 0;
    if (clen - ip < 2) return -1;
    const uint32_t off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8);
    ip += 2;
    if (off == 0 || off > op) return -1;
    if (ml == 15 && !warp_read_len_ext(src, clen, ip, ml, lane)) return -1;
    ml += 4;
    if (ml > (uint64_t)(cap - op)) return -2;
    __syncwarp();  // literals and earlier matches are visible to every lane
    warp_match_copy(dst + op, off, (uint32_t)ml, lane);
    op += (uint32_t)ml;
    __syncwarp();
    return 1;
}

// Returns the number of bytes produced, -1 for a malformed stream, -2 if it would overrun cap.
// An empty stream decodes to 0 bytes.
__device__ __forceinline__ int64_t warp_lz4_decode(const uint8_t *__restrict__ src, uint32_t clen,
                                                   uint8_t *dst, uint32_t cap, SeqTable *tab, int lane,
                                                   bool open_end = false) {
    if (clen == 0) return 0;
    uint32_t ip = 0, op = 0, count = 0, batch_ip = 0;
    // A window's speculative reads reach at most 305 bytes past its start (31 + token + two
    // extension bytes + 269 literals + offset).  The last stretch of the stream, where that could
    // leave it, is decoded a sequence at a time (which also holds all end-of-stream rules).
    constexpr uint32_t kWindowReach = 306;
    for (;;) {
        bool other = clen - ip < kWindowReach;             // the tail of the stream: no window, one sequence
        if (!other) {
            // ---- every lane decodes the token that would start at its byte of the window
            const uint32_t p = ip + lane;
            uint32_t nxt = 0xFFFFu, ta = 0, tb = 0;
            {
                const uint32_t t = src[p], e1 = src[p + 1];
                const bool x1 = (t >> 4) == 15u;
                const uint32_t ll = (t >> 4) + (x1 ? e1 : 0u);
                const uint32_t lit = p + 1 + (x1 ? 1u : 0u);
                const uint32_t oq = lit + ll;                     // offset bytes at oq, oq + 1
                const uint32_t off = (uint32_t)src[oq] | ((uint32_t)src[oq + 1] << 8);
                const uint32_t e2 = src[oq + 2];
                const bool x2 = (t & 15u) == 15u;
                const uint32_t ml = (t & 15u) + 4u + (x2 ? e2 : 0u);
                const bool regular = !(x1 && e1 == 255u) && !(x2 && e2 == 255u);
                if (regular) {
                    nxt = oq + 2u + (x2 ? 1u : 0u) - ip;
                    ta = off | (ll << 16);
                    tb = (lit - batch_ip) | (ml << 16);
                }
            }
            // ---- walk the chain of real token starts inside the window
            uint32_t starts = 0, cur = 0;
            while (cur < 32) {
                const uint32_t x = __shfl_sync(0xffffffffu, nxt, cur);
                if (x == 0xFFFFu) { other = true; break; }
                starts |= 1u << cur;
                cur = x;
            }
            if ((starts >> lane) & 1u) {
                const uint32_t slot = count + __popc(starts & ((1u << lane) - 1u));
                tab->a[slot] = ta; tab->b[slot] = tb;
            }
            count += __popc(starts);
            ip += cur;
        }
        if (other || count > kBatchFill) {
            if (count) {
                const int r = warp_flush_batch(src, batch_ip, dst, op, cap, count, tab, lane);
                if (r < 0) return r;
                count = 0;
                __syncwarp();
            }
            if (other) {
                const int r = warp_decode_one(src, clen, dst, cap, ip, op, lane, open_end);
                if (r < 0) return r;
                if (r == 0) break;
            }
            batch_ip = ip;
        }
    }
    return (int64_t)op;
}

// K4 copy half: the records of one frame, 32 at a time
__device__ __forceinline__ int64_t warp_lz4_copy(const uint8_t *__restrict__ src, uint32_t clen, uint8_t *dst,
                                                 uint32_t cap, const uint64_t *__restrict__ rec, uint32_t nrec,
                                                 int lane) {
    if (clen == 0) return 0;
    uint32_t ip = 0, op = 0, e = 0;
    bool tail = nrec == 0xFFFFFFFFu;                       // no table: sequence by sequence from the start
    for (;;) {
        bool one = tail;                                   // decode one sequence with warp_decode_one this turn?
        if (!tail) {
            const uint32_t rem = nrec - e;                 // nrec >= 1: the table always ends with a tail record
            const uint64_t r = (uint32_t)lane < rem ? rec[e + lane] : 0ull;
            const uint32_t kind = (uint32_t)(r >> 62);
            const uint32_t stop = __ballot_sync(0xffffffffu, (uint32_t)lane < rem && kind != 0);
            const uint32_t count = stop ? (uint32_t)(__ffs(stop) - 1) : (rem < 32u ? rem : 32u);
            if (count) {
                uint32_t a = 0, b = 0, insz = 0;
                if ((uint32_t)lane < count) {
                    a = (uint32_t)r;                       // offset | literals << 16
                    const uint32_t ll = a >> 16, ml = (uint32_t)(r >> 32) & 0xFFFFu;
                    insz = 1u + (ll >= 15u ? 1u : 0u) + ll + 2u + (ml >= 19u ? 1u : 0u);
                    b = ml << 16;
                }
                uint32_t incl = insz;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if ((uint32_t)lane < count) b |= incl - insz + 1u + ((a >> 16) >= 15u ? 1u : 0u);   // literal position
                const int rc = warp_copy_batch(src, ip, dst, op, cap, count, a, b, lane);
                if (rc < 0) return rc;
                ip += __shfl_sync(0xffffffffu, incl, 31);
                e += count;
                __syncwarp();
            }
            if (stop) {
                one = true;
                tail = __shfl_sync(0xffffffffu, kind, (int)count) != 1u;
                e += 1;
            } else if (e >= nrec) {
                return -1;                                 // cannot happen with a table from lz4_parse_kernel
            }
        }
        if (one) {
            const int rc = warp_decode_one(src, clen, dst, cap, ip, op, lane);
            if (rc < 0) return rc;
            if (rc == 0) break;
        }
    }
    return (int64_t)op;
}

struct DecodeArgs {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const uint32_t *frame_len;
    uint32_t nframes;
    int64_t typesize_override;
    uint8_t *dst;       // final destination
    uint8_t *scratch;   // staging for frames that still need an unshuffle (same offsets)
    const uint64_t *dst_off;
    const uint32_t *dst_cap;
    uint32_t *out_len;
    uint32_t *status;
    FrameMeta *meta;    // filter still to run on frame f (mode 0: none)
    // split decode only: the sequence records lz4_parse_kernel wrote
    const uint64_t *table;
    const uint64_t *table_off;
    const uint32_t *nrec;
    // fused kernel only: when not null, only the frames with only[f] != 0 are decoded (the frames the
    // chunk-parallel decoder of lz4_decode2.cuh had no table room for)
    const uint32_t *only;
    uint32_t fuse_unshuffle;    // 1: a byte-shuffled frame (typesize 2 / 4, aligned) is un-shuffled by the warp that decoded it
    unsigned long long *ticket; // not null: persistent warps, frames handed out in order (zero before the launch)
};

__device__ __forceinline__ uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// header checks in the order of blosc.go:296-303, 165-185, 385-390
__device__ __forceinline__ uint32_t check_header(const uint8_t *fr, uint32_t flen, uint32_t &flags,
                                                 uint32_t &codec, uint32_t &tsz, uint32_t &norig,
                                                 uint32_t &ncomp) {
    if (flen < 16) return kEInvalidHeader;
    const uint32_t ver = fr[0];
    codec = fr[1]; flags = fr[2]; tsz = fr[3];
    norig = rd32(fr + 4); ncomp = rd32(fr + 12);
    if (ver != 2) return kEInvalidVersion;
    if (ncomp > flen || ncomp < 16) return kEInvalidData;
    return kOk;
}

// ---- fused K1 epilogue (decompressBackend un-shuffles inside the same call, blosc.go:417-426) --------------
// The warp that decoded a byte-shuffled frame into the stage buffer un-shuffles it straight into dst, while the
// other warps of the SM are still decoding: the HBM-bound transpose runs under the issue-bound decode instead
// of as a kernel of its own behind it.  Typesizes 2 and 4 with 16-byte aligned bases and E % 16 == 0 (every
// frame of C3 / C5); anything else is left to filter_batch_kernel (meta keeps its mode).
// One lane turns 16 elements per step: a 16-byte vector from each plane in, T vectors out.
__device__ __forceinline__ void warp_unshuffle_frame(const uint8_t *sp, uint8_t *__restrict__ dp, uint32_t T, uint32_t E,
                                                     uint32_t n, int lane) {
    const uint32_t nq = E >> 4;
    if (T == 4) {
        for (uint32_t q = lane; q < nq; q += kWarp) {
            const uint4 p0 = __ldcg(reinterpret_cast<const uint4 *>(sp + 16ull * q));
            const uint4 p1 = __ldcg(reinterpret_cast<const uint4 *>(sp + (uint64_t)E + 16ull * q));
            const uint4 p2 = __ldcg(reinterpret_cast<const uint4 *>(sp + 2ull * E + 16ull * q));
            const uint4 p3 = __ldcg(reinterpret_cast<const uint4 *>(sp + 3ull * E + 16ull * q));
            uint4 o;
            uint8_t *d = dp + 64ull * q;
            transpose4x4(p0.x, p1.x, p2.x, p3.x, o.x, o.y, o.z, o.w); stg128_stream(d, o);
            transpose4x4(p0.y, p1.y, p2.y, p3.y, o.x, o.y, o.z, o.w); stg128_stream(d + 16, o);
            transpose4x4(p0.z, p1.z, p2.z, p3.z, o.x, o.y, o.z, o.w); stg128_stream(d + 32, o);
            transpose4x4(p0.w, p1.w, p2.w, p3.w, o.x, o.y, o.z, o.w); stg128_stream(d + 48, o);
        }
    } else {   // T == 2
        for (uint32_t q = lane; q < nq; q += kWarp) {
            const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(sp + 16ull * q));
            const uint4 b = __ldcg(reinterpret_cast<const uint4 *>(sp + (uint64_t)E + 16ull * q));
            uint8_t *d = dp + 32ull * q;
            stg128_stream(d, make_uint4(prmt(a.x, b.x, 0x5140), prmt(a.x, b.x, 0x7362), prmt(a.y, b.y, 0x5140), prmt(a.y, b.y, 0x7362)));
            stg128_stream(d + 16, make_uint4(prmt(a.z, b.z, 0x5140), prmt(a.z, b.z, 0x7362), prmt(a.w, b.w, 0x5140), prmt(a.w, b.w, 0x7362)));
        }
    }
    for (uint32_t i = E * T + lane; i < n; i += kWarp) dp[i] = __ldcg(sp + i);   // the n % T bytes behind the last element
}
__device__ __forceinline__ bool unshuffle_fusable(const FrameMeta &m, const uint8_t *sp, const uint8_t *dp, uint32_t n) {
    if (m.mode != 1 || (m.typesize != 4 && m.typesize != 2)) return false;
    const uint32_t E = n / m.typesize;
    return E >= 16 && (E & 15u) == 0 && ((((uintptr_t)sp) | ((uintptr_t)dp)) & 15u) == 0;
}

// one frame, one warp
template <bool kSplit>
__device__ __forceinline__ void warp_decode_frame(const DecodeArgs &a, uint32_t f, SeqTable *seq_table, int lane) {
    if (!kSplit && a.only && !a.only[f]) return;
    const uint8_t *fr = a.frames + a.frame_off[f];
    const uint32_t flen = a.frame_len[f];
    uint32_t flags = 0, codec = 0, tsz = 0, norig = 0, ncomp = 0;
    uint32_t st = check_header(fr, flen, flags, codec, tsz, norig, ncomp);
    FrameMeta m; m.mode = 0; m.typesize = 0;
    uint32_t produced = 0;
    if (st == kOk) {
        const bool is_memcpy = (flags & 0x2u) != 0;
        if (!is_memcpy) {
            if (codec < 1 || codec > 5) st = kEInvalidCodec;          // blosc.go:403-407
            else if (codec != 1 && codec != 2) st = kEUnsupported;    // Snappy/ZLIB/ZSTD: host side
        }
        if (st == kOk) {
            const uint32_t plen = ncomp - 16;
            // effective typesize and filter (blosc.go:417-426): bitshuffle flag wins
            uint64_t T = a.typesize_override > 0 ? (uint64_t)a.typesize_override : (uint64_t)tsz;
            uint32_t mode = (flags & 0x4u) ? 2u : ((flags & 0x1u) ? 1u : 0u);
            const bool active = mode != 0 && T > 1 && (uint64_t)norig >= T;
            m.mode = active ? mode : 0u;
            m.typesize = active ? (uint32_t)T : 0u;
            // the caller may give less room than NBytesOrig when the stream provably cannot
            // reach it (a hostile header): the decode then ends short -> size mismatch
            const uint32_t cap = a.dst_cap[f];
            uint8_t *out = (active ? a.scratch : a.dst) + a.dst_off[f];
            if (is_memcpy) {
                if (plen != norig) st = kESizeMismatch;               // blosc.go:398-400,429-431
                else if (cap < norig) st = kEDstTooSmall;
                else { warp_copy(out, fr + 16, plen, lane); produced = plen; }
            } else {
                const uint32_t dcap = cap < norig ? cap : norig;
                const int64_t got = kSplit ? warp_lz4_copy(fr + 16, plen, out, dcap, a.table + a.table_off[f], a.nrec[f], lane)
                                           : warp_lz4_decode(fr + 16, plen, out, dcap, seq_table, lane);
                if (got == -1) st = kEDecompressionFailed;            // blosc.go:410-413
                else if (got == -2) st = dcap == norig ? kEDecompressionFailed : kEDstTooSmall;
                else if ((uint64_t)got != norig) { st = kESizeMismatch; produced = (uint32_t)got; }  // blosc.go:429-431
                else produced = norig;
            }
        }
    }
    if (st == kOk && a.fuse_unshuffle) {
        const uint8_t *sp = a.scratch + a.dst_off[f];
        uint8_t *dp = a.dst + a.dst_off[f];
        if (unshuffle_fusable(m, sp, dp, norig)) {
            __syncwarp();                                   // the frame is complete in the stage buffer
            warp_unshuffle_frame(sp, dp, m.typesize, norig / m.typesize, norig, lane);
            m.mode = 0; m.typesize = 0;                     // nothing left for filter_batch_kernel
        }
    }
    if (lane == 0) {
        if (st != kOk) { m.mode = 0; m.typesize = 0; if (st != kESizeMismatch) produced = 0; }
        a.status[f] = st;
        a.out_len[f] = produced;
        a.meta[f] = m;
    }
}

// With a ticket (zero before the launch) the warps are persistent and take the next frame as soon as they are done
// with one: no warp waits for the slowest frame of its CTA, and the launch does not end on a tail of half-empty
// SMs.  Without one, warp w of CTA b takes frame 4 b + w.
template <bool kSplit>
__global__ void __launch_bounds__(kCodecThreads, 8) lz4_decode_kernel(DecodeArgs a) {
    __shared__ SeqTable seq_tables[kCodecWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!a.ticket) {
        const uint32_t f = blockIdx.x * kCodecWarps + warp;
        if (f < a.nframes) warp_decode_frame<kSplit>(a, f, &seq_tables[warp], lane);
        return;
    }
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= a.nframes) break;
        warp_decode_frame<kSplit>(a, (uint32_t)t, &seq_tables[warp], lane);
        __syncwarp();
    }
}

// ---- split decode: parse kernel -> sequence table -> copy kernel ----------------------------------
// The fused kernel above runs at 64 registers / 32 warps per SM because of its copy half; its parse
// half needs 28.  Split, the parse runs at 48 warps per SM (the whole 8 GiB C3 batch in 4.4 ms against
// ~7 ms inside the fused kernel) and writes one 8-byte record per sequence; the copy kernel then
// reads 32 records per batch instead of parsing.  Records (u64), kind in bits 62..63:
//   0  regular sequence: offset | literals << 16 | match length << 32
//   1  a token the window parse does not handle starts at stream position (low 32 bits): the copy
//      kernel decodes that one sequence with warp_decode_one
//   2  from stream position (low 32 bits) on, decode sequence by sequence to the end (the last 306
//      bytes of a stream, the closing token, anything malformed, a full table)
// All error reporting stays in the copy kernel (offsets, capacity, end-of-stream rules).
// a frame's table holds dst_cap / 4 + kSeqSlack records: every sequence but the last emits >= 4 bytes

struct ParseArgs {
    const uint8_t *frames;
    const uint64_t *frame_off;
    const uint32_t *frame_len;
    const uint32_t *dst_cap;
    uint32_t nframes;
    uint64_t *table;            // all frames' records
    const uint64_t *table_off;  // first record of frame f (exclusive scan of dst_cap / 4 + kSeqSlack)
    uint32_t *nrec;             // records written for frame f (~0: no room in the table, decode without it)
    uint64_t table_cap;         // records the table can hold in all
    unsigned long long *ticket = nullptr;   // not null: persistent warps (zero before the launch)
};

// the sequence records of one LZ4 block; returns how many were written (room >= 15)
__device__ __forceinline__ uint32_t warp_lz4_parse(const uint8_t *__restrict__ src, uint32_t clen, uint64_t *tab,
                                                   uint32_t room, int lane) {
    uint32_t ip = 0, n = 0;
    constexpr uint32_t kWindowReach = 306;
    constexpr uint64_t kLong = 1ull << 62, kTail = 2ull << 62;
    for (;;) {
        if (clen - ip < kWindowReach || n + 14 > room) { if (lane == 0) tab[n] = kTail | ip; n++; break; }
        const uint32_t p = ip + lane;
        uint32_t nxt = 0xFFFFu;
        uint64_t rec = 0;
        {
            const uint32_t t = src[p], e1 = src[p + 1];
            const bool x1 = (t >> 4) == 15u;
            const uint32_t ll = (t >> 4) + (x1 ? e1 : 0u);
            const uint32_t oq = p + 1 + (x1 ? 1u : 0u) + ll;                  // offset bytes at oq, oq + 1
            const uint32_t off = (uint32_t)src[oq] | ((uint32_t)src[oq + 1] << 8);
            const uint32_t e2 = src[oq + 2];
            const bool x2 = (t & 15u) == 15u;
            const uint32_t ml = (t & 15u) + 4u + (x2 ? e2 : 0u);
            if (!(x1 && e1 == 255u) && !(x2 && e2 == 255u)) {
                nxt = oq + 2u + (x2 ? 1u : 0u) - ip;
                rec = (uint64_t)(off | (ll << 16)) | ((uint64_t)ml << 32);
            }
        }
        uint32_t starts = 0, cur = 0;
        bool other = false;
        while (cur < 32) {
            const uint32_t x = __shfl_sync(0xffffffffu, nxt, cur);
            if (x == 0xFFFFu) { other = true; break; }
            starts |= 1u << cur;
            cur = x;
        }
        if ((starts >> lane) & 1u) tab[n + __popc(starts & ((1u << lane) - 1u))] = rec;
        n += __popc(starts);
        ip += cur;
        if (other) {
            // a token with longer extensions (or the closing token, or nonsense): find its end, lengths only
            uint32_t q = ip;
            const uint32_t tok = src[q++];
            uint64_t ll = tok >> 4, ml = tok & 15u;
            bool ok = !(ll == 15 && !warp_read_len_ext(src, clen, q, ll, lane)) && ll <= (uint64_t)(clen - q);
            if (ok) { q += (uint32_t)ll; ok = q != clen && clen - q >= 2; }
            if (ok) { q += 2; ok = !(ml == 15 && !warp_read_len_ext(src, clen, q, ml, lane)); }
            if (!ok) { if (lane == 0) tab[n] = kTail | ip; n++; break; }      // the copy kernel meets it and decides
            if (lane == 0) tab[n] = kLong | ip;
            n++;
            ip = q;
        }
    }
    return n;
}

__device__ __forceinline__ void warp_parse_frame(const ParseArgs &a, uint32_t f, int lane) {
    const uint8_t *fr = a.frames + a.frame_off[f];
    uint32_t flags = 0, codec = 0, tsz = 0, norig = 0, ncomp = 0;
    const uint32_t st = check_header(fr, a.frame_len[f], flags, codec, tsz, norig, ncomp);
    if (st != kOk || (flags & 0x2u) || (codec != 1 && codec != 2)) { if (lane == 0) a.nrec[f] = 0; return; }
    const uint32_t room = a.dst_cap[f] / 4u + kSeqSlack;
    if (a.table_off[f] + room > a.table_cap) {            // capacities that overlap in dst: not sized for
        if (lane == 0) a.nrec[f] = 0xFFFFFFFFu;
        return;
    }
    const uint32_t n = warp_lz4_parse(fr + 16, ncomp - 16, a.table + a.table_off[f], room, lane);
    if (lane == 0) a.nrec[f] = n;
}

__global__ void __launch_bounds__(kCodecThreads, 12) lz4_parse_kernel(ParseArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!a.ticket) {
        const uint32_t f = blockIdx.x * kCodecWarps + warp;
        if (f < a.nframes) warp_parse_frame(a, f, lane);
        return;
    }
    for (;;) {                                                  // persistent warps (see lz4_decode_kernel)
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= a.nframes) break;
        warp_parse_frame(a, (uint32_t)t, lane);
        __syncwarp();
    }
}

// ---- the same two halves over a table of bare LZ4 streams (the blocks of blocks.cuh) ----------------
struct StreamArgs {
    const uint8_t *src;         // stream t: src + src_off[t], clen[t] bytes
    const uint64_t *src_off;
    const uint32_t *clen;
    uint8_t *dst, *scratch;     // decodes to exactly cap[t] bytes at (kind & 4 ? scratch : dst) + dst_off[t]
    const uint64_t *dst_off;
    const uint32_t *cap;
    const uint32_t *kind;       // low 2 bits: 0 nothing, 1 stored (copy cap bytes), 2 one LZ4 block, 3 not ours
    const uint32_t *owner;      // status[owner[t]] is raised when stream t fails
    uint32_t *status;
    uint32_t nstreams;
    uint64_t *table;            // sequence records
    const uint64_t *table_off;  // exclusive scan of cap / 4 + kSeqSlack
    uint32_t *nrec;             // ~0: no room in the table (outputs that overlap in dst), decoded without it
    uint64_t table_cap;         // records the table can hold in all
};

__global__ void __launch_bounds__(kCodecThreads, 12) lz4_parse_streams_kernel(StreamArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * kCodecWarps + warp;
    if (t >= a.nstreams || (a.kind[t] & 3u) != 2u) return;
    const uint32_t room = a.cap[t] / 4u + kSeqSlack;
    if (a.table_off[t] + room > a.table_cap) { if (lane == 0) a.nrec[t] = 0xFFFFFFFFu; return; }
    const uint32_t n = warp_lz4_parse(a.src + a.src_off[t], a.clen[t], a.table + a.table_off[t], room, lane);
    if (lane == 0) a.nrec[t] = n;
}

__global__ void __launch_bounds__(kCodecThreads, 8) lz4_copy_streams_kernel(StreamArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * kCodecWarps + warp;
    if (t >= a.nstreams) return;
    const uint32_t kind = a.kind[t];
    if ((kind & 3u) == 0u || (kind & 3u) == 3u) return;
    const uint8_t *src = a.src + a.src_off[t];
    uint8_t *out = ((kind & 4u) ? a.scratch : a.dst) + a.dst_off[t];
    const uint32_t cap = a.cap[t];
    if ((kind & 3u) == 1u) { warp_copy(out, src, cap, lane); return; }
    const int64_t got = warp_lz4_copy(src, a.clen[t], out, cap, a.table + a.table_off[t], a.nrec[t], lane);
    if (got != (int64_t)cap && lane == 0) atomicMax(a.status + a.owner[t], (uint32_t)kEDecompressionFailed);
}

// ---- decode with a side-car index ------------------------------------------
// Frames that K3 produced with independent segments come with an index of sequence boundaries: one
// (payload offset, output position) pair per 64 KiB segment.  One warp decodes one sub-stream; the
// frame itself is an ordinary single LZ4 block (the reference decodes it without the index).
// status[] must be zero before the launch; index_finish_kernel turns it into out_len / meta.
struct IndexedDecodeArgs {
    DecodeArgs d;
    const uint64_t *index;      // segs_per_frame entries per frame, ~0 = no entry
    uint32_t segs_per_frame;
    unsigned long long *ticket; // zero before launch: warps pull (frame, segment) items from it
};

// One (frame, segment) item.  Most items are empty (segments that open no sequence, e.g. the
// incompressible planes of a shuffled frame), which is why the warps of the persistent kernel below
// pull items from a ticket instead of owning one each: a CTA slot is never held by idle warps.
__device__ __forceinline__ void warp_decode_index_item(const IndexedDecodeArgs &ia, uint32_t f, uint32_t sg,
                                                       SeqTable *tab, int lane) {
    const DecodeArgs &a = ia.d;
    const uint8_t *fr = a.frames + a.frame_off[f];
    const uint32_t flen = a.frame_len[f];
    uint32_t flags = 0, codec = 0, tsz = 0, norig = 0, ncomp = 0;
    uint32_t st = check_header(fr, flen, flags, codec, tsz, norig, ncomp);
    const bool is_memcpy = (flags & 0x2u) != 0;
    if (st == kOk && !is_memcpy) {
        if (codec < 1 || codec > 5) st = kEInvalidCodec;
        else if (codec != 1 && codec != 2) st = kEUnsupported;
    }
    if (st == kOk && a.dst_cap[f] < norig) st = kEDstTooSmall;
    if (st != kOk) { if (sg == 0 && lane == 0) atomicMax(a.status + f, st); return; }
    const uint32_t plen = ncomp - 16;
    const uint64_t T = a.typesize_override > 0 ? (uint64_t)a.typesize_override : (uint64_t)tsz;
    const uint32_t mode = (flags & 0x4u) ? 2u : ((flags & 0x1u) ? 1u : 0u);
    const bool active = mode != 0 && T > 1 && (uint64_t)norig >= T;
    uint8_t *out = (active ? a.scratch : a.dst) + a.dst_off[f];
    if (is_memcpy) {                                            // 64 KiB slices of the raw copy
        if (plen != norig) { if (sg == 0 && lane == 0) atomicMax(a.status + f, (uint32_t)kESizeMismatch); return; }
        for (uint64_t b = (uint64_t)sg * 65536u; b < norig; b += (uint64_t)ia.segs_per_frame * 65536u) {
            const uint32_t n = norig - b < 65536u ? (uint32_t)(norig - b) : 65536u;
            warp_copy(out + b, fr + 16 + b, n, lane);
        }
        return;
    }
    const uint64_t *idx = ia.index + (uint64_t)f * ia.segs_per_frame;
    uint64_t e0 = idx[sg];
    if (e0 == ~0ull) { if (sg != 0) return; e0 = 0; }          // the stream itself starts at (0, 0)
    uint64_t e1 = ~0ull;
    for (uint32_t k = sg + 1; k < ia.segs_per_frame && e1 == ~0ull; k++) e1 = idx[k];
    const uint32_t ip0 = (uint32_t)e0, op0 = (uint32_t)(e0 >> 32);
    const bool last = e1 == ~0ull;
    const uint32_t ip1 = last ? plen : (uint32_t)e1, op1 = last ? norig : (uint32_t)(e1 >> 32);
    if (ip0 > ip1 || ip1 > plen || op0 > op1 || op1 > norig) {  // a corrupt index must not become a wild copy
        if (lane == 0) atomicMax(a.status + f, (uint32_t)kEDecompressionFailed);
        return;
    }
    if (ip0 == ip1 && op0 == op1) return;
    const int64_t got = warp_lz4_decode(fr + 16 + ip0, ip1 - ip0, out + op0, op1 - op0, tab, lane, !last);
    if (got < 0) { if (lane == 0) atomicMax(a.status + f, (uint32_t)kEDecompressionFailed); }
    else if ((uint64_t)got != op1 - op0) { if (lane == 0) atomicMax(a.status + f, (uint32_t)kESizeMismatch); }
}

__global__ void __launch_bounds__(kCodecThreads, 8) lz4_decode_indexed_kernel(IndexedDecodeArgs ia) {
    __shared__ SeqTable seq_tables[kCodecWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t items = (uint64_t)ia.d.nframes * ia.segs_per_frame;
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(ia.ticket, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= items) break;
        warp_decode_index_item(ia, (uint32_t)(item / ia.segs_per_frame), (uint32_t)(item % ia.segs_per_frame),
                               &seq_tables[warp], lane);
        __syncwarp();
    }
}

// out_len / meta of the indexed decode, from the accumulated status and the header
__global__ void index_finish_kernel(DecodeArgs a) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.nframes) return;
    uint32_t flags = 0, codec = 0, tsz = 0, norig = 0, ncomp = 0;
    const uint32_t hs = check_header(a.frames + a.frame_off[f], a.frame_len[f], flags, codec, tsz, norig, ncomp);
    const uint32_t st = a.status[f];
    FrameMeta m; m.mode = 0; m.typesize = 0;
    uint32_t produced = 0;
    if (st == kOk && hs == kOk) {
        const uint64_t T = a.typesize_override > 0 ? (uint64_t)a.typesize_override : (uint64_t)tsz;
        const uint32_t mode = (flags & 0x4u) ? 2u : ((flags & 0x1u) ? 1u : 0u);
        if (mode != 0 && T > 1 && (uint64_t)norig >= T) { m.mode = mode; m.typesize = (uint32_t)T; }
        produced = norig;
    }
    a.out_len[f] = produced;
    a.meta[f] = m;
}

// The staging buffer of a decompress batch is sized from the caller's total_dst_bytes: an output slot that
// reaches beyond it gets capacity 0 (the frame then reports B2B_EDST_TOO_SMALL) instead of a wild write.
__global__ void clip_caps_kernel(const uint64_t *dst_off, const uint32_t *dst_cap, uint64_t total_dst, uint32_t nframes,
                                 uint32_t *cap_eff) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    const uint64_t o = dst_off[f];
    cap_eff[f] = (o <= total_dst && dst_cap[f] <= total_dst - o) ? dst_cap[f] : 0u;
}

// header-only pass for b2b_frame_info_batch_dev
__global__ void frame_info_kernel(const uint8_t *frames, const uint64_t *frame_off,
                                  const uint32_t *frame_len, uint32_t nframes, uint32_t *orig_len,
                                  uint32_t *status) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    uint32_t flags, codec, tsz, norig = 0, ncomp;
    const uint32_t st = check_header(frames + frame_off[f], frame_len[f], flags, codec, tsz, norig, ncomp);
    // only the header-level checks of ParseHeader (blosc.go:165-185) gate the size
    const bool have = st == kOk || st == kEInvalidData;
    orig_len[f] = have ? norig : 0u;
    status[f] = (st == kEInvalidData) ? kOk : st;
}

}  // namespace b2b
