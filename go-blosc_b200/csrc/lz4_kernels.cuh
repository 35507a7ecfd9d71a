// lz4_kernels.cuh -- K3 LZ4 block compression, K4 LZ4 block decompression, frame header
// parsing and the pack (gather) kernel.  One warp owns one frame.
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
constexpr int kHashLogDefault = 12;                 // per-warp hash table: 2^12 x u16 = 8 KiB

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

// Returns the number of bytes produced, -1 for a malformed stream, -2 if it would overrun cap.
// Strictness follows the oracle: zero offsets, offsets beyond the output so far, reads past
// the stream, writes past cap and a final token with a non-zero match nibble are errors;
// an empty stream decodes to 0 bytes (pierrec UncompressBlock, used at codec.go:79).
__device__ __forceinline__ int64_t warp_lz4_decode(const uint8_t *__restrict__ src, uint32_t clen,
                                                   uint8_t *dst, uint32_t cap, int lane) {
    if (clen == 0) return 0;
    uint32_t ip = 0, op = 0;
    for (;;) {
        if (ip >= clen) return -1;
        const uint32_t tok = src[ip++];
        uint64_t ll = tok >> 4;
        if (ll == 15 && !warp_read_len_ext(src, clen, ip, ll, lane)) return -1;
        if (ll > (uint64_t)(clen - ip)) return -1;
        if (ll > (uint64_t)(cap - op)) return -2;
        if (ll) warp_copy(dst + op, src + ip, (uint32_t)ll, lane);
        ip += (uint32_t)ll; op += (uint32_t)ll;
        uint64_t ml = tok & 15u;
        if (ip == clen) { if (ml != 0) return -1; break; }
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

__global__ void __launch_bounds__(kCodecThreads) lz4_decode_kernel(DecodeArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t f = blockIdx.x * kCodecWarps + warp;
    if (f >= a.nframes) return;
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
                const int64_t got = warp_lz4_decode(fr + 16, plen, out, dcap, lane);
                if (got == -1) st = kEDecompressionFailed;            // blosc.go:410-413
                else if (got == -2) st = dcap == norig ? kEDecompressionFailed : kEDstTooSmall;
                else if ((uint64_t)got != norig) st = kESizeMismatch; // blosc.go:429-431
                else produced = norig;
            }
        }
    }
    if (lane == 0) {
        if (st != kOk) { m.mode = 0; m.typesize = 0; produced = 0; }
        a.status[f] = st;
        a.out_len[f] = produced;
        a.meta[f] = m;
    }
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

// =========================================================================================
// K3: warp-cooperative LZ4 block compressor
// =========================================================================================
// 32 candidate positions per step (consecutive while matches are being found, spread out
// with the reference compressor's adaptive skip once literals pile up), a per-warp
// 2^HL x u16 hash table in shared memory, MATCH.ANY for repeats inside the step, a bounded
// look at the next four starts (the hit lanes already hold their candidates, so a slightly
// later but longer match costs nothing to find), backward and forward extension 32 lanes
// wide, cooperative literal copies.  Limits follow the reference's compressor (pierrec
// CompressBlock, called at codec.go:66): no match starts or is extended inside the last 14
// bytes, so the memcpy decision for short frames is the reference's.  The output is one
// valid LZ4 block per frame.
template <int HL> __device__ __forceinline__ uint32_t lz4_hash4(uint32_t seq) {
    return (seq * 2654435761u) >> (32 - HL);
}

constexpr uint32_t kLazyWindow = 2;   // later starts considered after the first hit
constexpr uint32_t kLazyWords = 4;    // bounded look-ahead: 4 + 4 * 4 = 20 bytes

// writes a length extension (value already reduced by 15) at out, returns bytes written
__device__ __forceinline__ uint32_t warp_put_len_ext(uint8_t *out, uint32_t v, int lane) {
    const uint32_t full = v / 255u, last = v - full * 255u;
    for (uint32_t i = lane; i < full; i += kWarp) out[i] = 255;
    if (lane == 0) out[full] = (uint8_t)last;
    return full + 1;
}

template <int HL>
__device__ __forceinline__ uint32_t warp_lz4_encode(const uint8_t *__restrict__ src, uint32_t n,
                                                    uint8_t *__restrict__ out, uint16_t *table,
                                                    int lane) {
    uint32_t op = 0, anchor = 0;
    if (n > 14) {
        for (uint32_t i = lane; i < (1u << HL); i += kWarp) table[i] = 0;
        __syncwarp();
        const uint32_t mfl = n - 14;      // a match may start at p < mfl (pierrec: sn = n - mfLimit)
        const uint32_t mlimit = n - 14;   // and is extended only below mlimit
        uint32_t si = 0;
        while (si < mfl) {
            // adaptive skip of the reference compressor: about 3 probes per 4 + lits/128 bytes
            const uint32_t lits = si - anchor;
            uint32_t stride = (4u + (lits >> 7)) / 3u;
            if (stride < 1) stride = 1;
            const uint64_t p64 = (uint64_t)si + (uint64_t)lane * stride;
            const bool valid = p64 < mfl;
            const uint32_t p = valid ? (uint32_t)p64 : 0u;
            const uint32_t seq = valid ? load32u(src + p) : 0u;
            const uint32_t h = lz4_hash4<HL>(seq);
            const uint32_t c16 = table[h];
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            const uint32_t same = __match_any_sync(0xffffffffu, seq) & vmask;
            // nearest earlier lane of this step with the same 4 bytes, else the table entry
            const uint32_t lower = same & ((1u << lane) - 1u);
            int64_t cand = 0;
            bool ok = false;
            if (valid && lower) {
                const uint32_t src_lane = 31u - (uint32_t)__clz((int)lower);
                const uint64_t dist = (uint64_t)(lane - src_lane) * stride;
                cand = (int64_t)p - (int64_t)dist;
                ok = dist < 65536;
            }
            if (valid && !ok) {
                cand = (int64_t)((p & ~0xFFFFu) + c16);
                if (cand >= (int64_t)p) cand -= 65536;
                ok = cand >= 0 && ((int64_t)p - cand) < 65536 && load32u(src + cand) == seq;
            }
            const uint32_t hit = __ballot_sync(0xffffffffu, ok);
            int pick = hit ? __ffs(hit) - 1 : 31;
            if (hit && stride == 1) {
                // bounded comparison of the first hit with the hits at the next kLazyWindow starts
                const uint32_t window = hit & (((2u << kLazyWindow) - 1u) << pick);
                uint32_t score = 0;
                if ((window >> lane) & 1u) {
                    uint32_t len = 4;
                    const uint8_t *a = src + p + 4, *b = src + (uint32_t)cand + 4;
                    if (p + 4 + 4 * kLazyWords <= mlimit) {
                        // all loads are issued before the first compare (one memory round trip)
                        uint32_t x[kLazyWords];
#pragma unroll
                        for (uint32_t k = 0; k < kLazyWords; k++) x[k] = load32u(a + 4 * k) ^ load32u(b + 4 * k);
                        bool open = true;
#pragma unroll
                        for (uint32_t k = 0; k < kLazyWords; k++) {
                            if (open) {
                                if (x[k]) { len += (uint32_t)(__ffs((int)x[k]) - 1) >> 3; open = false; }
                                else len += 4;
                            }
                        }
                    }
                    // longer wins; a later start pays one byte per position; ties go to the earlier lane
                    score = ((64u + len - (uint32_t)(lane - pick)) << 5) | (31u - (uint32_t)lane);
                }
                const uint32_t best = __reduce_max_sync(0xffffffffu, score);
                pick = 31 - (int)(best & 31u);
            }
            __syncwarp();
            // every probed position up to the chosen start is recorded (of lanes holding the
            // same 4 bytes the last one wins, so the table keeps the nearest occurrence)
            const uint32_t upto = same & ((2u << pick) - 1u);
            if (valid && lane <= pick && (upto >> lane) == 1u) table[h] = (uint16_t)p;
            __syncwarp();
            if (hit == 0) {
                const uint64_t nx = (uint64_t)si + 32ull * stride;
                si = nx < mfl ? (uint32_t)nx : mfl;
                continue;
            }
            uint32_t mp = __shfl_sync(0xffffffffu, p, pick);                         // match start
            uint32_t mc = (uint32_t)__shfl_sync(0xffffffffu, (uint32_t)cand, pick);  // its source
            const uint32_t offset = mp - mc;
            // forward extension from mp + 4, 128 bytes per step
            uint32_t mend = mp + 4;
            {
                uint32_t cpos = mc + 4;
                for (;;) {
                    const uint32_t a = mend + 4u * lane;
                    uint32_t x = 0xFFFFFFFFu;
                    if (a < mlimit) {
                        x = load32u(src + a) ^ load32u(src + cpos + 4u * lane);
                        const uint32_t avail = mlimit - a;
                        if (avail < 4) x |= 0xFFFFFFFFu << (8u * avail);
                    }
                    const uint32_t diff = __ballot_sync(0xffffffffu, x != 0);
                    if (diff == 0) { mend += 128; cpos += 128; continue; }
                    const int fl = __ffs(diff) - 1;
                    const uint32_t xf = __shfl_sync(0xffffffffu, x, fl);
                    mend += 4u * fl + ((uint32_t)(__ffs((int)xf) - 1) >> 3);
                    break;
                }
            }
            // backward extension over the pending literals
            while (mp > anchor) {
                const uint32_t k = lane + 1;
                const bool eq = (mp >= anchor + k) && (mc >= k) && src[mp - k] == src[mc - k];
                const uint32_t neq = ~__ballot_sync(0xffffffffu, eq);
                const uint32_t back = neq ? (uint32_t)(__ffs((int)neq) - 1) : 32u;
                mp -= back; mc -= back;
                if (back < 32) break;
            }
            // emit: token | literal length ext | literals | offset | match length ext
            const uint32_t ll = mp - anchor, ml = mend - mp - 4;
            const uint32_t tok_pos = op++;
            if (ll >= 15) op += warp_put_len_ext(out + op, ll - 15, lane);
            warp_copy(out + op, src + anchor, ll, lane);
            op += ll;
            if (lane == 0) {
                out[tok_pos] = (uint8_t)(((ll < 15 ? ll : 15u) << 4) | (ml < 15 ? ml : 15u));
                out[op] = (uint8_t)offset;
                out[op + 1] = (uint8_t)(offset >> 8);
            }
            op += 2;
            if (ml >= 15) op += warp_put_len_ext(out + op, ml - 15, lane);
            si = mend; anchor = mend;
        }
    }
    // last literals
    const uint32_t ll = n - anchor;
    const uint32_t tok_pos = op++;
    if (lane == 0) out[tok_pos] = (uint8_t)((ll < 15 ? ll : 15u) << 4);
    if (ll >= 15) op += warp_put_len_ext(out + op, ll - 15, lane);
    warp_copy(out + op, src + anchor, ll, lane);
    op += ll;
    return op;
}

struct EncodeArgs {
    const uint8_t *in;          // (shuffled) input, frame f at src_off[f]
    const uint64_t *src_off;
    const uint32_t *src_len;
    uint32_t nframes;
    uint8_t *comp;              // scratch: LZ4 block of frame f at comp_off[f]
    const uint64_t *comp_off;
    uint32_t *comp_len;         // out: payload bytes actually stored (c, or n for memcpy)
    uint32_t *frame_len;        // out: 16 + payload bytes (0 when status != 0)
    uint32_t *flags;            // out: header flags
    uint32_t *status;           // out
    uint32_t shuffle_flag;      // B2B_FLAG_SHUFFLE / B2B_FLAG_BITSHUFFLE / 0 (set even when T<=1)
    uint32_t keep_raw;          // 1: raw-block API, never substitute the memcpy payload
};

template <int HL>
__global__ void __launch_bounds__(kCodecThreads, HL <= 11 ? 12 : (HL == 12 ? 7 : 3)) lz4_encode_kernel(EncodeArgs a) {
    extern __shared__ __align__(16) uint16_t tables[];   // kCodecWarps x 2^HL entries
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t f = blockIdx.x * kCodecWarps + warp;
    if (f >= a.nframes) return;
    const uint32_t n = a.src_len[f];
    uint32_t st = kOk, c = 0, flags = a.shuffle_flag, flen = 0;
    if (n == 0) st = kEInvalidData;                       // blosc.go:269-271
    else if (n > 0xFFFFFFFFu - 16u) st = kEDataTooLarge;  // header fields are u32
    else {
        c = warp_lz4_encode<HL>(a.in + a.src_off[f], n, a.comp + a.comp_off[f],
                                tables + ((size_t)warp << HL), lane);
        if (c >= n && !a.keep_raw) { c = n; flags |= 0x2u; }  // blosc.go:342-345: store uncompressed
        flen = 16 + c;
    }
    if (lane == 0) {
        a.comp_len[f] = c; a.frame_len[f] = flen; a.flags[f] = flags; a.status[f] = st;
    }
}

// =========================================================================================
// Pack: header + payload of every frame gathered to its packed position
// =========================================================================================
struct PackArgs {
    const uint8_t *comp;        // LZ4 blocks (scratch)
    const uint64_t *comp_off;
    const uint8_t *raw;         // what a memcpy frame stores (shuffled bytes, or the caller's
    const uint64_t *src_off;    //   original bytes under B2B_OPT_REF_MEMCPY_QUIRK)
    const uint32_t *src_len;
    const uint32_t *comp_len;
    const uint32_t *flags;
    const uint32_t *status;
    const uint64_t *frame_off;  // packed offsets (16-byte aligned)
    uint8_t *dst;
    uint32_t nframes;
    uint32_t tiles_per_frame;
    uint32_t codec;
    uint32_t typesize_u8;       // uint8(opts.TypeSize)
};

__global__ void __launch_bounds__(kFilterThreads) pack_frames_kernel(PackArgs a) {
    const uint32_t f = blockIdx.x / a.tiles_per_frame, tile0 = blockIdx.x % a.tiles_per_frame;
    if (f >= a.nframes || a.status[f] != 0) return;
    const uint32_t n = a.src_len[f], c = a.comp_len[f], flags = a.flags[f];
    uint8_t *out = a.dst + a.frame_off[f];
    if (tile0 == 0 && threadIdx.x == 0) {
        // blosc.go:358-366: [2, codec, flags, uint8(T), n, n, 16 + c]
        uint4 h;
        h.x = 2u | (a.codec << 8) | (flags << 16) | (a.typesize_u8 << 24);
        h.y = n; h.z = n; h.w = 16u + c;
        *reinterpret_cast<uint4 *>(out) = h;
    }
    const uint8_t *payload = (flags & 0x2u) ? a.raw + a.src_off[f] : a.comp + a.comp_off[f];
    for (uint64_t t = tile0; t * kTileBytes < c; t += a.tiles_per_frame) {
        const uint64_t b = t * kTileBytes, len = c - b < kTileBytes ? c - b : kTileBytes;
        cta_copy(out + 16 + b, payload + b, len);
    }
}

}  // namespace b2b
