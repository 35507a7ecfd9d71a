// lz4_encode.cuh -- K3: LZ4 block compression of a batch of frames, segment-parallel.
//
// Replaces lz4Codec.Compress (codec.go:63-75 -> pierrec/lz4 CompressBlock) and the frame
// assembly of compressBackend (blosc.go:336-373).  The reference can only decode ONE LZ4
// block per frame, so a frame is still one block on the wire, but it is produced in
// parallel:
//
//   encode   one warp per 64 KiB SEGMENT of a frame (warps take (frame, segment) items from an
//            atomic ticket: byte planes of very different compressibility would otherwise leave
//            most warps of a CTA idle).  A segment references only itself and the last kWarmBytes (4.5 KiB)
//            of its predecessor (entered into the hash table first, so runs and periodic patterns
//            continue across the boundary), so positions fit a table entry and the reference
//            compressor's adaptive skip restarts at every segment (after a byte shuffle: at every
//            byte plane).  The warp writes the segment's sequences in final wire format into a
//            scratch slot, EXCEPT the token / literal run of its first sequence and its trailing
//            literals, which depend on the neighbouring segments.
//   finalize one thread per frame walks the segment summaries: trailing literals of segment
//            k are carried into the first sequence of the next segment that has a match
//            (they are contiguous in the input), which fixes every segment's place in the
//            block, the block size c and the memcpy decision (c >= n, blosc.go:342-345).  It
//            also writes the optional side-car decode index (one entry per segment).
//   pack     (after the offsets scan, K5) one CTA per segment writes the merged first token,
//            its literals and the segment body at the frame's packed position; memcpy frames
//            copy the raw bytes instead.  Segment 0 writes the 16-byte header.
//
// Match finder: a 2^HL-entry shared-memory hash table per warp whose 32-bit entries hold a
// 17-bit position and 15 check bits of the multiplicative 4-byte hash (a candidate is only
// fetched from global memory when the check bits agree).  Compressible regions are parsed by
// STRIPS (one 61-byte strip per lane, see "dense parse" below); incompressible regions by the
// reference's skip schedule (32 probes per step at a growing stride, MATCH.ANY for repeats
// inside the step, first hit wins, 32-lane-wide backward / forward extension).
#pragma once
#include "common.cuh"

#ifndef B2B_TRACE
#define B2B_TRACE(...) ((void)0)      // tests/emu only: event hook of the CPU shim, nothing in the CUDA build
#endif
#ifndef B2B_STAT
#define B2B_STAT(i, v) ((void)0)      // tests/emu only: counters of the CPU shim
#endif

namespace b2b {

constexpr uint32_t kSegBytes = 65536;
constexpr uint32_t kSegSlot = 65840;   // align16(65536 + 65536/255 + 32): worst case of one segment
constexpr int kEncWarps = 4;           // segments (warps) per CTA
constexpr int kEncThreads = kEncWarps * 32;
constexpr int kHashLogDefault = 10;    // 2^10 x u32 = 4 KiB (+ 3 KiB of match lists) per warp: 32 resident warps per SM
constexpr uint32_t kWarmBytes = 4608;  // tail of the previous segment pre-loaded into the hash table:
constexpr uint32_t kWarmDense = 512;   //   its last kWarmDense bytes at every position, the rest at every 4th

struct SegMeta {
    uint32_t first_ll;   // literals before the first match (whole segment if there is none)
    uint32_t body_len;   // bytes in the scratch slot
    uint32_t trail_ll;   // literals after the last match
    uint32_t info;       // bit 8: has a match; bits 0..3: match-length nibble of the first token
};
struct SegPlace {
    uint32_t out_off;    // payload offset of this segment's (merged) first token
    uint32_t lit_total;  // literals of that token: carried ones + first_ll
    uint32_t trail_dst;  // payload offset of the literals this segment hands on: its trailing literals, or (no match) all of
                         // it.  They lie inside the literal run of the next token, whichever segment holds that
    uint32_t pad;
};

__host__ __device__ __forceinline__ uint32_t seg_count(uint32_t n) { return (n + kSegBytes - 1) / kSegBytes; }
// scratch bytes of one frame's slots: exact LZ4 bound for single-segment frames
__host__ __device__ __forceinline__ uint64_t frame_slot_bytes(uint32_t n) {
    const uint32_t s = seg_count(n);
    if (s > 1) return (uint64_t)s * kSegSlot;
    return ((uint64_t)n + n / 255u + 32ull + 15ull) & ~15ull;
}
__host__ __device__ __forceinline__ uint32_t len_ext_bytes(uint32_t v) { return v >= 15 ? (v - 15) / 255u + 1u : 0u; }

// Unaligned 32-bit load as two aligned words and a funnel shift (no branch).  Callers only use
// it at least 8 bytes before the end of the frame (all match limits are >= 14 bytes short of
// it), so the second word is always inside the buffer.
__device__ __forceinline__ uint32_t enc_load32u(const uint8_t *p) {
    const uint32_t r = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t *q = reinterpret_cast<const uint32_t *>((uintptr_t)p - r);
    return __funnelshift_r(q[0], q[1], 8u * r);
}

// The input of a segment seen as aligned 32-bit words: w = org rounded down to 4 bytes, sh = what was
// rounded off.  Byte position p of the segment is byte (p + sh) of w; all index arithmetic is 32-bit
// and the loads stay in the global address space (LDG with a scaled index).
struct WordView {
    const uint32_t *w;
    uint32_t sh;
};
__device__ __forceinline__ WordView word_view(const uint8_t *org) {
    WordView v;
    v.sh = (uint32_t)((uintptr_t)org & 3u);
    v.w = reinterpret_cast<const uint32_t *>(org - v.sh);
    return v;
}
// unaligned 32-bit load at byte position p (same safety rule as enc_load32u)
__device__ __forceinline__ uint32_t wv_load32(const WordView &v, uint32_t p) {
    const uint32_t a = p + v.sh;
    const uint32_t *q = v.w + (a >> 2);
    return __funnelshift_r(q[0], q[1], a << 3);
}
// 8 bytes at position p as two words (three aligned loads)
__device__ __forceinline__ void wv_load64(const WordView &v, uint32_t p, uint32_t &lo, uint32_t &hi) {
    const uint32_t a = p + v.sh;
    const uint32_t *q = v.w + (a >> 2);
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    lo = __funnelshift_r(w0, w1, a << 3);
    hi = __funnelshift_r(w1, w2, a << 3);
}

// 8 bytes at position p and the byte after them (still three aligned loads)
__device__ __forceinline__ void wv_load72(const WordView &v, uint32_t p, uint32_t &lo, uint32_t &hi, uint32_t &b8) {
    const uint32_t a = p + v.sh;
    const uint32_t *q = v.w + (a >> 2);
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    lo = __funnelshift_r(w0, w1, a << 3);
    hi = __funnelshift_r(w1, w2, a << 3);
    b8 = (w2 >> ((a & 3u) << 3)) & 0xFFu;
}

// the same in two halves, so that the loads can be issued one turn before their use
__device__ __forceinline__ void wv_issue64(const WordView &v, uint32_t p, uint32_t &w0, uint32_t &w1, uint32_t &w2) {
    const uint32_t *q = v.w + ((p + v.sh) >> 2);
    w0 = q[0]; w1 = q[1]; w2 = q[2];
}
__device__ __forceinline__ void wv_finish64(const WordView &v, uint32_t p, uint32_t w0, uint32_t w1, uint32_t w2,
                                            uint32_t &lo, uint32_t &hi) {
    const uint32_t a = (p + v.sh) << 3;
    lo = __funnelshift_r(w0, w1, a);
    hi = __funnelshift_r(w1, w2, a);
}

// Hash of the window at a position: lo = its first 4 bytes, hi = the bytes after them.  HB = bytes hashed:
//   4  typed arrays behind a byte shuffle (short matches count; the default)
//   5  unshuffled input: a 4-byte context is too ambiguous in text and in low-entropy integers (the most recent
//      occurrence is rarely the longest one); the reference compressor hashes 6 bytes
//   6  the reference's own width
// phase: position inside a bit-shuffle group (0 otherwise): after the bit shuffle byte k of group g is only
// comparable with byte k of an earlier group, so the place is part of the key (candidates at offsets 8T * j).
template <int HB>
__device__ __forceinline__ uint32_t enc_hash(uint32_t lo, uint32_t hi, uint32_t phase) {
    if constexpr (HB == 4) return (lo ^ (phase * 0x9E3779B1u)) * 2654435761u;
    else if constexpr (HB == 5) return (lo ^ (((hi & 0xFFu) | (phase << 8)) * 0x9E3779B1u)) * 2654435761u;
    else return (lo ^ (((hi & 0xFFFFu) | (phase << 16)) * 0x9E3779B1u)) * 2654435761u;
}

// writes a length extension (value already reduced by 15) at out, returns bytes written
__device__ __forceinline__ uint32_t warp_put_len_ext(uint8_t *out, uint32_t v, int lane) {
    const uint32_t full = v / 255u, last = v - full * 255u;
    for (uint32_t i = lane; i < full; i += kWarp) out[i] = 255;
    if (lane == 0) out[full] = (uint8_t)last;
    return full + 1;
}

// ---- emission helpers ----------------------------------------------------------------------
// A lane's output stream leaves as aligned 32-bit words: bytes collect in a 64-bit accumulator and
// are stored four at a time once the write position is 4-byte aligned; the unaligned head and the
// last < 4 bytes go out as single bytes (their neighbours belong to other lanes).  32 lanes writing
// scattered single bytes cost one LSU wavefront per byte; this cuts them by about four.
struct LaneWriter {
    uint8_t *p;        // where the first byte still held in acc goes
    uint32_t acc;
    uint32_t nb;       // bytes held in acc (< 4)
    __device__ __forceinline__ void begin(uint8_t *at) { p = at; acc = 0; nb = 0; }
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {       // n <= 4 bytes, v clean above them
        if ((uintptr_t)p & 3u) {                                        // unaligned head: single bytes
            for (uint32_t i = 0; i < n; i++) {
                if ((uintptr_t)p & 3u) { *p++ = (uint8_t)v; v >>= 8; }
                else { acc |= (v & 0xFFu) << (8u * nb); nb++; v >>= 8; }
            }
            return;
        }
        const uint32_t s = 8u * nb;
        const uint32_t lo = acc | (v << s), hi = __funnelshift_l(v, 0u, s);   // hi = v >> (32 - s), 0 for s = 0
        nb += n;
        if (nb >= 4) { *reinterpret_cast<uint32_t *>(p) = lo; p += 4; acc = hi; nb -= 4; }
        else acc = lo;
    }
    __device__ __forceinline__ void flush() { while (nb) { *p++ = (uint8_t)acc; acc >>= 8; nb--; } }
};

// Byte-loop copy for the emission paths (two loads in flight; the ranges never overlap).  The 16-byte realigning
// path of warp_copy would set the register footprint of the whole kernel for the sake of the odd long literal run
// (72 -> 69 registers, encode 20.6 -> 19.7 ms per 8 GiB on C3, +12 % on bit-shuffled float64).
__device__ __forceinline__ void warp_copy_lean(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    for (uint32_t i = lane; i < n; i += 2 * kWarp) {
        const bool p1 = i + kWarp < n;
        const uint8_t v0 = src[i];
        uint8_t v1 = 0;
        if (p1) v1 = src[i + kWarp];
        dst[i] = v0;
        if (p1) dst[i + kWarp] = v1;
    }
}

struct EncState {
    uint8_t *body;
    uint32_t op;
    bool have_first;
    SegMeta m;
};

// One sequence written by the whole warp: ll literals starting at lit, then the match (offset,
// mlc = match length - 4).  The first sequence of a segment only leaves its offset / match
// extension in the body: its token and literals are written by the pack pass.
__device__ __forceinline__ void emit_coop(EncState &st, const uint8_t *lit, uint32_t ll, uint32_t offset,
                                          uint32_t mlc, int lane) {
    uint8_t *body = st.body;
    uint32_t op = st.op;
    if (!st.have_first) {
        st.have_first = true;
        st.m.first_ll = ll;
        st.m.info = 0x100u | (mlc < 15 ? mlc : 15u);
        if (lane < 2) body[op + lane] = (uint8_t)(offset >> (8 * lane));
        op += 2;
    } else if (ll <= 28) {
        // short literal run: token, literals and offset leave in ONE predicated byte store
        uint32_t v = ((ll < 15 ? ll : 15u) << 4) | (mlc < 15 ? mlc : 15u);
        const uint32_t ext = ll >= 15 ? 1u : 0u;                 // 15..28 literals: one extension byte
        if ((uint32_t)lane == 1 && ext) v = ll - 15;
        if ((uint32_t)lane > ext && (uint32_t)lane <= ext + ll) v = lit[lane - 1 - ext];
        if ((uint32_t)lane == ext + ll + 1) v = offset;
        if ((uint32_t)lane == ext + ll + 2) v = offset >> 8;
        if ((uint32_t)lane < ext + ll + 3) body[op + lane] = (uint8_t)v;
        op += ext + ll + 3;
    } else {
        const uint32_t tok_pos = op++;
        op += warp_put_len_ext(body + op, ll - 15, lane);
        warp_copy_lean(body + op, lit, ll, lane);
        op += ll;
        if (lane == 0) body[tok_pos] = (uint8_t)(0xF0u | (mlc < 15 ? mlc : 15u));
        if (lane < 2) body[op + lane] = (uint8_t)(offset >> (8 * lane));
        op += 2;
    }
    if (mlc >= 15) op += warp_put_len_ext(body + op, mlc - 15, lane);
    st.op = op;
}

// 32-lane forward extension: first position >= from where org[pos] != org[pos - offset], capped at mlimit
__device__ __forceinline__ uint32_t extend_coop(const WordView &in, uint32_t from, uint32_t offset,
                                                uint32_t mlimit, int lane) {
    uint32_t mend = from;
    for (;;) {
        const uint32_t a = mend + 4u * lane;
        uint32_t x = 0xFFFFFFFFu;
        if (a < mlimit) {
            x = wv_load32(in, a) ^ wv_load32(in, a - offset);
            const uint32_t avail = mlimit - a;
            if (avail < 4) x |= 0xFFFFFFFFu << (8u * avail);
        }
        const uint32_t diff = __ballot_sync(0xffffffffu, x != 0);
        if (diff == 0) { mend += 128; continue; }
        const int fl = __ffs(diff) - 1;
        const uint32_t xf = __shfl_sync(0xffffffffu, x, fl);
        return mend + 4u * fl + ((uint32_t)(__ffs((int)xf) - 1) >> 3);
    }
}

// ---- dense parse: every lane owns a strip of the input -------------------------------------
// In compressible regions one step covers 32 strips of kStrip bytes.  Each lane walks ITS strip
// like a scalar LZ4 compressor (probe / insert at every position, follow a hit, jump behind it);
// the hash table is shared, so lanes find each other's positions.  The loop is a two-state machine
// (probing | following a match, 4 bytes per turn) so that lanes in different phases still share
// most instructions.  A lane follows its own match at most kStripCap bytes beyond its strip; what
// is still matching there ("open") is finished 32 lanes wide.  Afterwards matches that overlap a
// match of an earlier strip are dropped or trimmed (prefix maximum of the match ends over the
// lanes), literal runs are measured from the previous surviving match, a prefix sum of the encoded
// sizes places every lane's sequences, and each lane writes its own bytes.
// kStrip is deliberately not a divisor of the power-of-two periods typed arrays have after a
// shuffle: with period P, the nearest earlier occurrence of a position lies in a lower lane's
// strip at an EARLIER turn (P mod kStrip turns before), so it is already in the table.
constexpr uint32_t kStrip = 61;
constexpr uint32_t kStripCap = 61;       // own match followed this far past the strip end
constexpr uint32_t kListMax = 14;        // matches a lane records per strip (a 15th would need 61 bytes of
                                         // back-to-back 4-byte matches; the lane then stops probing).  14 rather
                                         // than 16 keeps a CTA at 26.5 KiB of shared memory: 8 CTAs per SM, not 7
constexpr uint32_t kDeadTrialBytes = 20480;   // start of a bit-shuffled segment that runs without the noise-place rule and audits it
                                              // (a cold table sees a period only after its first repeat)
constexpr uint32_t kDeadTrialWarm = 6144;     // the same for a segment that starts from its predecessor's warm-up window (periods up to 4 KiB are in the table)
constexpr uint32_t kDeadHits = 8;        // matches inside the candidate noise places, in one audited step, that switch the rule off
constexpr uint32_t kLaneLitEmit = 32;    // literal runs up to here are written by the owning lane

struct LaneLists {
    uint32_t a[kListMax][32];            // start (17 bits) | length << 17 (the last match of a lane keeps its end in a register)
    uint16_t off[kListMax][32];
};

// One segment.  org: first byte of the warm-up window (W bytes before the segment, 0 for the
// first one), L the segment length, tail: bytes of the frame after it.  Positions below are
// relative to org.  Limits follow the reference compressor relative to the END OF THE FRAME:
// no match starts or is extended inside the frame's last 14 bytes (pierrec mfLimit), which
// also keeps the reference's memcpy decision for short frames.
//
// Dense steps (stride 1) parse the WHOLE 32-position window at once: every hit lane follows its
// own match (bounded), the greedy chain "first hit at or after the end of the previous match"
// is resolved with shuffles, and the selected sequences are written by their own lanes at
// offsets from a warp prefix sum.  Only the first sequence of a step (its literals may reach
// far back) and a match longer than the per-lane bound go through the 32-lane-wide path.
template <int HL, int HB, bool PH>
__device__ __forceinline__ SegMeta warp_encode_segment(const uint8_t *__restrict__ org, uint32_t W,
                                                       uint32_t L, uint64_t tail,
                                                       uint8_t *__restrict__ body, uint32_t *table,
                                                       LaneLists *lists, int lane, uint32_t dense_lits,
                                                       uint32_t strip_full, uint32_t strip_cap, bool cold_start,
                                                       uint32_t writer_min_lits, uint32_t ph0_, uint32_t phmask_) {
    // PH: the input is bit-shuffled and the frame is long enough for it to matter: ph0 = place of org inside a
    // group, phmask = group size - 1 (a power of two); both zero (and folded away) otherwise
    const uint32_t ph0 = PH ? ph0_ : 0u, phmask = PH ? phmask_ : 0u;
    EncState st;
    st.body = body; st.op = 0; st.have_first = false;
    st.m.first_ll = L; st.m.body_len = 0; st.m.trail_ll = L; st.m.info = 0;
    const int64_t fl = (int64_t)L + (int64_t)tail - 14;          // frame limit in segment coordinates
    uint32_t mfl = L >= 4 ? L - 3 : 0;                           // a match of 4 bytes must fit the segment
    if (fl < (int64_t)mfl) mfl = fl > 0 ? (uint32_t)fl : 0u;
    uint32_t mlimit = L;
    if (fl < (int64_t)L) mlimit = fl > 0 ? (uint32_t)fl : 0u;
    if (mfl == 0) return st.m;
    mfl += W; mlimit += W;

    const WordView in = word_view(org);
    for (uint32_t i = lane; i < (1u << HL); i += kWarp) table[i] = 0;
    __syncwarp();
    // warm-up: enter the tail of the previous segment (ascending, so the nearest position wins).  The
    // older part goes in at every 4th position: a repeat of 8 bytes or more still meets an entered
    // position within 3 bytes and backward extension recovers the start, so periods up to 4 KiB
    // continue across a segment boundary at a quarter of the cost.
    const uint32_t Wsparse = W > kWarmDense ? W - kWarmDense : 0u;
    for (uint32_t q0 = 0; q0 < Wsparse; q0 += 16u * kWarp) {        // four loads in flight per lane
        uint32_t hv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t q = q0 + 4u * ((uint32_t)j * kWarp + (uint32_t)lane);
            uint32_t wlo = 0, whi = 0;
            if (q < Wsparse) wv_load64(in, q, wlo, whi);
            hv[j] = enc_hash<HB>(wlo, whi, (ph0 + q) & phmask);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t q = q0 + 4u * ((uint32_t)j * kWarp + (uint32_t)lane);
            if (q < Wsparse) table[hv[j] >> (32 - HL)] = (((hv[j] >> (17 - HL)) & 0x7FFFu) << 17) | q;
        }
    }
    __syncwarp();
    for (uint32_t q0 = Wsparse; q0 < W; q0 += 4u * kWarp) {
        uint32_t hv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t q = q0 + (uint32_t)j * kWarp + (uint32_t)lane;
            uint32_t wlo = 0, whi = 0;
            if (q + 3 < W) wv_load64(in, q, wlo, whi);
            hv[j] = enc_hash<HB>(wlo, whi, (ph0 + q) & phmask);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t q = q0 + (uint32_t)j * kWarp + (uint32_t)lane;
            if (q + 3 < W) table[hv[j] >> (32 - HL)] = (((hv[j] >> (17 - HL)) & 0x7FFFu) << 17) | q;
        }
    }
    __syncwarp();
    // Bit-shuffled typed arrays: a group of 8 * typesize bytes holds, place by place, one bit of eight elements -- the low
    // mantissa bits first (noise), the high bits, sign and exponent last (nearly constant from group to group).  Looking
    // for matches in the noise places costs most of the encoder's time there and the few it finds (4 bytes, far back) push
    // the useful positions out of the 1 024-entry table.  So the warp measures, on up to 64 pairs of neighbouring groups
    // of its segment, at which place the group stops being noise (first place whose byte repeats from group to group in
    // at least half of the pairs) and probes / enters positions only from one byte past it.  float64 (C4): 48 of 64 places
    // are noise.  Unaudited: size 1.030 -> 1.019 of the oracle's, encode 9.1 -> 5.6 ms per 2 GiB, compress 203 -> 301 GB/s;
    // with the audit below (which the bit planes of a ramp need): 1.026, 6.8 ms, 271 GB/s in the bench's 16 GiB run.
    // (Groups that start with compressible places, e.g. int16 counters, have no noise prefix and are left alone.)
    uint32_t dead = 0, dead_c = 0;
    if (PH && phmask) {
        const uint32_t G = phmask + 1u;
        const uint32_t q0 = W + ((G - ((ph0 + W) & phmask)) & phmask);   // first group that starts inside the segment
        const uint32_t avail = W + L > q0 + G ? (W + L - q0) / G - 1u : 0u;   // pairs (g, g + 1) wholly inside it
        const uint32_t np = avail < 64u ? avail : 64u;
        if (np >= 16u) {
            uint32_t p = 0;
            for (; p < G; p++) {
                uint32_t c = 0;
#pragma unroll
                for (uint32_t u = 0; u < 2; u++) {
                    const uint32_t g = (uint32_t)lane + 32u * u;
                    if (g < np) c += org[q0 + g * G + p] == org[q0 + (g + 1u) * G + p] ? 1u : 0u;
                }
                c = __reduce_add_sync(0xffffffffu, c);
                if (2u * c >= np) break;
            }
            if (p >= 8u && p < G) dead_c = p + 4u;
        }
    }
    // ... but "does not repeat from group to group" is not yet "noise": the bit planes of a RAMP do not repeat that way
    // either and still compress 40:1 through matches that start in those places.  So the rule is AUDITED: the dense steps
    // of the segment's first 20 KiB (6 KiB when the table starts warm) and every eighth one after them run without it and count the matches that start in
    // the candidate places; more than kDeadHits in a step switch the rule off for the rest of the segment.
    uint32_t dstep = 0, trial_hits = 0;
    bool audit = false;
    uint32_t anchor = W, si = W;
    uint32_t rep = 0;                                            // this lane's last match offset
    uint32_t ramp = cold_start ? 0u : 5u;                        // dense steps taken so far (5: no ramp)
    while (si < mfl) {
        // adaptive skip of the reference compressor: about 3 probes per 4 + lits/128 bytes
        const uint32_t lits = si - anchor;
        const uint32_t stride = lits < dense_lits ? 1u : (4u + (lits >> 7)) / 3u;
        if (stride == 1) {
            // ------------------------------------------------------------ dense step: strips
            // The first dense steps of a segment run on 1, 2, 4, 8, 16 lanes: the table is cold, and inside
            // ONE step a lane cannot see what a lower lane will only reach in a later turn, so a period
            // that first repeats inside a full step is mostly missed there (a 1024-byte period lost 79 %
            // of its first 1952 bytes).  With the step size doubling, an earlier occurrence of a position
            // lies in an earlier step (or in its own strip), which the table already holds.
            // Only the first segment of a frame pays for this (small frames live there entirely; a later
            // segment can lose at most the first repeat inside one step, about a kilobyte).
            const uint32_t strip = strip_full;
            const uint32_t nlanes = ramp < 5 ? (1u << ramp) : 32u;
            ramp++;
            if (PH) {
                audit = dead_c != 0 && (si - W < (cold_start ? kDeadTrialBytes : kDeadTrialWarm) || (dstep & 7u) == 0u);
                dead = audit ? 0u : dead_c;
                dstep++;
            }
            uint32_t pos = si + (uint32_t)lane * strip;
            if (pos > mfl || (uint32_t)lane >= nlanes) pos = mfl;
            const uint32_t send = pos + strip < mfl ? pos + strip : mfl;
            const uint32_t cap = send + strip_cap < mlimit ? send + strip_cap : mlimit;
            uint32_t cnt = 0, e = 0, moff = 0, mst = 0, my_last = 0;
            bool ext = false, my_open = false;
            // the 8 + 8 bytes the next extension turn compares, loaded a turn ahead (the candidate is
            // typically kilobytes back: an L2 round trip that would otherwise open every turn).  Doing
            // the same for the probe window and the previous-offset window was measured and did not pay.
            uint32_t xa0 = 0, xa1 = 0, xa2 = 0, xb0 = 0, xb1 = 0, xb2 = 0;
            {   // the next step's strips: bring their lines into L2 while this step is parsed
                const uint32_t a = si + nlanes * strip + 64u * (uint32_t)lane;
#ifdef __CUDA_ARCH__
                if (a < mlimit) asm volatile("prefetch.global.L2 [%0];" ::"l"(org + a));
#else
                (void)a;
#endif
            }
            while (__any_sync(0xffffffffu, ext || pos < send)) {
                if (!ext) {
                    if (PH && dead && pos < send) {
                        // bit-shuffled input: nothing is looked for (or entered) where the group is noise, see `dead` above
                        const uint32_t phs = (ph0 + pos) & phmask;
                        if (phs + 3 < dead) { pos += dead - 3 - phs; if (pos > send) pos = send; }
                    }
                    if (pos < send) {
                        // ---- four consecutive positions per turn: one 12-byte window, four probes in flight
                        uint32_t v0, v1;
                        uint32_t v2 = 0;
                        if constexpr (HB == 6) wv_load72(in, pos, v0, v1, v2);
                        else wv_load64(in, pos, v0, v1);
                        uint32_t seq[4], ent[4], chk[4];
                        seq[0] = v0; seq[1] = __funnelshift_r(v0, v1, 8);
                        seq[2] = __funnelshift_r(v0, v1, 16); seq[3] = __funnelshift_r(v0, v1, 24);
                        const uint32_t nv = send - pos < 4u ? send - pos : 4u;
                        uint32_t hh[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint32_t hv = enc_hash<HB>(seq[k], __funnelshift_r(v1, v2, 8 * k), (ph0 + pos + k) & phmask);
                            hh[k] = hv >> (32 - HL); chk[k] = (hv >> (17 - HL)) & 0x7FFFu;
                            ent[k] = table[hh[k]];
                        }
                        // the lane's previous offset is tried as well (constant strides, periodic data).  Its
                        // window is typically kilobytes back (an L2 hit): the loads go out here and are only
                        // looked at after the table candidates and the inserts
                        uint32_t r0 = 0, r1 = 0, r2 = 0;
                        const bool rep_ok = rep != 0 && rep <= pos;
                        if (rep_ok) wv_issue64(in, pos - rep, r0, r1, r2);
                        // first position with a table candidate.  An entry whose check bits agree is taken
                        // unverified: the first turn of the extension compares the bytes anyway.
                        int pick = -1;
                        uint32_t pc = 0;
                        bool sure = false;
#pragma unroll
                        for (int k = 3; k >= 0; k--) {
                            // c is the candidate position if the check bits agree, else >= 2^17 (> any position)
                            const uint32_t c = ent[k] ^ (chk[k] << 17), p = pos + k;
                            if ((uint32_t)k < nv && p - c - 1u < 65535u) { pick = k; pc = c; }   // c < p, p - c <= 65535
                        }
                        // positions up to the chosen one are recorded (like the scalar compressor, which
                        // does not enter the positions it jumps over)
                        const uint32_t upto = pick < 0 ? nv : (uint32_t)pick + 1u;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if ((uint32_t)k < upto) table[hh[k]] = (chk[k] << 17) | (pos + k);
                        // the previous offset wins at the same or an earlier position (fewer, longer sequences)
                        if (rep_ok) {
                            uint32_t y0, y1;
                            wv_finish64(in, pos - rep, r0, r1, r2, y0, y1);
                            const uint32_t lim = pick < 0 ? nv : (uint32_t)pick + 1u;
                            if (lim > 3 && __funnelshift_r(y0, y1, 24) == seq[3]) { pick = 3; sure = true; }
                            if (lim > 2 && __funnelshift_r(y0, y1, 16) == seq[2]) { pick = 2; sure = true; }
                            if (lim > 1 && __funnelshift_r(y0, y1, 8) == seq[1]) { pick = 1; sure = true; }
                            if (y0 == seq[0]) { pick = 0; sure = true; }
                            if (sure) pc = pos + (uint32_t)pick - rep;
                        }
                        if (pick < 0) {
                            // repeats inside the group (runs, periods 1..3): nearest earlier position
                            if (nv > 3 && (seq[3] == seq[0])) { pick = 3; pc = pos; }
                            if (nv > 3 && (seq[3] == seq[1])) { pick = 3; pc = pos + 1; }
                            if (nv > 3 && (seq[3] == seq[2])) { pick = 3; pc = pos + 2; }
                            if (nv > 2 && (seq[2] == seq[0])) { pick = 2; pc = pos; }
                            if (nv > 2 && (seq[2] == seq[1])) { pick = 2; pc = pos + 1; }
                            if (nv > 1 && (seq[1] == seq[0])) { pick = 1; pc = pos; }
                            sure = true;
                        }
                        B2B_TRACE(1, ph0 + pos, nv, pick, pick >= 0 ? pos + pick - pc : 0, sure);
                        if (pick >= 0) {
                            ext = true; mst = pos + pick; moff = mst - pc; e = sure ? mst + 4 : mst;
                            if (e + 8 <= cap) { wv_issue64(in, e, xa0, xa1, xa2); wv_issue64(in, e - moff, xb0, xb1, xb2); }
                        } else pos += nv;
                    }
                } else {
                    bool done = true, cancel = false;
                    if (e + 8 <= cap) {
                        uint32_t a0, a1, b0, b1;
                        wv_finish64(in, e, xa0, xa1, xa2, a0, a1);
                        wv_finish64(in, e - moff, xb0, xb1, xb2, b0, b1);
                        const uint32_t x0 = a0 ^ b0, x1 = a1 ^ b1;
                        if (x0) { cancel = e == mst; e += (uint32_t)(__ffs((int)x0) - 1) >> 3; }
                        else if (x1) e += 4u + ((uint32_t)(__ffs((int)x1) - 1) >> 3);
                        else {
                            e += 8; done = false;
                            if (e + 8 <= cap) { wv_issue64(in, e, xa0, xa1, xa2); wv_issue64(in, e - moff, xb0, xb1, xb2); }
                        }
                    } else {
                        if (e == mst) {
                            if (wv_load32(in, e) == wv_load32(in, e - moff)) e += 4; else cancel = true;
                        }
                        if (!cancel) while (e < cap && org[e] == org[e - moff]) e++;
                    }
                    if (cancel) { ext = false; pos = mst + 1; done = false; }
                    if (done) {
                        my_open = e >= cap && cap < mlimit;
                        lists->a[cnt][lane] = mst | ((e - mst) << 17);
                        lists->off[cnt][lane] = (uint16_t)moff;
                        cnt++;
                        pos = cnt < kListMax ? e : send; rep = moff; my_last = e; ext = false;
                    }
                }
            }
            const uint32_t region_end = si + nlanes * strip < mfl ? si + nlanes * strip : mfl;
            const uint32_t any_match = __ballot_sync(0xffffffffu, cnt != 0);
            if (any_match == 0) { si = region_end; continue; }
            // ---- open matches are finished 32 lanes wide, in stream order; one that an earlier
            // finished match already covers is left alone (it is dropped below)
            {
                uint32_t opens = __ballot_sync(0xffffffffu, my_open);
                uint32_t erun = 0;
                while (opens) {
                    const int j = __ffs(opens) - 1;
                    const uint32_t ce = __shfl_sync(0xffffffffu, my_last, j);
                    const uint32_t oj = __shfl_sync(0xffffffffu, moff, j);
                    if (ce > erun) {
                        erun = extend_coop(in, ce, oj, mlimit, lane);
                        if (lane == j) my_last = erun;
                    }
                    opens &= opens - 1;
                    opens &= __ballot_sync(0xffffffffu, my_last > erun);
                }
            }
            // ---- drop / trim against the matches of earlier strips
            uint32_t ein;                                      // everything before ein is taken
            {
                uint32_t incl = my_last;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d && t > incl) incl = t;
                }
                ein = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0 || ein < anchor) ein = anchor;
            }
            uint32_t kf = 0, f_ms = 0;                         // first surviving match, its (trimmed) start
            while (kf < cnt) {
                const uint32_t a = lists->a[kf][lane];
                const uint32_t ms = a & 0x1FFFFu;
                const uint32_t me = kf + 1 == cnt ? my_last : ms + (a >> 17);
                if (me > ein && (ms >= ein || me - ein >= 4)) { f_ms = ms > ein ? ms : ein; break; }
                kf++;
            }
            const bool kept = kf < cnt;
            uint32_t prev;                                     // end of the previous surviving match
            uint32_t last_kept_end;
            {
                uint32_t incl = kept ? my_last : 0u;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d && t > incl) incl = t;
                }
                prev = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0 || prev < anchor) prev = anchor;
                last_kept_end = __shfl_sync(0xffffffffu, incl, 31);
            }
            const uint32_t keptm = __ballot_sync(0xffffffffu, kept);
            if (keptm == 0) { si = region_end; continue; }
            // the first sequence of the segment leaves only its offset / match extension in the body
            const bool seg_first = !st.have_first && lane == __ffs(keptm) - 1;
            // ---- backward extension over the pending literals (down to the previous surviving match,
            // possibly in another strip) and the encoded size of this lane's sequences
            uint32_t size = 0, lit_sum = 0, nseq = 0;          // lit_sum: literals the lanes write themselves
            {
                uint32_t pe = prev;
                for (uint32_t k = kf; k < cnt; k++) {
                    const uint32_t a = lists->a[k][lane];
                    uint32_t ms = k == kf ? f_ms : (a & 0x1FFFFu);
                    const uint32_t me = k + 1 == cnt ? my_last : (a & 0x1FFFFu) + (a >> 17);
                    uint32_t mc = ms - lists->off[k][lane];
                    while (ms > pe && mc > 0 && org[ms - 1] == org[mc - 1]) { ms--; mc--; }
                    if (PH && audit && ((ph0 + ms) & phmask) + 7u < dead_c) trial_hits++;   // (its four bytes lie in candidate places)
                    const uint32_t len = me - ms;
                    lists->a[k][lane] = ms | ((len < 0x7FFFu ? len : 0x7FFFu) << 17);
                    const uint32_t ll = ms - pe, mlc = len - 4;
                    if (seg_first && k == kf) size += 2u + len_ext_bytes(mlc);
                    else size += 1u + len_ext_bytes(ll) + ll + 2u + len_ext_bytes(mlc);
                    if (ll <= kLaneLitEmit) lit_sum += ll;
                    nseq++;
                    pe = me;
                }
            }
            uint32_t incl = size;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const bool use_writer = __reduce_add_sync(0xffffffffu, lit_sum) >= writer_min_lits * __reduce_add_sync(0xffffffffu, nseq);
            // ---- every lane writes its own sequences; long literal runs are left to the warp
            uint32_t g_dst0 = 0, g_src0 = 0, g_len0 = 0, g_dst1 = 0, g_src1 = 0, g_len1 = 0;
            // Two ways to write them: steps whose lanes write two or more literals per sequence on average
            // (low-entropy data: runs of 8..30 literals between short matches) go through the aligned-word
            // writer, which cuts the scattered single-byte stores by four (C5: 39 -> 30 ms per 8 GiB);
            // steps of mostly token + offset, and steps whose literal runs are long enough to be copied
            // 32 lanes wide anyway (C4), are cheaper byte by byte.
            if (use_writer)
            {
                LaneWriter w;
                w.begin(body + st.op + (incl - size));
                uint32_t pe = prev, ng = 0;
                for (uint32_t k = kf; k < cnt; k++) {
                    const uint32_t a = lists->a[k][lane];
                    const uint32_t ms = a & 0x1FFFFu;
                    const uint32_t me = k + 1 == cnt ? my_last : ms + (a >> 17);
                    const uint32_t offset = lists->off[k][lane];
                    const uint32_t ll = ms - pe, mlc = me - ms - 4;
                    if (seg_first && k == kf) {
                        st.m.first_ll = ll;
                        st.m.info = 0x100u | (mlc < 15 ? mlc : 15u);
                    } else {
                        w.put(((ll < 15 ? ll : 15u) << 4) | (mlc < 15 ? mlc : 15u), 1);
                        if (ll >= 15) {
                            uint32_t v = ll - 15;
                            while (v >= 255) { w.put(255, 1); v -= 255; }
                            w.put(v, 1);
                        }
                        if (ll <= kLaneLitEmit) {
                            uint32_t i = 0;
                            for (; i + 4 <= ll; i += 4) w.put(wv_load32(in, pe + i), 4);
                            if (i < ll) w.put(wv_load32(in, pe + i) & (0xFFFFFFFFu >> (8u * (4u - (ll - i)))), ll - i);
                        } else {
                            w.flush();                                  // the run is copied 32 lanes wide below
                            if (ng == 0) { g_dst0 = (uint32_t)(w.p - body); g_src0 = pe; g_len0 = ll; }
                            else { g_dst1 = (uint32_t)(w.p - body); g_src1 = pe; g_len1 = ll; }
                            ng++;
                            w.p += ll;
                        }
                    }
                    w.put(offset, 2);
                    if (mlc >= 15) {
                        uint32_t v = mlc - 15;
                        while (v >= 255) { w.put(255, 1); v -= 255; }
                        w.put(v, 1);
                    }
                    pe = me;
                }
                w.flush();
            }
            else
            {
                uint8_t *b = body + st.op + (incl - size);
                uint32_t pe = prev, ng = 0;
                for (uint32_t k = kf; k < cnt; k++) {
                    const uint32_t a = lists->a[k][lane];
                    const uint32_t ms = a & 0x1FFFFu;
                    const uint32_t me = k + 1 == cnt ? my_last : ms + (a >> 17);
                    const uint32_t offset = lists->off[k][lane];
                    const uint32_t ll = ms - pe, mlc = me - ms - 4;
                    if (seg_first && k == kf) {
                        st.m.first_ll = ll;
                        st.m.info = 0x100u | (mlc < 15 ? mlc : 15u);
                    } else {
                        *b++ = (uint8_t)(((ll < 15 ? ll : 15u) << 4) | (mlc < 15 ? mlc : 15u));
                        if (ll >= 15) {
                            uint32_t v = ll - 15;
                            while (v >= 255) { *b++ = 255; v -= 255; }
                            *b++ = (uint8_t)v;
                        }
                        if (ll <= kLaneLitEmit) {
                            const uint8_t *lit = org + pe;
                            uint32_t i = 0;
                            for (; i + 4 <= ll; i += 4) {
                                const uint8_t v0 = lit[i], v1 = lit[i + 1], v2 = lit[i + 2], v3 = lit[i + 3];
                                b[i] = v0; b[i + 1] = v1; b[i + 2] = v2; b[i + 3] = v3;
                            }
                            for (; i < ll; i++) b[i] = lit[i];
                        } else {
                            if (ng == 0) { g_dst0 = (uint32_t)(b - body); g_src0 = pe; g_len0 = ll; }
                            else { g_dst1 = (uint32_t)(b - body); g_src1 = pe; g_len1 = ll; }
                            ng++;
                        }
                        b += ll;
                    }
                    b[0] = (uint8_t)offset; b[1] = (uint8_t)(offset >> 8);
                    b += 2;
                    if (mlc >= 15) {
                        uint32_t v = mlc - 15;
                        while (v >= 255) { *b++ = 255; v -= 255; }
                        *b++ = (uint8_t)v;
                    }
                    pe = me;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint32_t gd_r = r ? g_dst1 : g_dst0, gs_r = r ? g_src1 : g_src0, gl_r = r ? g_len1 : g_len0;
                uint32_t gaps = __ballot_sync(0xffffffffu, gl_r != 0);
                while (gaps) {
                    const int j = __ffs(gaps) - 1;
                    gaps &= gaps - 1;
                    const uint32_t gd = __shfl_sync(0xffffffffu, gd_r, j), gs = __shfl_sync(0xffffffffu, gs_r, j);
                    const uint32_t gl = __shfl_sync(0xffffffffu, gl_r, j);
                    warp_copy_lean(body + gd, org + gs, gl, lane);
                }
            }
            if (!st.have_first) {
                // the lane that wrote the first sequence of the segment publishes its summary
                const int fl0 = __ffs(keptm) - 1;
                st.m.first_ll = __shfl_sync(0xffffffffu, st.m.first_ll, fl0);
                st.m.info = __shfl_sync(0xffffffffu, st.m.info, fl0);
                st.have_first = true;
            }
            if (PH && audit) {
                const uint32_t hits = __reduce_add_sync(0xffffffffu, trial_hits);
                trial_hits = 0;
                B2B_STAT(24, hits);
                if (hits > kDeadHits) dead_c = 0;
            }
            st.op += total;
            anchor = last_kept_end;
            si = region_end > anchor ? region_end : anchor;
            __syncwarp();
            continue;
        }
        const uint32_t p = si + (uint32_t)lane * stride;
        const bool valid = p < mfl;
        uint32_t seq = 0, seq_hi = 0;
        if (valid) wv_load64(in, p, seq, seq_hi);
        const uint32_t hv = enc_hash<HB>(seq, seq_hi, (ph0 + p) & phmask);
        const uint32_t h = hv >> (32 - HL);
        const uint32_t chk = (hv >> (17 - HL)) & 0x7FFFu;
        const uint32_t ent = table[h];
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        const uint32_t same = __match_any_sync(0xffffffffu, seq) & vmask;
        // nearest earlier lane of this step with the same 4 bytes, else the table entry
        const uint32_t lower = same & ((1u << lane) - 1u);
        uint32_t cand = 0;
        bool ok = false;
        if (valid && lower) {
            const uint32_t src_lane = 31u - (uint32_t)__clz((int)lower);
            cand = p - (uint32_t)(lane - src_lane) * stride;
            ok = true;
        } else if (valid && (ent >> 17) == chk) {
            cand = ent & 0x1FFFFu;
            ok = cand < p && p - cand < 65536u && wv_load32(in, cand) == seq;
        }
        const uint32_t hit = __ballot_sync(0xffffffffu, ok);

        // ---------------------------------------------------------------- strided probing
        const int pick = hit ? __ffs(hit) - 1 : 31;
        __syncwarp();
        // every probed position up to the chosen start is recorded
        const uint32_t upto = same & ((2u << pick) - 1u);
        if (valid && lane <= pick && (upto >> lane) == 1u) table[h] = (chk << 17) | p;
        __syncwarp();
        if (hit == 0) {
            si += 32u * stride;
            continue;
        }
        uint32_t mp = __shfl_sync(0xffffffffu, p, pick);      // match start
        uint32_t mc = __shfl_sync(0xffffffffu, cand, pick);   // its source
        const uint32_t offset = mp - mc;
        const uint32_t mend = extend_coop(in, mp + 4, offset, mlimit, lane);
        // backward extension over the pending literals (never into the previous segment)
        while (mp > anchor) {
            const uint32_t k = lane + 1;
            const bool eq = (mp >= anchor + k) && (mc >= k) && org[mp - k] == org[mc - k];
            const uint32_t neq = ~__ballot_sync(0xffffffffu, eq);
            const uint32_t back = neq ? (uint32_t)(__ffs((int)neq) - 1) : 32u;
            mp -= back; mc -= back;
            if (back < 32) break;
        }
        emit_coop(st, org + anchor, mp - anchor, offset, mend - mp - 4, lane);
        si = mend; anchor = mend; rep = offset;
    }
    st.m.body_len = st.op;
    st.m.trail_ll = W + L - anchor;
    return st.m;
}

struct EncodeArgs {
    const uint8_t *in;          // (shuffled) input, frame f at src_off[f]
    const uint64_t *src_off;
    const uint32_t *src_len;
    uint32_t nframes;
    uint32_t segs_grid;         // items per frame in the ticket space; an item strides over segments
    uint8_t *comp;              // scratch: segment s of frame f at comp_off[f] + s * kSegSlot
    const uint64_t *comp_off;
    const uint64_t *seg_base;   // index of frame f's first segment in meta / place
    SegMeta *meta;
    unsigned long long *ticket; // zero before launch
    uint32_t tune[4];           // experiment knobs (0: default)
    uint32_t independent;       // 1: no warm-up window, segments never reference each other (decode index)
    uint32_t phase_mask;        // 8 * typesize - 1 when the input went through the bit shuffle and that is a power of two
                                // (candidates are then looked for at the same place of an earlier group only), else 0
    uint32_t planes;            // typesize when the input went through the byte shuffle (plane = len / typesize), else 0
    uint64_t comp_cap, seg_cap; // bytes of `comp`, entries of `meta`: sized from the caller's total_src_bytes; a frame
                                // that does not fit (sources that overlap, a bound that was not one) is skipped
    uint64_t src_cap = ~0ull;   // bytes of `in` (the filter scratch is exactly that large): same rule
};

// scratch was sized from host-known bounds; a frame beyond them gets B2B_EDST_TOO_SMALL instead of a wild write
__device__ __forceinline__ bool frame_fits_scratch(uint64_t comp_off, uint64_t seg_base, uint32_t n, uint64_t comp_cap,
                                                   uint64_t seg_cap, uint64_t src_off = 0, uint64_t src_cap = ~0ull) {
    return comp_off + frame_slot_bytes(n) <= comp_cap && seg_base + seg_count(n) <= seg_cap &&
           src_off <= src_cap && n <= src_cap - src_off;
}

template <int HL, int HB, bool PH>
__global__ void __launch_bounds__(kEncThreads, HL <= 10 ? 7 : (HL == 11 ? 5 : (HL == 12 ? 2 : 1)))
lz4_encode_kernel(EncodeArgs a) {
#ifndef B2B_EMU
    extern __shared__ __align__(16) uint32_t enc_tables[];   // kEncWarps x 2^HL entries
#else
    static uint32_t enc_tables[kEncWarps << 13];             // tests/emu: the CPU shim has no dynamic shared memory
#endif
    __shared__ LaneLists enc_lists[kEncWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t items = (uint64_t)a.nframes * a.segs_grid;
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(a.ticket, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= items) break;
        // segment-major ticket order: segment k of every frame before segment k+1 of any.  After a byte
        // shuffle segment k is byte plane k, so segments of equal cost sit together and the launch ends
        // on whatever the last plane costs instead of on one late, expensive segment
        const uint32_t f = (uint32_t)(item % a.nframes), s0 = (uint32_t)(item / a.nframes);
        const uint32_t n = a.src_len[f];
        const uint32_t nseg = seg_count(n);
        const uint8_t *frame = a.in + a.src_off[f];
        if (!frame_fits_scratch(a.comp_off[f], a.seg_base[f], n, a.comp_cap, a.seg_cap, a.src_off[f], a.src_cap)) continue;
        // short frames are mostly the raw tail the bit shuffle leaves alone: no place in the key there
        const uint32_t pmask = (PH && n >= 32u * (a.phase_mask + 1u)) ? a.phase_mask : 0u;
        for (uint32_t s = s0; s < nseg; s += a.segs_grid) {
            const uint32_t B = s * kSegBytes;
            const uint32_t L = n - B < kSegBytes ? n - B : kSegBytes;
            // after a byte shuffle a segment that starts on a plane boundary has nothing to learn from the
            // tail of the previous plane (another byte of the element): no warm-up window there
            const bool plane_start = a.planes > 1 && B % (n / a.planes) == 0;
            const uint32_t W = (a.independent || plane_start) ? 0u : (B < kWarmBytes ? B : kWarmBytes);
            const SegMeta m = warp_encode_segment<HL, HB, PH>(frame + B - W, W, L, (uint64_t)n - B - L,
                                                      a.comp + a.comp_off[f] + (uint64_t)s * kSegSlot,
                                                      enc_tables + ((size_t)warp << HL), &enc_lists[warp], lane,
                                                      a.tune[0] ? a.tune[0] : 256u,
                                                      a.tune[1] ? (a.tune[1] < 64u ? a.tune[1] : 64u) : kStrip,
                                                      a.tune[2] ? a.tune[2] : kStripCap, B == 0,
                                                      a.tune[3] ? a.tune[3] - 1u : 2u, (B - W) & pmask, pmask);
            if (lane == 0) a.meta[a.seg_base[f] + s] = m;
            __syncwarp();
        }
    }
}

// ---- finalize: one thread per frame ------------------------------------------------------
struct FinalizeArgs {
    const uint32_t *src_len;
    const uint64_t *seg_base;
    const SegMeta *meta;
    SegPlace *place;
    uint32_t nframes;
    uint32_t shuffle_flag;      // B2B_FLAG_SHUFFLE / B2B_FLAG_BITSHUFFLE / 0 (set even when T<=1)
    uint32_t keep_raw;          // raw-block API: never substitute the memcpy payload
    uint32_t *comp_len;         // out: payload bytes stored (c, or n for memcpy)
    uint32_t *frame_len;        // out: 16 + payload (0 when status != 0)
    uint32_t *flags;            // out: header flags
    uint32_t *final_ll;         // out: literals of the closing token
    uint32_t *final_off;        // out: payload offset of the closing token
    uint32_t *status;
    // side-car decode index (optional): entry (f, s) = payload offset of the token that opens segment
    // s's sequences | output position of that token's first literal << 32; ~0 if the segment has none
    uint64_t *index;
    uint32_t segs_per_frame;
    const uint64_t *comp_off;   // with comp_cap / seg_cap / src_off / src_cap: the same capacity check as the encoder's
    uint64_t comp_cap, seg_cap;
    const uint64_t *src_off = nullptr;
    uint64_t src_cap = ~0ull;
};

// segments [lo, hi) of a frame learn that the token in front of their literal run takes hdr bytes (four loads in flight)
__device__ __forceinline__ void fix_trail_dst(SegPlace *place, uint32_t lo, uint32_t hi, uint32_t hdr, uint32_t lane) {
    for (uint32_t q = lo + lane; q < hi; q += 128) {
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = q + 32u * u < hi ? place[q + 32u * u].trail_dst : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) if (q + 32u * u < hi) place[q + 32u * u].trail_dst = v[u] + hdr;
    }
}

// One WARP per frame: the walk over the segments is a serial recurrence (where a segment's first token goes depends
// on everything before it), but its cost was the dependent load of one summary per step -- one thread walking the
// 16 384 segments of a 1 GiB frame took 2.6 ms, as long as the encoder.  The lanes load 32 summaries at once, every
// lane replays the 32 steps from shuffles (same state in every lane), and lane j keeps what step j produced.
__global__ void finalize_frames_kernel(FinalizeArgs a) {
    const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (f >= a.nframes) return;
    const uint32_t n = a.src_len[f];
    uint32_t st = 0, flags = a.shuffle_flag, c = 0, flen = 0;
    const uint64_t ibase = (uint64_t)f * a.segs_per_frame;
    bool walk = false;
    if (n == 0) st = 1;                                   // ErrInvalidData, blosc.go:269-271
    else if (n > 0xFFFFFFFFu - 16u) st = 6;               // header fields are u32
    else if (!frame_fits_scratch(a.comp_off[f], a.seg_base[f], n, a.comp_cap, a.seg_cap, a.src_off ? a.src_off[f] : 0, a.src_cap))
        st = 11;                                          // B2B_EDST_TOO_SMALL: the scratch bound was not one
    else walk = true;
    if (!walk) {
        if (a.index) for (uint32_t s = lane; s < a.segs_per_frame; s += 32) a.index[ibase + s] = ~0ull;
    } else {
        const uint32_t nseg = seg_count(n);
        const uint64_t base = a.seg_base[f];
        uint64_t out = 0, carry = 0;
        // the literal run that is open: its token will stand at payload offset `out` (fixed until the run closes), its
        // literals come from the segments run_first .. (trailing literals of a segment with a match, then whole segments
        // without one); where a segment's share starts inside the run is known at once, the token's length bytes in
        // front of it only when the run closes -- they are added then (to the lanes of this group, and in memory to the
        // segments of earlier groups: the incompressible planes of one large frame are a run of a thousand segments)
        uint32_t run_first = 0;
        if (a.index) for (uint32_t s = nseg + lane; s < a.segs_per_frame; s += 32) a.index[ibase + s] = ~0ull;
        for (uint32_t s0 = 0; s0 < nseg; s0 += 32) {
            const uint32_t s = s0 + lane;
            SegMeta m; m.first_ll = 0; m.body_len = 0; m.trail_ll = 0; m.info = 0;
            if (s < nseg) m = a.meta[base + s];
            SegPlace pl; pl.out_off = 0; pl.lit_total = 0; pl.trail_dst = 0; pl.pad = 0;
            uint64_t idx = ~0ull;
            const uint32_t cnt = nseg - s0 < 32u ? nseg - s0 : 32u;
            for (uint32_t j = 0; j < cnt; j++) {
                const uint32_t info = __shfl_sync(0xffffffffu, m.info, (int)j), first = __shfl_sync(0xffffffffu, m.first_ll, (int)j);
                const uint32_t body = __shfl_sync(0xffffffffu, m.body_len, (int)j), trail = __shfl_sync(0xffffffffu, m.trail_ll, (int)j);
                if (info & 0x100u) {
                    const uint64_t lt = carry + first;
                    const uint32_t hdr = 1u + len_ext_bytes((uint32_t)lt);
                    // the run closes: its contributors learn where the literals start
                    if (s >= run_first && lane < j) pl.trail_dst += hdr;
                    if (run_first < s0) fix_trail_dst(a.place + base, run_first, s0, hdr, lane);
                    if (lane == j) {
                        pl.out_off = (uint32_t)(out > 0xFFFFFFFFull ? 0xFFFFFFFFull : out);
                        pl.lit_total = (uint32_t)lt;
                        idx = (uint64_t)pl.out_off | (((uint64_t)(s0 + j) * kSegBytes + first - lt) << 32);
                    }
                    out += 1ull + hdr - 1u + lt + body;
                    carry = trail;
                    run_first = s0 + j;                   // its trailing literals open the next run, at its very start
                    if (lane == j) pl.trail_dst = (uint32_t)out;
                } else {
                    if (lane == j) pl.trail_dst = (uint32_t)(out + carry);
                    carry += trail;                       // no match: the whole segment is carried
                }
            }
            if (s < nseg) {
                a.place[base + s] = pl;
                if (a.index && s < a.segs_per_frame) a.index[ibase + s] = idx;
            }
            __threadfence_block();
            __syncwarp();                                 // (the stores above are read-modify-written by other lanes when a run closes)
        }
        {   // the closing token closes the last run
            const uint32_t hdr = 1u + len_ext_bytes((uint32_t)carry);
            fix_trail_dst(a.place + base, run_first, nseg, hdr, lane);
        }
        if (lane == 0) {
            a.final_ll[f] = (uint32_t)carry;
            a.final_off[f] = (uint32_t)(out > 0xFFFFFFFFull ? 0xFFFFFFFFull : out);
        }
        out += 1ull + len_ext_bytes((uint32_t)carry) + carry;
        if (out >= n && !a.keep_raw) { c = n; flags |= 0x2u; }   // blosc.go:342-345: store uncompressed
        else c = (uint32_t)out;
        flen = 16 + c;
    }
    if (lane == 0) { a.comp_len[f] = c; a.frame_len[f] = flen; a.flags[f] = flags; a.status[f] = st; }
}

// ---- pack: one CTA per segment -------------------------------------------------------------
struct PackArgs {
    const uint8_t *in;          // what the LZ4 block was made from (shuffled bytes)
    const uint8_t *raw;         // what a memcpy frame stores (shuffled bytes, or the caller's
                                //   original bytes under B2B_OPT_REF_MEMCPY_QUIRK)
    const uint64_t *src_off;
    const uint32_t *src_len;
    const uint8_t *comp;
    const uint64_t *comp_off;
    const uint64_t *seg_base;
    const SegMeta *meta;
    const SegPlace *place;
    const uint32_t *comp_len, *flags, *final_ll, *final_off;
    uint32_t *status;
    uint64_t dst_cap = 0;       // bytes of dst (0: not checked): a frame that would reach beyond gets B2B_EDST_TOO_SMALL
    const uint64_t *frame_off;  // packed offsets (16-byte aligned)
    uint8_t *dst;
    uint32_t nframes, segs_grid, codec, typesize_u8;
    uint32_t header;            // 1: write the 16-byte frame header (0 for the raw-block API)
};

// token with literal length lt and match nibble, followed by the length extension; returns bytes
__device__ __forceinline__ uint32_t cta_put_token(uint8_t *dst, uint32_t lt, uint32_t nib) {
    if (threadIdx.x == 0) dst[0] = (uint8_t)(((lt < 15 ? lt : 15u) << 4) | nib);
    if (lt < 15) return 1;
    const uint32_t v = lt - 15, full = v / 255u;
    for (uint32_t i = threadIdx.x; i < full; i += blockDim.x) dst[1 + i] = 255;
    if (threadIdx.x == 0) dst[1 + full] = (uint8_t)(v - full * 255u);
    return 2 + full;
}

// bytes a token with lt literals takes before the literals
__device__ __forceinline__ uint32_t token_hdr_bytes(uint32_t lt) { return 1u + len_ext_bytes(lt); }

// segment s of frame f's LZ4 block, assembled at its place in `out` (the start of the block) by the whole CTA.
// Every segment moves exactly the input bytes it OWNS: a literal run that crosses segment boundaries (the trailing
// literals of a segment, whole segments without a match -- the incompressible byte planes of a shuffled frame --
// and the leading literals of the segment that closes the run) is copied piecewise by its owners; the finalize pass
// gives every segment the place of its share (SegPlace.trail_dst).  (Round 1 let the closing segment copy the whole run:
// in ONE large frame that is a single CTA moving half the frame -- 256 MiB: 12.7 of 14.9 ms.)
__device__ __forceinline__ void cta_pack_lz4_segment(const PackArgs &a, uint32_t f, uint32_t s, uint32_t n,
                                                     uint32_t nseg, uint64_t base, uint8_t *out) {
    const uint32_t B = s * kSegBytes;
    const uint32_t L = n - B < kSegBytes ? n - B : kSegBytes;
    const uint8_t *frame = a.in + a.src_off[f];
    const SegMeta m = a.meta[base + s];
    if (m.info & 0x100u) {
        const SegPlace pl = a.place[base + s];
        uint8_t *d = out + pl.out_off;
        const uint32_t hdr = cta_put_token(d, pl.lit_total, m.info & 15u);
        // my own leading literals close the run; the segment body follows them
        cta_copy(d + hdr + (pl.lit_total - m.first_ll), frame + B, m.first_ll);
        cta_copy(d + hdr + pl.lit_total, a.comp + a.comp_off[f] + (uint64_t)s * kSegSlot, m.body_len);
    }
    if (s == nseg - 1) cta_put_token(out + a.final_off[f], a.final_ll[f], 0);      // closing token: the last literals
    // my trailing literals (the whole segment when it has no match) open or continue the literal run of the next token:
    // the finalize pass has worked out where they go
    const uint32_t mine = (m.info & 0x100u) ? m.trail_ll : L;
    if (mine == 0) return;
    const uint32_t dst_off = a.place[base + s].trail_dst;
    cta_copy(out + dst_off, frame + (B + L - mine), mine);
}

__global__ void __launch_bounds__(kFilterThreads, 8) pack_frames_kernel(PackArgs a) {
    const uint64_t item = blockIdx.x;
    const uint32_t f = (uint32_t)(item / a.segs_grid), s0 = (uint32_t)(item % a.segs_grid);
    if (f >= a.nframes || a.status[f] != 0) return;
    const uint32_t n = a.src_len[f], c = a.comp_len[f], flags = a.flags[f];
    const uint32_t nseg = seg_count(n);
    if (a.dst_cap && (a.frame_off[f] > a.dst_cap || (uint64_t)c + (a.header ? 16u : 0u) > a.dst_cap - a.frame_off[f])) {
        // only reachable when total_src_bytes understated the batch; every CTA of the frame takes this exit
        if (s0 == 0 && threadIdx.x == 0) a.status[f] = 11;
        return;
    }
    uint8_t *out = a.dst + a.frame_off[f];
    if (a.header) {
        if (s0 == 0 && threadIdx.x == 0) {
            // blosc.go:358-366: [2, codec, flags, uint8(T), n, n, 16 + c]
            uint4 h;
            h.x = 2u | (a.codec << 8) | (flags << 16) | (a.typesize_u8 << 24);
            h.y = n; h.z = n; h.w = 16u + c;
            *reinterpret_cast<uint4 *>(out) = h;
        }
        out += 16;
    }
    const uint64_t base = a.seg_base[f];
    for (uint32_t s = s0; s < nseg; s += a.segs_grid) {
        if (flags & 0x2u) {                                   // memcpy frame
            const uint32_t B = s * kSegBytes;
            const uint32_t L = n - B < kSegBytes ? n - B : kSegBytes;
            cta_copy(out + B, a.raw + a.src_off[f] + B, L);
            continue;
        }
        cta_pack_lz4_segment(a, f, s, n, nseg, base, out);
    }
}

}  // namespace b2b
