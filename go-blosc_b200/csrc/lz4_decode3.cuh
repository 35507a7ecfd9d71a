// lz4_decode3.cuh -- K4, third arrangement: one LANE per frame ("scalar decoders in lockstep").
//
// Replaces lz4Codec.Decompress (codec.go:77-84 -> pierrec UncompressBlock) for batches of MANY small frames.
// An LZ4 block is a serial chain of 10..20-byte sequences; the one-warp-per-frame kernels (lz4_kernels.cuh) spend
// about 64 warp instructions on each of them (24 in the parse half, 40 in the copy half) because a whole warp
// does the bookkeeping of one chain.  Here a chain belongs to ONE thread, which decodes it like a scalar
// decoder -- token, literals, offset, match, with exactly the checks of warp_decode_one in the same order --
// and the 32 lanes of a warp run 32 such decoders in lockstep, one TURN at a time (a vote per turn keeps them
// converged): a turn parses what is due (a token, or an offset and match length) and moves at most 16 bytes.
// Sequences of a typed-array stream take two turns (literals, match), so the lanes of a warp stay in phase.
//
// What makes this work on a GPU is that a lane never touches global memory a byte at a time (32 lanes on 32
// different frames = one cache-line wavefront per lane and instruction):
//   in    the stream is read as aligned 16-byte vectors into a 32-byte ring per lane (shared memory), and
//         tokens, offsets and literals are read from there;
//   out   output bytes collect in a 64-byte ring per lane and leave as one aligned 16-byte vector per completed
//         chunk; the ring doubles as the history for matches at offsets <= 32;
//   far   a match at a larger offset reads its (up to 16) bytes as the two aligned 16-byte vectors that cover
//         them -- the lane's own earlier output -- into a 32-byte scratch, and copies from there;
//   bulk  a literal run of 16 bytes or more at a 16-byte aligned output position moves as words: five words of
//         the in-ring, a funnel shift, one 16-byte store (the incompressible byte planes of a shuffled frame,
//         stored frames).
// The three buffers of a lane are 132 bytes apart from the next lane's (33 words: bank = lane + word), so the 32
// lanes of a shared-memory access hit different banks whatever their positions are.
// No record table, no parse kernel: scratch is the stage buffer only.
#pragma once
#include "common.cuh"
#include "lz4_kernels.cuh"

namespace b2b {

constexpr int kLaneThreads = 128;
constexpr uint32_t kLaneBuf = 132;        // bytes of shared memory per lane: in-ring 32 | far scratch 32 | out-ring 64 | pad 4
constexpr uint32_t kLaneIn = 0, kLaneFar = 32, kLaneOut = 64;
constexpr uint32_t kLaneNear = 32;        // offsets up to here are served by the out-ring

struct LaneDec {
    uint8_t *sb;                // this lane's shared memory
    const uint8_t *srcv;        // stream, rounded down to 16 bytes; u = stream position + g0
    uint8_t *outv;              // output, rounded down to 16 bytes; v = output position + a0
    uint32_t g0, a0;
    uint32_t uend;              // g0 + clen
    uint32_t vcap;              // a0 + capacity
    uint32_t u;                 // next stream byte
    uint32_t v;                 // next output byte
    uint32_t in_loaded;         // stream chunks [0, in_loaded) have been brought in (the ring keeps the last two)
    uint32_t ll, ml, off;       // what is left of the current sequence
    int state;                  // 0 token + literals, 1 offset + match, 2 done
    int rc;                     // 0 stream complete, -1 malformed, -2 over capacity (state 2)
};

__device__ __forceinline__ void lane_in_ensure(LaneDec &d, uint32_t u) {      // byte u of the stream is in the ring
    while (d.in_loaded <= (u >> 4)) {
        const uint4 x = *reinterpret_cast<const uint4 *>(d.srcv + 16ull * d.in_loaded);
        uint32_t *w = reinterpret_cast<uint32_t *>(d.sb + kLaneIn + 16u * (d.in_loaded & 1u));
        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
        d.in_loaded++;
    }
}
__device__ __forceinline__ uint32_t lane_in_byte(LaneDec &d, uint32_t u) {
    lane_in_ensure(d, u);
    return d.sb[kLaneIn + (u & 31u)];
}
// output chunk q (16 bytes) is complete: to global memory
__device__ __forceinline__ void lane_flush_chunk(LaneDec &d, uint32_t q) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(d.sb + kLaneOut + 16u * (q & 3u));
    if (q == 0 && d.a0) {
        for (uint32_t b = d.a0; b < 16u; b++) d.outv[b] = d.sb[kLaneOut + b];
    } else {
        stg128(d.outv + 16ull * q, make_uint4(w[0], w[1], w[2], w[3]));
    }
}
// The byte loops of a turn carry no branches: the stream bytes they read are brought in before the loop (one call,
// at most two vectors) and the chunks they complete leave after it (at most two), so the lanes of a warp only
// differ in their trip counts.
__device__ __forceinline__ void lane_flush_upto(LaneDec &d, uint32_t v_before) {      // chunks completed since v_before
    for (uint32_t q = v_before >> 4; q < (d.v >> 4); q++) lane_flush_chunk(d, q);
}
// length bytes behind a nibble of 15; false: ran off the stream.  (No early returns here or in lane_turn: the
// compiler only re-converges the lanes of a warp at the end of STRUCTURED control flow, and 32 decoders that run one
// after the other are 32 times slower than 32 that run together.)
__device__ __forceinline__ bool lane_len_ext(LaneDec &d, uint64_t &len) {
    bool ok = true, more = true;
    while (more) {
        if (d.u >= d.uend) { ok = false; more = false; }
        else {
            const uint32_t b = lane_in_byte(d, d.u++);
            len += b;
            more = b == 255u;
            if (len > 0xFFFFFFFFull) { ok = false; more = false; }
        }
    }
    return ok;
}

// One turn of one lane.  Order of the checks = warp_decode_one (lz4_kernels.cuh).
__device__ __forceinline__ void lane_turn(LaneDec &d) {
    int fail = 0;                                        // -1 / -2: the stream ends here with that code
    if (d.state == 0) {
        if (d.ll == 0xFFFFFFFFu) {                       // a new token
            if (d.u >= d.uend) fail = -1;
            else {
                const uint32_t tok = lane_in_byte(d, d.u++);
                uint64_t l = tok >> 4;
                if (l == 15 && !lane_len_ext(d, l)) fail = -1;
                else if (l > (uint64_t)(d.uend - d.u)) fail = -1;
                else if (l > (uint64_t)(d.vcap - d.v)) fail = -2;
                else { d.ll = (uint32_t)l; d.ml = tok & 15u; }   // the nibble; the length follows with the offset
            }
        }
        if (fail == 0) {
            // up to 16 literal bytes
            if (d.ll >= 16u && (d.v & 15u) == 0) {
                lane_in_ensure(d, d.u + 15u);
                const uint32_t wi = d.u >> 2, sh = 8u * (d.u & 3u);
                const uint32_t *ring = reinterpret_cast<const uint32_t *>(d.sb + kLaneIn);
                const uint32_t x0 = ring[wi & 7u], x1 = ring[(wi + 1u) & 7u], x2 = ring[(wi + 2u) & 7u],
                               x3 = ring[(wi + 3u) & 7u], x4 = ring[(wi + 4u) & 7u];
                const uint4 o = make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh),
                                           __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh));
                uint32_t *w = reinterpret_cast<uint32_t *>(d.sb + kLaneOut + (d.v & 63u));
                w[0] = o.x; w[1] = o.y; w[2] = o.z; w[3] = o.w;
                stg128(d.outv + d.v, o);
                d.u += 16u; d.v += 16u; d.ll -= 16u;
            } else {
                uint32_t n = d.ll < 16u ? d.ll : 16u;
                if (n > 16u - (d.v & 15u) && d.ll >= 32u) n = 16u - (d.v & 15u);     // a long run: get the output aligned first
                if (n) lane_in_ensure(d, d.u + n - 1u);
                const uint32_t v0 = d.v;
                for (uint32_t k = 0; k < n; k++) d.sb[kLaneOut + ((v0 + k) & 63u)] = d.sb[kLaneIn + ((d.u + k) & 31u)];
                d.v += n; d.u += n; d.ll -= n;
                lane_flush_upto(d, v0);
            }
            if (d.ll == 0) {
                if (d.u == d.uend) { d.state = 2; d.rc = d.ml != 0 ? -1 : 0; }       // closing token
                else { d.state = 1; d.off = 0; }
            }
        }
    } else {
        if (d.off == 0) {                                // offset and match length are due
            if (d.uend - d.u < 2u) fail = -1;
            else {
                const uint32_t lo = lane_in_byte(d, d.u), hi = lane_in_byte(d, d.u + 1u);
                d.u += 2u;
                const uint32_t off = lo | (hi << 8);
                uint64_t m = d.ml;
                if (off == 0 || off > d.v - d.a0) fail = -1;
                else if (m == 15 && !lane_len_ext(d, m)) fail = -1;
                else if (m + 4 > (uint64_t)(d.vcap - d.v)) fail = -2;
                else { d.off = off; d.ml = (uint32_t)(m + 4); }
            }
        }
        if (fail == 0) {
            const uint32_t n = d.ml < 16u ? d.ml : 16u;
            const uint32_t v0 = d.v;
            if (d.off <= kLaneNear) {
                const uint32_t s = v0 - d.off;           // (bytes this loop writes may be read by it again: in order)
                for (uint32_t k = 0; k < n; k++) d.sb[kLaneOut + ((v0 + k) & 63u)] = d.sb[kLaneOut + ((s + k) & 63u)];
            } else {
                const uint32_t s = d.v - d.off;          // its bytes [s, s + n) left the ring as whole chunks
                const uint8_t *g = d.outv + 16ull * (s >> 4);
                const uint4 x = *reinterpret_cast<const uint4 *>(g), y = *reinterpret_cast<const uint4 *>(g + 16);
                uint32_t *w0 = reinterpret_cast<uint32_t *>(d.sb + kLaneFar + 16u * ((s >> 4) & 1u));
                uint32_t *w1 = reinterpret_cast<uint32_t *>(d.sb + kLaneFar + 16u * (((s >> 4) + 1u) & 1u));
                w0[0] = x.x; w0[1] = x.y; w0[2] = x.z; w0[3] = x.w;
                w1[0] = y.x; w1[1] = y.y; w1[2] = y.z; w1[3] = y.w;
                for (uint32_t k = 0; k < n; k++) d.sb[kLaneOut + ((v0 + k) & 63u)] = d.sb[kLaneFar + ((s + k) & 31u)];
            }
            d.v += n;
            lane_flush_upto(d, v0);
            d.ml -= n;
            if (d.ml == 0) { d.state = 0; d.ll = 0xFFFFFFFFu; }
        }
    }
    if (fail != 0) { d.state = 2; d.rc = fail; }
}

// Returns the number of bytes produced, -1 for a malformed stream, -2 if it would overrun cap (like warp_lz4_decode).
// All lanes of the warp call it together (clen == 0xFFFFFFFF: this lane has no stream).
__device__ __forceinline__ int64_t lane_lz4_decode(const uint8_t *src, uint32_t clen, uint8_t *out, uint32_t cap,
                                                   uint8_t *sb, bool stored) {
    LaneDec d;
    d.sb = sb;
    const bool have = clen != 0xFFFFFFFFu;
    d.g0 = have ? (uint32_t)((uintptr_t)src & 15u) : 0u;
    d.a0 = have ? (uint32_t)((uintptr_t)out & 15u) : 0u;
    d.srcv = src - d.g0; d.outv = out - d.a0;
    d.uend = d.g0 + (have ? clen : 0u);
    d.vcap = d.a0 + cap;
    d.u = d.g0; d.v = d.a0; d.in_loaded = 0;
    d.ll = 0xFFFFFFFFu; d.ml = 0; d.off = 0;
    d.state = have && clen != 0 ? 0 : 2; d.rc = 0;
    if (have && stored && clen != 0) {                   // a stored frame is one literal run without a token
        d.ll = clen; d.ml = 0;
    }
    while (__any_sync(0xffffffffu, d.state != 2)) {
        if (d.state != 2) lane_turn(d);
    }
    if (!have) return 0;
    if (d.rc < 0) return d.rc;
    // the last, partial chunk
    for (uint32_t b = d.v & ~15u; b < d.v; b++)
        if (b >= d.a0) d.outv[b] = d.sb[kLaneOut + (b & 63u)];
    return (int64_t)(d.v - d.a0);
}

// one frame per lane: header checks and status words exactly as warp_decode_frame (lz4_kernels.cuh)
__global__ void __launch_bounds__(kLaneThreads, 8) lz4_lane_decode_kernel(DecodeArgs a) {
    __shared__ __align__(16) uint8_t s_buf[kLaneThreads * kLaneBuf];
    const uint32_t f = blockIdx.x * kLaneThreads + threadIdx.x;
    uint8_t *sb = s_buf + threadIdx.x * kLaneBuf;
    const bool inb = f < a.nframes;
    uint32_t flags = 0, codec = 0, tsz = 0, norig = 0, ncomp = 0, st = kOk, produced = 0;
    FrameMeta m; m.mode = 0; m.typesize = 0;
    const uint8_t *fr = nullptr;
    uint8_t *out = nullptr;
    uint32_t clen = 0xFFFFFFFFu, dcap = 0;
    bool is_memcpy = false;
    if (inb) {
        fr = a.frames + a.frame_off[f];
        st = check_header(fr, a.frame_len[f], flags, codec, tsz, norig, ncomp);
        if (st == kOk) {
            is_memcpy = (flags & 0x2u) != 0;
            if (!is_memcpy) {
                if (codec < 1 || codec > 5) st = kEInvalidCodec;          // blosc.go:403-407
                else if (codec != 1 && codec != 2) st = kEUnsupported;    // Snappy/ZLIB/ZSTD: host side
            }
        }
        if (st == kOk) {
            const uint32_t plen = ncomp - 16;
            const uint64_t T = a.typesize_override > 0 ? (uint64_t)a.typesize_override : (uint64_t)tsz;
            const uint32_t mode = (flags & 0x4u) ? 2u : ((flags & 0x1u) ? 1u : 0u);
            const bool active = mode != 0 && T > 1 && (uint64_t)norig >= T;
            m.mode = active ? mode : 0u;
            m.typesize = active ? (uint32_t)T : 0u;
            const uint32_t cap = a.dst_cap[f];
            out = (active ? a.scratch : a.dst) + a.dst_off[f];
            dcap = cap < norig ? cap : norig;
            if (is_memcpy) {
                if (plen != norig) st = kESizeMismatch;                   // blosc.go:398-400, 429-431
                else if (cap < norig) st = kEDstTooSmall;
                else clen = plen;
            } else {
                clen = plen;
            }
        }
    }
    const int64_t got = lane_lz4_decode(fr ? fr + 16 : nullptr, clen, out, dcap, sb, is_memcpy);
    if (!inb) return;
    if (st == kOk) {
        if (is_memcpy) produced = norig;
        else if (got == -1) st = kEDecompressionFailed;                   // blosc.go:410-413
        else if (got == -2) st = dcap == norig ? kEDecompressionFailed : kEDstTooSmall;
        else if ((uint64_t)got != norig) { st = kESizeMismatch; produced = (uint32_t)got; }   // blosc.go:429-431
        else produced = norig;
    }
    if (st != kOk) { m.mode = 0; m.typesize = 0; if (st != kESizeMismatch) produced = 0; }
    a.status[f] = st;
    a.out_len[f] = produced;
    a.meta[f] = m;
}

}  // namespace b2b
