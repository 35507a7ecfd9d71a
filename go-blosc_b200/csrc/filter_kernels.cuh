// filter_kernels.cuh -- K1 byte shuffle / unshuffle and K2 bit shuffle / unshuffle.
//
// Replaces shuffle.go:16-295 and the amd64/arm64 assembly behind it (shuffle_amd64.s,
// shuffle_arm64.s).  Semantics (bit-exact):
//   byte shuffle   dst[j*E + i] = src[i*T + j]         E = n / T, tail [E*T, n) copied raw
//   bit  shuffle   per group of 8 elements and byte position j an 8x8 bit transpose
//                  (MSB first), groups are 8*T bytes and local; E % 8 leftover elements and
//                  the n % T tail copied raw
//   T <= 1 or n < T  : identity
//
// One launch covers a whole batch: CTA b works on frame b / tiles_per_frame and strides
// over that frame's 16 KiB tiles.  The filter mode and typesize are uniform for a compress
// batch and per frame (from the header) on decompress, hence the in-kernel dispatch.
//
// HBM-bound (2n algorithmic bytes).  Fast paths for T in {2,4,8,16} need 16-byte aligned
// frame bases (and E % 16 == 0 for the byte shuffle); every global access is then a
// coalesced 16-byte vector.  K1 stages the tile through shared memory as T byte planes
// (registers -> PRMT-packed plane words -> smem -> 16-byte rows out); K2 is register-only
// (each thread owns one 8-element group: byte gather by PRMT 4x4 transposes, then three
// delta swaps per 8x8 bit matrix).  Everything else takes the generic byte-granular path.
#pragma once
#include "common.cuh"

namespace b2b {

// =========================================================================================
// K1 fast tiles
// =========================================================================================
template <int T> struct ShufCfg {
    static constexpr int TE = kTileBytes / T;  // elements per tile
    static constexpr int EPV = 16 / T;         // elements per 16-byte vector
};

// scatter the 16 loaded bytes (EPV elements starting at tile element e) into the T smem planes
template <int T>
__host__ __device__ __forceinline__ void planes_from_vec(uint8_t *smem, uint32_t e, const uint4 &v) {
    constexpr int TE = ShufCfg<T>::TE;
    if constexpr (T == 2) {
        uint2 p0 = make_uint2(prmt(v.x, v.y, 0x6420), prmt(v.z, v.w, 0x6420));
        uint2 p1 = make_uint2(prmt(v.x, v.y, 0x7531), prmt(v.z, v.w, 0x7531));
        *reinterpret_cast<uint2 *>(smem + e) = p0;
        *reinterpret_cast<uint2 *>(smem + TE + e) = p1;
    } else if constexpr (T == 4) {
        uint32_t p0, p1, p2, p3;
        transpose4x4(v.x, v.y, v.z, v.w, p0, p1, p2, p3);
        *reinterpret_cast<uint32_t *>(smem + 0 * TE + e) = p0;
        *reinterpret_cast<uint32_t *>(smem + 1 * TE + e) = p1;
        *reinterpret_cast<uint32_t *>(smem + 2 * TE + e) = p2;
        *reinterpret_cast<uint32_t *>(smem + 3 * TE + e) = p3;
    } else if constexpr (T == 8) {
        // elements (x,y) and (z,w): plane j gets (elem0.byte j, elem1.byte j)
        uint32_t lo01 = prmt(v.x, v.z, 0x5140);  // x0 z0 x1 z1
        uint32_t lo23 = prmt(v.x, v.z, 0x7362);  // x2 z2 x3 z3
        uint32_t hi01 = prmt(v.y, v.w, 0x5140);
        uint32_t hi23 = prmt(v.y, v.w, 0x7362);
        *reinterpret_cast<uint16_t *>(smem + 0 * TE + e) = (uint16_t)lo01;
        *reinterpret_cast<uint16_t *>(smem + 1 * TE + e) = (uint16_t)(lo01 >> 16);
        *reinterpret_cast<uint16_t *>(smem + 2 * TE + e) = (uint16_t)lo23;
        *reinterpret_cast<uint16_t *>(smem + 3 * TE + e) = (uint16_t)(lo23 >> 16);
        *reinterpret_cast<uint16_t *>(smem + 4 * TE + e) = (uint16_t)hi01;
        *reinterpret_cast<uint16_t *>(smem + 5 * TE + e) = (uint16_t)(hi01 >> 16);
        *reinterpret_cast<uint16_t *>(smem + 6 * TE + e) = (uint16_t)hi23;
        *reinterpret_cast<uint16_t *>(smem + 7 * TE + e) = (uint16_t)(hi23 >> 16);
    } else {  // T == 16: one element, one byte per plane
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++) smem[j * TE + e] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
    }
}

// gather EPV elements (16 bytes) starting at tile element e back out of the T smem planes
template <int T>
__host__ __device__ __forceinline__ uint4 vec_from_planes(const uint8_t *smem, uint32_t e) {
    constexpr int TE = ShufCfg<T>::TE;
    uint4 v;
    if constexpr (T == 2) {
        uint2 p0 = *reinterpret_cast<const uint2 *>(smem + e);
        uint2 p1 = *reinterpret_cast<const uint2 *>(smem + TE + e);
        v.x = prmt(p0.x, p1.x, 0x5140); v.y = prmt(p0.x, p1.x, 0x7362);
        v.z = prmt(p0.y, p1.y, 0x5140); v.w = prmt(p0.y, p1.y, 0x7362);
    } else if constexpr (T == 4) {
        uint32_t p0 = *reinterpret_cast<const uint32_t *>(smem + 0 * TE + e);
        uint32_t p1 = *reinterpret_cast<const uint32_t *>(smem + 1 * TE + e);
        uint32_t p2 = *reinterpret_cast<const uint32_t *>(smem + 2 * TE + e);
        uint32_t p3 = *reinterpret_cast<const uint32_t *>(smem + 3 * TE + e);
        transpose4x4(p0, p1, p2, p3, v.x, v.y, v.z, v.w);
    } else if constexpr (T == 8) {
        uint32_t q[8];
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = *reinterpret_cast<const uint16_t *>(smem + j * TE + e);
        // q[j] = (elem0.byte j, elem1.byte j)
        uint32_t a = prmt(q[0], q[1], 0x5140), b = prmt(q[2], q[3], 0x5140);  // e0b0 e0b1 e1b0 e1b1 ...
        uint32_t c = prmt(q[4], q[5], 0x5140), d = prmt(q[6], q[7], 0x5140);
        v.x = prmt(a, b, 0x5410); v.y = prmt(c, d, 0x5410);
        v.z = prmt(a, b, 0x7632); v.w = prmt(c, d, 0x7632);
    } else {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; j++) w[j >> 2] |= (uint32_t)smem[j * TE + e] << (8 * (j & 3));
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return v;
}

// One tile of the forward byte shuffle.  src/dst: frame bases (16-byte aligned), E % 16 == 0.
template <int T>
__device__ __forceinline__ void shuffle_tile(const uint8_t *__restrict__ src,
                                             uint8_t *__restrict__ dst, uint64_t E, uint64_t tile,
                                             uint8_t *smem) {
    constexpr int TE = ShufCfg<T>::TE, EPV = ShufCfg<T>::EPV;
    constexpr int NV = kTileBytes / 16 / kFilterThreads;  // vectors per thread
    const uint64_t e0 = tile * TE;
    const uint32_t ve = (uint32_t)(E - e0 < (uint64_t)TE ? E - e0 : (uint64_t)TE);  // multiple of 16
    const uint8_t *s = src + e0 * T;
    uint4 v[NV];
#pragma unroll
    for (int it = 0; it < NV; it++) {
        uint32_t idx = it * kFilterThreads + threadIdx.x;
        if (idx * EPV < ve) v[it] = ldg128_stream(s + 16ull * idx);
    }
#pragma unroll
    for (int it = 0; it < NV; it++) {
        uint32_t idx = it * kFilterThreads + threadIdx.x;
        if (idx * EPV < ve) planes_from_vec<T>(smem, idx * EPV, v[it]);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < NV; it++) {
        uint32_t b = 16u * (it * kFilterThreads + threadIdx.x);
        uint32_t j = b / TE, off = b % TE;
        if (off < ve) {
            uint4 o = *reinterpret_cast<const uint4 *>(smem + b);
            stg128_stream(dst + (uint64_t)j * E + e0 + off, o);
        }
    }
    __syncthreads();
}

template <int T>
__device__ __forceinline__ void unshuffle_tile(const uint8_t *__restrict__ src,
                                               uint8_t *__restrict__ dst, uint64_t E,
                                               uint64_t tile, uint8_t *smem) {
    constexpr int TE = ShufCfg<T>::TE, EPV = ShufCfg<T>::EPV;
    constexpr int NV = kTileBytes / 16 / kFilterThreads;
    const uint64_t e0 = tile * TE;
    const uint32_t ve = (uint32_t)(E - e0 < (uint64_t)TE ? E - e0 : (uint64_t)TE);
    uint4 v[NV];
#pragma unroll
    for (int it = 0; it < NV; it++) {
        uint32_t b = 16u * (it * kFilterThreads + threadIdx.x);
        uint32_t j = b / TE, off = b % TE;
        if (off < ve) v[it] = ldg128_stream(src + (uint64_t)j * E + e0 + off);
    }
#pragma unroll
    for (int it = 0; it < NV; it++) {
        uint32_t b = 16u * (it * kFilterThreads + threadIdx.x);
        if (b % TE < ve) *reinterpret_cast<uint4 *>(smem + b) = v[it];
    }
    __syncthreads();
    uint8_t *d = dst + e0 * T;
#pragma unroll
    for (int it = 0; it < NV; it++) {
        uint32_t idx = it * kFilterThreads + threadIdx.x;
        if (idx * EPV < ve) stg128_stream(d + 16ull * idx, vec_from_planes<T>(smem, idx * EPV));
    }
    __syncthreads();
}

// Generic byte shuffle tile (any T >= 2, any alignment): direct gather, coalesced writes.
__device__ __forceinline__ void shuffle_tile_generic(const uint8_t *__restrict__ src,
                                                     uint8_t *__restrict__ dst, uint64_t E,
                                                     uint64_t T, uint64_t e0, uint32_t ve,
                                                     bool inverse) {
    // ve * T <= ~2 * kTileBytes except for huge T (then ve == 1)
    // ve * T <= kTileBytes, or ve == 1 and T < 2^32: 32-bit index arithmetic suffices
    const uint32_t total = ve * (uint32_t)T, t32 = (uint32_t)T;
    if (!inverse) {
        for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
            uint32_t j = idx / ve, i = idx - j * ve;
            dst[(uint64_t)j * E + e0 + i] = src[(e0 + i) * T + j];
        }
    } else {
        for (uint32_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
            uint32_t i = idx / t32, j = idx - i * t32;
            dst[(e0 + i) * T + j] = src[(uint64_t)j * E + e0 + i];
        }
    }
}

// ---- K1 staged tiles: any typesize up to kStageMaxT, any alignment, any element count ------------------
// What the 16-byte fast path cannot take (T = 3, 5, 6, 7, 12, ...; frame bases that are not 16-byte aligned;
// E % 16 != 0, i.e. most frame sizes that are not powers of two) used to go through shuffle_tile_generic:
// byte-granular global accesses on both sides, 0.75 - 1.6 TB/s.  Here the CONTIGUOUS side of the transpose is
// staged in shared memory and every global access is a 16-byte vector on a 16-byte boundary:
//   * the contiguous side is loaded / stored from the aligned address below its first byte (the few bytes in
//     front belong to the same 16-byte block, so the access stays inside the buffer);
//   * on the plane side one thread owns one ALIGNED 16-byte vector of one plane segment: it gathers its 16
//     bytes from (scatters them to) the staged elements at stride T; only the first and the last vector of a
//     segment, which reach outside it, move single bytes.
// Lanes run over the planes first, so the 32 lanes of a shared-memory access touch neighbouring bytes.
constexpr uint32_t kStageTile = 16384;                             // bytes of one staged tile
constexpr uint32_t kStageLin = kStageTile + 32;                    // + misalignment + over-read
constexpr uint32_t kStageMaxT = 16;                                // (T = 32 measured no faster than the byte-granular path)

// forward: elements [e0, e0 + ve) of the frame -> T plane segments
__device__ __forceinline__ void shuffle_tile_staged(const uint8_t *__restrict__ s, uint8_t *__restrict__ d, uint64_t E,
                                                    uint32_t T, uint64_t e0, uint32_t ve, uint8_t *lin) {
    const uint32_t nb = ve * T;
    const uint8_t *g0 = s + e0 * T;
    const uint32_t mis = (uint32_t)((uintptr_t)g0 & 15u);
    const uint32_t nvec = (mis + nb + 15u) >> 4;
    for (uint32_t v = threadIdx.x; v < nvec; v += kFilterThreads)
        *reinterpret_cast<uint4 *>(lin + 16u * v) = ldg128_stream(g0 - mis + 16ull * v);
    __syncthreads();
    const uint32_t d0 = (uint32_t)((uintptr_t)(d + e0) & 15u), dE = (uint32_t)(E & 15u);   // plane j starts at (d0 + j * dE) & 15
    const uint32_t NV = (15u + ve + 15u) >> 4, items = NV * T;       // vectors a segment can touch
    uint32_t v = threadIdx.x / T, j = threadIdx.x - v * T;
    const uint32_t dv = kFilterThreads / T, dj = kFilterThreads - dv * T;
    for (uint32_t idx = threadIdx.x; idx < items; idx += kFilterThreads) {
        const uint32_t aj = (d0 + j * dE) & 15u;
        const uint32_t lo = 16u * v, hi = lo + 16u;                   // this vector's bytes, counted from the aligned address
        if (lo < aj + ve && hi > aj) {
            uint8_t *gp = d + (uint64_t)j * E + e0 - aj;              // 16-byte aligned
            if (lo >= aj && hi <= aj + ve) {
                const uint8_t *p = lin + mis + (lo - aj) * T + j;     // element lo - aj, byte j
                uint32_t w[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    w[k] = (uint32_t)p[0] | ((uint32_t)p[T] << 8) | ((uint32_t)p[2u * T] << 16) | ((uint32_t)p[3u * T] << 24);
                    p += 4u * T;
                }
                stg128_stream(gp + lo, make_uint4(w[0], w[1], w[2], w[3]));
            } else {
                for (uint32_t b = (lo > aj ? lo : aj); b < hi && b < aj + ve; b++) gp[b] = lin[mis + (b - aj) * T + j];
            }
        }
        v += dv; j += dj;
        if (j >= T) { j -= T; v++; }
    }
    __syncthreads();
}

// inverse: T plane segments [e0, e0 + ve) -> elements
__device__ __forceinline__ void unshuffle_tile_staged(const uint8_t *__restrict__ s, uint8_t *__restrict__ d, uint64_t E,
                                                      uint32_t T, uint64_t e0, uint32_t ve, uint8_t *lin) {
    const uint32_t nb = ve * T;
    uint8_t *g0 = d + e0 * T;
    const uint32_t mis = (uint32_t)((uintptr_t)g0 & 15u);
    const uint32_t s0 = (uint32_t)((uintptr_t)(s + e0) & 15u), sE = (uint32_t)(E & 15u);   // plane j starts at (s0 + j * sE) & 15
    const uint32_t NV = (15u + ve + 15u) >> 4, items = NV * T;
    uint32_t v = threadIdx.x / T, j = threadIdx.x - v * T;
    const uint32_t dv = kFilterThreads / T, dj = kFilterThreads - dv * T;
    for (uint32_t idx = threadIdx.x; idx < items; idx += kFilterThreads) {
        const uint32_t aj = (s0 + j * sE) & 15u;
        const uint32_t lo = 16u * v, hi = lo + 16u;
        if (lo < aj + ve && hi > aj) {
            const uint8_t *gp = s + (uint64_t)j * E + e0 - aj;        // 16-byte aligned
            const uint4 x = ldg128_stream(gp + lo);
            const uint32_t w[4] = {x.x, x.y, x.z, x.w};
            if (lo >= aj && hi <= aj + ve) {
                uint8_t *p = lin + mis + (lo - aj) * T + j;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    p[0] = (uint8_t)w[k]; p[T] = (uint8_t)(w[k] >> 8); p[2u * T] = (uint8_t)(w[k] >> 16); p[3u * T] = (uint8_t)(w[k] >> 24);
                    p += 4u * T;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t b = lo + (uint32_t)k;
                    if (b >= aj && b < aj + ve) lin[mis + (b - aj) * T + j] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
                }
            }
        }
        v += dv; j += dj;
        if (j >= T) { j -= T; v++; }
    }
    __syncthreads();
    const uint32_t nvec = (mis + nb + 15u) >> 4;
    for (uint32_t q = threadIdx.x; q < nvec; q += kFilterThreads) {
        const uint32_t lo = 16u * q, hi = lo + 16u;
        if (lo >= mis && hi <= mis + nb) stg128_stream(g0 - mis + lo, *reinterpret_cast<const uint4 *>(lin + lo));
        else for (uint32_t b = (lo > mis ? lo : mis); b < hi && b < mis + nb; b++) (g0 - mis)[b] = lin[b];
    }
    __syncthreads();
}

// =========================================================================================
// K2 fast groups: one thread owns one group of 8 elements (8*T contiguous bytes)
// =========================================================================================
template <int T>
__host__ __device__ __forceinline__ void bitshuffle_group(const uint8_t *__restrict__ src,
                                                 uint8_t *__restrict__ dst) {
    constexpr int NW = 2 * T;  // 32-bit words in the group
    uint32_t w[NW], o[NW];
#pragma unroll
    for (int k = 0; k < NW / 4; k++) {
        uint4 v = ldg128_stream(src + 16 * k);
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
    if constexpr (T == 2) {
        uint32_t lo0 = prmt(w[0], w[1], 0x6420), hi0 = prmt(w[2], w[3], 0x6420);
        uint32_t lo1 = prmt(w[0], w[1], 0x7531), hi1 = prmt(w[2], w[3], 0x7531);
        bit_transpose8(lo0, hi0); bit_transpose8(lo1, hi1);
        o[0] = lo0; o[1] = hi0; o[2] = lo1; o[3] = hi1;
    } else {
        constexpr int WPE = T / 4;  // words per element
#pragma unroll
        for (int h = 0; h < WPE; h++) {
            uint32_t lo[4], hi[4];
            transpose4x4(w[0 * WPE + h], w[1 * WPE + h], w[2 * WPE + h], w[3 * WPE + h], lo[0],
                         lo[1], lo[2], lo[3]);
            transpose4x4(w[4 * WPE + h], w[5 * WPE + h], w[6 * WPE + h], w[7 * WPE + h], hi[0],
                         hi[1], hi[2], hi[3]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                bit_transpose8(lo[k], hi[k]);
                o[2 * (4 * h + k)] = lo[k];
                o[2 * (4 * h + k) + 1] = hi[k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NW / 4; k++)
        stg128_stream(dst + 16 * k, make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
}

template <int T>
__host__ __device__ __forceinline__ void bitunshuffle_group(const uint8_t *__restrict__ src,
                                                   uint8_t *__restrict__ dst) {
    constexpr int NW = 2 * T;
    uint32_t w[NW], o[NW];
#pragma unroll
    for (int k = 0; k < NW / 4; k++) {
        uint4 v = ldg128_stream(src + 16 * k);
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
    if constexpr (T == 2) {
        uint32_t lo0 = w[0], hi0 = w[1], lo1 = w[2], hi1 = w[3];
        bit_transpose8(lo0, hi0); bit_transpose8(lo1, hi1);
        o[0] = prmt(lo0, lo1, 0x5140); o[1] = prmt(lo0, lo1, 0x7362);
        o[2] = prmt(hi0, hi1, 0x5140); o[3] = prmt(hi0, hi1, 0x7362);
    } else {
        constexpr int WPE = T / 4;
#pragma unroll
        for (int h = 0; h < WPE; h++) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                lo[k] = w[2 * (4 * h + k)];
                hi[k] = w[2 * (4 * h + k) + 1];
                bit_transpose8(lo[k], hi[k]);
            }
            transpose4x4(lo[0], lo[1], lo[2], lo[3], o[0 * WPE + h], o[1 * WPE + h], o[2 * WPE + h],
                         o[3 * WPE + h]);
            transpose4x4(hi[0], hi[1], hi[2], hi[3], o[4 * WPE + h], o[5 * WPE + h], o[6 * WPE + h],
                         o[7 * WPE + h]);
        }
    }
#pragma unroll
    for (int k = 0; k < NW / 4; k++)
        stg128_stream(dst + 16 * k, make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
}

// generic bit (un)shuffle item: group g, byte position j; any T, any alignment
__host__ __device__ __forceinline__ void bitshuffle_item_generic(const uint8_t *__restrict__ src,
                                                        uint8_t *__restrict__ dst, uint64_t T,
                                                        uint64_t g, uint64_t j, bool inverse) {
    const uint64_t base = g * 8 * T;
    uint32_t lo = 0, hi = 0;
    if (!inverse) {
#pragma unroll
        for (int m = 0; m < 4; m++) lo |= (uint32_t)src[base + (uint64_t)m * T + j] << (8 * m);
#pragma unroll
        for (int m = 0; m < 4; m++) hi |= (uint32_t)src[base + (uint64_t)(m + 4) * T + j] << (8 * m);
        bit_transpose8(lo, hi);
        uint8_t *d = dst + base + 8 * j;
#pragma unroll
        for (int k = 0; k < 4; k++) { d[k] = (uint8_t)(lo >> (8 * k)); d[k + 4] = (uint8_t)(hi >> (8 * k)); }
    } else {
        const uint8_t *s = src + base + 8 * j;
#pragma unroll
        for (int k = 0; k < 4; k++) { lo |= (uint32_t)s[k] << (8 * k); hi |= (uint32_t)s[k + 4] << (8 * k); }
        bit_transpose8(lo, hi);
#pragma unroll
        for (int m = 0; m < 4; m++) {
            dst[base + (uint64_t)m * T + j] = (uint8_t)(lo >> (8 * m));
            dst[base + (uint64_t)(m + 4) * T + j] = (uint8_t)(hi >> (8 * m));
        }
    }
}

// =========================================================================================
// The batch filter kernel
// =========================================================================================
struct FilterArgs {
    const uint8_t *src;
    uint8_t *dst;
    FrameTable ft;
    const FrameMeta *meta;  // per frame (decompress) or null -> uniform below
    FrameMeta uniform;
    const uint32_t *status;  // optional: frames whose status != 0 are skipped
    int inverse;
    int copy_inactive;       // frames without an active filter: 1 = copy src->dst, 0 = leave
    uint64_t limit = 0;      // bytes of src / dst (0: not checked): a frame that reaches beyond is left alone
};

template <int T>
__device__ __forceinline__ void run_shuffle_fast(const uint8_t *s, uint8_t *d, uint64_t E,
                                                 uint32_t tile0, uint32_t tstride, bool inverse,
                                                 uint8_t *smem) {
    const uint64_t ntiles = (E + ShufCfg<T>::TE - 1) / ShufCfg<T>::TE;
    for (uint64_t t = tile0; t < ntiles; t += tstride) {
        if (!inverse) shuffle_tile<T>(s, d, E, t, smem);
        else unshuffle_tile<T>(s, d, E, t, smem);
    }
}

// ---- K2, coalesced form: every lane moves ONE 16-byte vector in and out --------------------
// A group of 8 elements (8*T bytes) spans L = T/2 consecutive lanes.  Lane r of a group holds
// 8/L whole elements after the load and must end up with output vector r = the 8x8 bit
// matrices of byte positions 2r and 2r+1, which need two bytes of EVERY element of the group:
// an all-to-all inside the group by __shfl_xor (L-1 rounds of 16/L bytes).  The partner order
// r^s is undone with a block-xor permutation (PRMT / register swaps).  Bit unshuffle is the
// mirror image.  All global accesses are 512 contiguous bytes per warp instruction.
__device__ __forceinline__ uint32_t sel4(uint32_t i, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t lo = (i & 1u) ? b : a, hi = (i & 1u) ? d : c;
    return (i & 2u) ? hi : lo;
}

// permute the L blocks of an 8-byte value: out[block b] = in[block b ^ r]
template <int L> __device__ __forceinline__ void xor_blocks(uint32_t &lo, uint32_t &hi, uint32_t r) {
    if constexpr (L == 8) {   // 1-byte blocks
        if (r & 1u) { lo = prmt(lo, lo, 0x2301); hi = prmt(hi, hi, 0x2301); }
        if (r & 2u) { lo = prmt(lo, lo, 0x1032); hi = prmt(hi, hi, 0x1032); }
        if (r & 4u) { const uint32_t t = lo; lo = hi; hi = t; }
    } else if constexpr (L == 4) {   // 2-byte blocks
        if (r & 1u) { lo = prmt(lo, lo, 0x1032); hi = prmt(hi, hi, 0x1032); }
        if (r & 2u) { const uint32_t t = lo; lo = hi; hi = t; }
    } else {   // L == 2: 4-byte blocks
        if (r & 1u) { const uint32_t t = lo; lo = hi; hi = t; }
    }
}

template <int T> __device__ __forceinline__ uint4 bitshuffle_vec(uint4 v, int lane) {
    constexpr int L = T / 2;
    const uint32_t r = (uint32_t)lane & (L - 1);
    const uint32_t FULL = 0xffffffffu;
    uint32_t xlo_a, xhi_a, xlo_b, xhi_b;   // X'_{2r} and X'_{2r+1}: blocks in partner order s
    if constexpr (T == 4) {
        // 4 elements per lane, 2 u16 units each; column u = unit u of the 4 elements (8 bytes)
        const uint32_t c0a = prmt(v.x, v.y, 0x5410), c0b = prmt(v.z, v.w, 0x5410);
        const uint32_t c1a = prmt(v.x, v.y, 0x7632), c1b = prmt(v.z, v.w, 0x7632);
        const uint32_t sa = r ? c0a : c1a, sb = r ? c0b : c1b;            // column of the partner
        const uint32_t ra = __shfl_xor_sync(FULL, sa, 1), rb = __shfl_xor_sync(FULL, sb, 1);
        const uint32_t oa = r ? c1a : c0a, ob = r ? c1b : c0b;            // own column r
        // block s=0: own elements, block s=1: the partner's; low bytes -> j=2r, high -> j=2r+1
        xlo_a = prmt(oa, ob, 0x6420); xhi_a = prmt(ra, rb, 0x6420);
        xlo_b = prmt(oa, ob, 0x7531); xhi_b = prmt(ra, rb, 0x7531);
    } else if constexpr (T == 8) {
        // 2 elements per lane (x,y) (z,w), 4 units each; column u = (e0.unit u, e1.unit u)
        const uint32_t c0 = prmt(v.x, v.z, 0x5410), c1 = prmt(v.x, v.z, 0x7632);
        const uint32_t c2 = prmt(v.y, v.w, 0x5410), c3 = prmt(v.y, v.w, 0x7632);
        uint32_t R[4];
        R[0] = sel4(r, c0, c1, c2, c3);
#pragma unroll
        for (int s = 1; s < 4; s++) R[s] = __shfl_xor_sync(FULL, sel4(r ^ s, c0, c1, c2, c3), s);
        xlo_a = prmt(R[0], R[1], 0x6420); xhi_a = prmt(R[2], R[3], 0x6420);
        xlo_b = prmt(R[0], R[1], 0x7531); xhi_b = prmt(R[2], R[3], 0x7531);
    } else {   // T == 16: one element per lane, 8 units
        uint32_t R[8];
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const uint32_t i = r ^ (uint32_t)s;
            const uint32_t w = sel4(i >> 1, v.x, v.y, v.z, v.w);
            const uint32_t u = (i & 1u) ? (w >> 16) : (w & 0xFFFFu);
            R[s] = s ? __shfl_xor_sync(FULL, u, s) : u;
        }
        const uint32_t t01 = prmt(R[0], R[1], 0x5140), t23 = prmt(R[2], R[3], 0x5140);
        const uint32_t t45 = prmt(R[4], R[5], 0x5140), t67 = prmt(R[6], R[7], 0x5140);
        xlo_a = prmt(t01, t23, 0x5410); xhi_a = prmt(t45, t67, 0x5410);
        xlo_b = prmt(t01, t23, 0x7632); xhi_b = prmt(t45, t67, 0x7632);
    }
    xor_blocks<L>(xlo_a, xhi_a, r);
    xor_blocks<L>(xlo_b, xhi_b, r);
    bit_transpose8(xlo_a, xhi_a);
    bit_transpose8(xlo_b, xhi_b);
    return make_uint4(xlo_a, xhi_a, xlo_b, xhi_b);
}

template <int T> __device__ __forceinline__ uint4 bitunshuffle_vec(uint4 v, int lane) {
    constexpr int L = T / 2;
    const uint32_t r = (uint32_t)lane & (L - 1);
    const uint32_t FULL = 0xffffffffu;
    uint32_t alo = v.x, ahi = v.y, blo = v.z, bhi = v.w;   // Y_{2r}, Y_{2r+1}
    bit_transpose8(alo, ahi);                              // X_{2r}: byte m = element m, byte 2r
    bit_transpose8(blo, bhi);                              // X_{2r+1}
    xor_blocks<L>(alo, ahi, r);                            // block s now belongs to lane r^s
    xor_blocks<L>(blo, bhi, r);
    if constexpr (T == 4) {
        // block = 4 elements; packet for a lane = its 4 elements' unit r: (a.byte, b.byte) x 4
        const uint32_t p0a = prmt(alo, blo, 0x5140), p0b = prmt(alo, blo, 0x7362);   // own block
        const uint32_t p1a = prmt(ahi, bhi, 0x5140), p1b = prmt(ahi, bhi, 0x7362);   // partner's
        const uint32_t qa = __shfl_xor_sync(FULL, p1a, 1), qb = __shfl_xor_sync(FULL, p1b, 1);
        // own elements e=0..3: unit r from p0, unit r^1 from q
        const uint32_t u0a = r ? qa : p0a, u0b = r ? qb : p0b;   // unit 0 of elements (0,1) (2,3)
        const uint32_t u1a = r ? p0a : qa, u1b = r ? p0b : qb;   // unit 1
        return make_uint4(prmt(u0a, u1a, 0x5410), prmt(u0a, u1a, 0x7632), prmt(u0b, u1b, 0x5410),
                          prmt(u0b, u1b, 0x7632));
    } else if constexpr (T == 8) {
        // block = 2 elements; packet s = (a.b0, b.b0, a.b1, b.b1) of block s = unit r of 2 elements
        uint32_t P[4];
        P[0] = prmt(alo, blo, 0x5140); P[1] = prmt(alo, blo, 0x7362);
        P[2] = prmt(ahi, bhi, 0x5140); P[3] = prmt(ahi, bhi, 0x7362);
        uint32_t U[4];   // U[s] = unit (r^s) of my two elements
        U[0] = P[0];
#pragma unroll
        for (int s = 1; s < 4; s++) U[s] = __shfl_xor_sync(FULL, P[s], s);
        // unit u of my elements = U[u ^ r]
        const uint32_t u0 = sel4(r, U[0], U[1], U[2], U[3]), u1 = sel4(r ^ 1u, U[0], U[1], U[2], U[3]);
        const uint32_t u2 = sel4(r ^ 2u, U[0], U[1], U[2], U[3]), u3 = sel4(r ^ 3u, U[0], U[1], U[2], U[3]);
        // element 0 = low halves of u0..u3, element 1 = high halves
        return make_uint4(prmt(u0, u1, 0x5410), prmt(u2, u3, 0x5410), prmt(u0, u1, 0x7632),
                          prmt(u2, u3, 0x7632));
    } else {   // T == 16: block = 1 element; packet s = (a.byte s, b.byte s)
        uint32_t U[8];
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const uint32_t aw = s < 4 ? alo : ahi, bw = s < 4 ? blo : bhi;
            const uint32_t pk = ((aw >> (8 * (s & 3))) & 0xFFu) | (((bw >> (8 * (s & 3))) & 0xFFu) << 8);
            U[s] = s ? __shfl_xor_sync(FULL, pk, s) : pk;
        }
        // unit u of my element = U[u ^ r]; word k = (unit 2k, unit 2k+1)
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i0 = (uint32_t)(2 * k) ^ r, i1 = (uint32_t)(2 * k + 1) ^ r;
            const uint32_t e0 = (i0 & 4u) ? sel4(i0 & 3u, U[4], U[5], U[6], U[7]) : sel4(i0 & 3u, U[0], U[1], U[2], U[3]);
            const uint32_t e1 = (i1 & 4u) ? sel4(i1 & 3u, U[4], U[5], U[6], U[7]) : sel4(i1 & 3u, U[0], U[1], U[2], U[3]);
            w[k] = (e0 & 0xFFFFu) | (e1 << 16);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int T>
__device__ __forceinline__ void run_bitshuffle_fast(const uint8_t *s, uint8_t *d, uint64_t G,
                                                    uint32_t tile0, uint32_t tstride,
                                                    bool inverse) {
    constexpr uint64_t GPT = kTileBytes / (8 * T);  // groups per tile
    const uint64_t ntiles = (G + GPT - 1) / GPT;
    for (uint64_t t = tile0; t < ntiles; t += tstride) {
        const uint64_t g0 = t * GPT;
        const uint32_t vg = (uint32_t)(G - g0 < GPT ? G - g0 : GPT);
        if constexpr (T == 2) {
            for (uint32_t g = threadIdx.x; g < vg; g += kFilterThreads) {
                const uint64_t b = (g0 + g) * 8 * T;
                if (!inverse) bitshuffle_group<T>(s + b, d + b);
                else bitunshuffle_group<T>(s + b, d + b);
            }
        } else {
            // one 16-byte vector per lane and pass; groups never straddle a warp (8*T <= 128 bytes)
            const uint32_t vbytes = vg * 8 * T;
            const uint8_t *ts = s + g0 * 8 * T;
            uint8_t *td = d + g0 * 8 * T;
            const int lane = threadIdx.x & 31;
            constexpr int NV = kTileBytes / 16 / kFilterThreads;
            uint4 v[NV];
#pragma unroll
            for (int it = 0; it < NV; it++) {
                const uint32_t b = 16u * (it * kFilterThreads + threadIdx.x);
                v[it] = b < vbytes ? ldg128_stream(ts + b) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int it = 0; it < NV; it++) {
                const uint32_t b = 16u * (it * kFilterThreads + threadIdx.x);
                const uint4 o = inverse ? bitunshuffle_vec<T>(v[it], lane) : bitshuffle_vec<T>(v[it], lane);
                if (b < vbytes) stg128_stream(td + b, o);
            }
        }
    }
}

// ---- K2, T = 8 and 16: shared-memory staged ------------------------------------------------------
// A group is 8 elements x T bytes.  The lane all-to-all above needs T/2 - 1 shuffles per 16 bytes
// and, for T = 16, runs out of registers; here the tile is staged in shared memory with the group
// rows padded (128 -> 144 bytes, 64 -> 72 bytes), one thread takes (group g, byte positions
// 4q..4q+3): 8 LDS.32, one per element (word index 36g + 4m + q resp. 18g + 2m + q: the 32 lanes of
// a warp hit 32 different banks), two 4x4 byte transposes, four 8x8 bit transposes, and 32
// contiguous output bytes.  Unshuffle is the mirror image (32 contiguous input bytes per thread,
// STS.32 into the padded rows, coalesced rows out).
template <int T> struct BitSmemCfg {
    static constexpr uint32_t kRow = 8 * T;                          // bytes of one group
    static constexpr uint32_t kPad = T == 16 ? 16 : 8;               // row padding
    static constexpr uint32_t kStride = kRow + kPad;
    static constexpr uint32_t kGroups = kTileBytes / kRow;           // groups per tile
    static constexpr uint32_t kQ = T / 4;                            // threads per group
};
constexpr uint32_t kFilterSmemBytes = BitSmemCfg<16>::kGroups * BitSmemCfg<16>::kStride;   // 18 432 >= both layouts
static_assert(BitSmemCfg<8>::kGroups * BitSmemCfg<8>::kStride <= kFilterSmemBytes, "smem");
static_assert(kStageLin <= kFilterSmemBytes, "smem");   // staged byte shuffle: the contiguous side of one tile

// tile byte b (16-byte chunk) <-> its place in the padded rows
template <int T> __device__ __forceinline__ uint32_t bit_smem_pos(uint32_t b) {
    using C = BitSmemCfg<T>;
    return (b / C::kRow) * C::kStride + (b % C::kRow);
}

template <int T>
__device__ __forceinline__ void run_bitshuffle_smem(const uint8_t *s, uint8_t *d, uint64_t G, uint32_t tile0,
                                                    uint32_t tstride, bool inverse, uint8_t *smem) {
    using C = BitSmemCfg<T>;
    const uint64_t ntiles = (G + C::kGroups - 1) / C::kGroups;
    for (uint64_t t = tile0; t < ntiles; t += tstride) {
        const uint64_t g0 = t * C::kGroups;
        const uint32_t vg = (uint32_t)(G - g0 < C::kGroups ? G - g0 : C::kGroups);
        const uint32_t vbytes = vg * C::kRow;
        const uint8_t *ts = s + g0 * C::kRow;
        uint8_t *td = d + g0 * C::kRow;
        if (!inverse) {
#pragma unroll
            for (int it = 0; it < kTileBytes / 16 / kFilterThreads; it++) {
                const uint32_t b = 16u * (it * kFilterThreads + threadIdx.x);
                if (b < vbytes) {
                    const uint4 v = ldg128_stream(ts + b);
                    uint8_t *p = smem + bit_smem_pos<T>(b);                 // 16-byte aligned for T = 16, 8 for T = 8
                    if constexpr (T == 16) *reinterpret_cast<uint4 *>(p) = v;
                    else { *reinterpret_cast<uint2 *>(p) = make_uint2(v.x, v.y); *reinterpret_cast<uint2 *>(p + 8) = make_uint2(v.z, v.w); }
                }
            }
            __syncthreads();
            for (uint32_t item = threadIdx.x; item < vg * C::kQ; item += kFilterThreads) {
                const uint32_t g = item / C::kQ, q = item % C::kQ;
                const uint8_t *row = smem + g * C::kStride + 4u * q;
                uint32_t w[8];
#pragma unroll
                for (int m = 0; m < 8; m++) w[m] = *reinterpret_cast<const uint32_t *>(row + T * m);
                uint32_t lo[4], hi[4];
                transpose4x4(w[0], w[1], w[2], w[3], lo[0], lo[1], lo[2], lo[3]);   // byte m of lo[k] = element m, byte 4q+k
                transpose4x4(w[4], w[5], w[6], w[7], hi[0], hi[1], hi[2], hi[3]);
#pragma unroll
                for (int k = 0; k < 4; k++) bit_transpose8(lo[k], hi[k]);
                uint8_t *o = td + g * C::kRow + 32u * q;                                // dst[8gT + 8j ..], j = 4q..4q+3
                stg128_stream(o, make_uint4(lo[0], hi[0], lo[1], hi[1]));
                stg128_stream(o + 16, make_uint4(lo[2], hi[2], lo[3], hi[3]));
            }
            __syncthreads();
        } else {
            for (uint32_t item = threadIdx.x; item < vg * C::kQ; item += kFilterThreads) {
                const uint32_t g = item / C::kQ, q = item % C::kQ;
                const uint8_t *in = ts + g * C::kRow + 32u * q;
                const uint4 a = ldg128_stream(in), b = ldg128_stream(in + 16);
                uint32_t lo[4] = {a.x, a.z, b.x, b.z}, hi[4] = {a.y, a.w, b.y, b.w};
#pragma unroll
                for (int k = 0; k < 4; k++) bit_transpose8(lo[k], hi[k]);             // byte m = element m, byte 4q+k
                uint32_t w[8];
                transpose4x4(lo[0], lo[1], lo[2], lo[3], w[0], w[1], w[2], w[3]);     // w[m] = element m, bytes 4q..4q+3
                transpose4x4(hi[0], hi[1], hi[2], hi[3], w[4], w[5], w[6], w[7]);
                uint8_t *row = smem + g * C::kStride + 4u * q;
#pragma unroll
                for (int m = 0; m < 8; m++) *reinterpret_cast<uint32_t *>(row + T * m) = w[m];
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < kTileBytes / 16 / kFilterThreads; it++) {
                const uint32_t b = 16u * (it * kFilterThreads + threadIdx.x);
                if (b < vbytes) {
                    const uint8_t *p = smem + bit_smem_pos<T>(b);
                    uint4 v;
                    if constexpr (T == 16) v = *reinterpret_cast<const uint4 *>(p);
                    else { const uint2 x = *reinterpret_cast<const uint2 *>(p), y = *reinterpret_cast<const uint2 *>(p + 8); v = make_uint4(x.x, x.y, y.x, y.y); }
                    stg128_stream(td + b, v);
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kFilterThreads, 6) filter_batch_kernel(FilterArgs a) {
    __shared__ __align__(16) uint8_t smem[kFilterSmemBytes];
    const uint32_t tpf = a.ft.tiles_per_frame;
    const uint32_t f = blockIdx.x / tpf, tile0 = blockIdx.x % tpf;
    if (f >= a.ft.nframes) return;
    if (a.status && a.status[f] != 0) return;
    const FrameMeta m = a.meta ? a.meta[f] : a.uniform;
    const uint64_t n = frame_len(a.ft, f), off = frame_off(a.ft, f);
    if (a.limit && (off > a.limit || n > a.limit - off)) return;   // the caller's bound was not one: K3 reports the frame
    const uint8_t *s = a.src + off;
    uint8_t *d = a.dst + off;
    const uint64_t T = m.typesize;
    const bool inverse = a.inverse != 0;
    const bool active = (m.mode == 1 || m.mode == 2) && T > 1 && n >= T;
    if (!active) {
        // identity (shuffle.go:17-19 and the default arm of ShuffleBuffer): plain copy
        if (s == d || !a.copy_inactive) return;
        for (uint64_t t = tile0; t * kTileBytes < n; t += tpf) {
            uint64_t b = t * kTileBytes, len = n - b < kTileBytes ? n - b : kTileBytes;
            cta_copy(d + b, s + b, len);
        }
        return;
    }
    const uint64_t E = n / T;
    const bool aligned = ((((uintptr_t)s) | ((uintptr_t)d)) & 15u) == 0;
    uint64_t covered;  // bytes handled by the transform; the rest is copied raw by tile 0
    if (m.mode == 1) {
        covered = E * T;
        const bool fast = aligned && (E % 16 == 0);
        if (fast && T == 4) run_shuffle_fast<4>(s, d, E, tile0, tpf, inverse, smem);
        else if (fast && T == 8) run_shuffle_fast<8>(s, d, E, tile0, tpf, inverse, smem);
        else if (fast && T == 2) run_shuffle_fast<2>(s, d, E, tile0, tpf, inverse, smem);
        else if (fast && T == 16) run_shuffle_fast<16>(s, d, E, tile0, tpf, inverse, smem);
        else if (T <= kStageMaxT) {
            // everything else up to 16 bytes per element: staged through shared memory, aligned vectors on both sides
            const uint64_t te = kStageTile / T;
            const uint64_t ntiles = (E + te - 1) / te;
            for (uint64_t t = tile0; t < ntiles; t += tpf) {
                const uint64_t e0 = t * te;
                const uint32_t ve = (uint32_t)(E - e0 < te ? E - e0 : te);
                if (!inverse) shuffle_tile_staged(s, d, E, (uint32_t)T, e0, ve, smem);
                else unshuffle_tile_staged(s, d, E, (uint32_t)T, e0, ve, smem);
            }
        } else {
            uint64_t te = kTileBytes / T;
            if (te == 0) te = 1;
            const uint64_t ntiles = (E + te - 1) / te;
            for (uint64_t t = tile0; t < ntiles; t += tpf) {
                uint64_t e0 = t * te;
                uint32_t ve = (uint32_t)(E - e0 < te ? E - e0 : te);
                shuffle_tile_generic(s, d, E, T, e0, ve, inverse);
            }
        }
    } else {
        const uint64_t G = E / 8;
        covered = G * 8 * T;
        if (aligned && T == 8) run_bitshuffle_smem<8>(s, d, G, tile0, tpf, inverse, smem);
        else if (aligned && T == 4) run_bitshuffle_fast<4>(s, d, G, tile0, tpf, inverse);
        else if (aligned && T == 2) run_bitshuffle_fast<2>(s, d, G, tile0, tpf, inverse);
        else if (aligned && T == 16) run_bitshuffle_smem<16>(s, d, G, tile0, tpf, inverse, smem);
        else {
            // items (g, j): 8 bytes each; a tile is kTileBytes / 8 items
            const uint64_t items = G * T, ipt = kTileBytes / 8;
            for (uint64_t t = tile0; t * ipt < items; t += tpf) {
                uint64_t q0 = t * ipt, vq = items - q0 < ipt ? items - q0 : ipt;
                for (uint64_t q = threadIdx.x; q < vq; q += blockDim.x) {
                    uint64_t g = (q0 + q) / T, j = (q0 + q) - g * T;
                    bitshuffle_item_generic(s, d, T, g, j, inverse);
                }
            }
        }
    }
    if (tile0 == 0)
        for (uint64_t i = covered + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

}  // namespace b2b
