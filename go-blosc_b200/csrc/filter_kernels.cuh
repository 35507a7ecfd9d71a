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

template <int T>
__device__ __forceinline__ void run_bitshuffle_fast(const uint8_t *s, uint8_t *d, uint64_t G,
                                                    uint32_t tile0, uint32_t tstride,
                                                    bool inverse) {
    constexpr uint64_t GPT = kTileBytes / (8 * T);  // groups per tile
    const uint64_t ntiles = (G + GPT - 1) / GPT;
    for (uint64_t t = tile0; t < ntiles; t += tstride) {
        const uint64_t g0 = t * GPT;
        const uint32_t vg = (uint32_t)(G - g0 < GPT ? G - g0 : GPT);
        for (uint32_t g = threadIdx.x; g < vg; g += kFilterThreads) {
            const uint64_t b = (g0 + g) * 8 * T;
            if (!inverse) bitshuffle_group<T>(s + b, d + b);
            else bitunshuffle_group<T>(s + b, d + b);
        }
    }
}

__global__ void __launch_bounds__(kFilterThreads, 6) filter_batch_kernel(FilterArgs a) {
    __shared__ __align__(16) uint8_t smem[kTileBytes];
    const uint32_t tpf = a.ft.tiles_per_frame;
    const uint32_t f = blockIdx.x / tpf, tile0 = blockIdx.x % tpf;
    if (f >= a.ft.nframes) return;
    if (a.status && a.status[f] != 0) return;
    const FrameMeta m = a.meta ? a.meta[f] : a.uniform;
    const uint64_t n = frame_len(a.ft, f), off = frame_off(a.ft, f);
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
        else {
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
        if (aligned && T == 8) run_bitshuffle_fast<8>(s, d, G, tile0, tpf, inverse);
        else if (aligned && T == 4) run_bitshuffle_fast<4>(s, d, G, tile0, tpf, inverse);
        else if (aligned && T == 2) run_bitshuffle_fast<2>(s, d, G, tile0, tpf, inverse);
        else if (aligned && T == 16) run_bitshuffle_fast<16>(s, d, G, tile0, tpf, inverse);
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
