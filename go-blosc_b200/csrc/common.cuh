// common.cuh -- shared device helpers for the b2b kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2b {

constexpr int kFilterThreads = 256;     // threads per filter CTA
constexpr int kTileBytes = 16384;       // bytes of one filter tile (one CTA pass)
constexpr int kWarp = 32;
// sequence records of the split decoder: a frame gets dst_cap / 4 + kSeqSlack of them (lz4_kernels.cuh)
constexpr uint32_t kSeqSlack = 40;

// Per-frame filter selection (uniform for a compress batch, per frame from the header on
// decompress).  mode: 0 none, 1 byte shuffle, 2 bit shuffle.
struct FrameMeta {
    uint32_t mode;
    uint32_t typesize;
};

// Frame table of a batch.  off/len may be null: then every frame is `uniform_len` bytes at
// f * uniform_len (this is also how the whole-buffer ShuffleBuffer call arrives, nframes = 1
// with a 64-bit length).
struct FrameTable {
    const uint64_t *off;
    const uint32_t *len;
    uint64_t uniform_len;
    uint32_t nframes;
    uint32_t tiles_per_frame;  // CTAs assigned to each frame; they stride over its tiles
};

__device__ __forceinline__ uint64_t frame_off(const FrameTable &t, uint32_t f) {
    return t.off ? t.off[f] : (uint64_t)f * t.uniform_len;
}
__device__ __forceinline__ uint64_t frame_len(const FrameTable &t, uint32_t f) {
    return t.len ? (uint64_t)t.len[f] : t.uniform_len;
}

// ---- 16-byte global accesses.  Streaming data is touched once: keep it out of L1. --------
// (The __host__ halves below exist only so that tests/host_kernel_check.cu can run the
// per-thread bit/byte permutation code on the CPU; the product never executes them.)
__host__ __device__ __forceinline__ uint4 ldg128_stream(const void *p) {
#ifdef __CUDA_ARCH__
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
#else
    return *reinterpret_cast<const uint4 *>(p);
#endif
}
__host__ __device__ __forceinline__ void stg128_stream(void *p, const uint4 &v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
#else
    *reinterpret_cast<uint4 *>(p) = v;
#endif
}
// plain (coherent) variants for buffers that the same kernel also writes
__device__ __forceinline__ uint4 ldg128(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void stg128(void *p, const uint4 &v) { *reinterpret_cast<uint4 *>(p) = v; }

__host__ __device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t ab = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((ab >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
#endif
}

// 4x4 byte transpose: in a=(a0..a3) b c d ; out p[k] = (a_k, b_k, c_k, d_k).  8 PRMT.
__host__ __device__ __forceinline__ void transpose4x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                             uint32_t &p0, uint32_t &p1, uint32_t &p2,
                                             uint32_t &p3) {
    uint32_t t0 = prmt(a, b, 0x5140), t1 = prmt(c, d, 0x5140);
    uint32_t t2 = prmt(a, b, 0x7362), t3 = prmt(c, d, 0x7362);
    p0 = prmt(t0, t1, 0x5410);
    p1 = prmt(t0, t1, 0x7632);
    p2 = prmt(t2, t3, 0x5410);
    p3 = prmt(t2, t3, 0x7632);
}

// 8x8 bit-matrix transpose along the ANTI-diagonal on (lo = bytes 0..3, hi = bytes 4..7):
// with byte m of the input = b[m], byte k of the output has bit (7-m) = b[m] bit (7-k),
// which is exactly the reference's MSB-first bit shuffle step (shuffle.go:184-200 and its
// inverse 261-276; the map is an involution).  Three delta swaps: 9, 18, 36.
__host__ __device__ __forceinline__ void bit_transpose8(uint32_t &lo, uint32_t &hi) {
    uint32_t t;
    t = (lo ^ (lo >> 9)) & 0x00550055u; lo ^= t ^ (t << 9);
    t = (hi ^ (hi >> 9)) & 0x00550055u; hi ^= t ^ (t << 9);
    t = (lo ^ (lo >> 18)) & 0x00003333u; lo ^= t ^ (t << 18);
    t = (hi ^ (hi >> 18)) & 0x00003333u; hi ^= t ^ (t << 18);
    t = (lo ^ (hi >> 4)) & 0x0F0F0F0Fu; lo ^= t; hi ^= t << 4;
}

// ---- warp-cooperative byte copy, any alignment of src and dst ----------------------------
// All 32 lanes call it with the same arguments.  Ranges must not overlap.  Reads stay
// inside [src, src+n).  The bulk moves as 16-byte stores aligned on dst, the source being
// re-aligned with funnel shifts when its alignment differs.
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n, int lane) {
    if (n < 96) {
        for (uint32_t i = lane; i < n; i += kWarp) dst[i] = src[i];
        return;
    }
    uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t sh = (uint32_t)((uintptr_t)src & 15u);
    uint32_t nvec;
    if (sh == 0) {
        nvec = n >> 4;
        uint32_t i = lane;
        for (; i + 3 * kWarp < nvec; i += 4 * kWarp) {
            uint4 a = ldg128(src + 16ull * i), b = ldg128(src + 16ull * (i + kWarp));
            uint4 c = ldg128(src + 16ull * (i + 2 * kWarp)), d = ldg128(src + 16ull * (i + 3 * kWarp));
            stg128(dst + 16ull * i, a); stg128(dst + 16ull * (i + kWarp), b);
            stg128(dst + 16ull * (i + 2 * kWarp), c); stg128(dst + 16ull * (i + 3 * kWarp), d);
        }
        for (; i < nvec; i += kWarp) stg128(dst + 16ull * i, ldg128(src + 16ull * i));
    } else {
        // byte-misaligned source: two aligned 16-byte loads per output vector (the second one is
        // the next lane's first: an L1 hit) and a funnel shift.  The last vectors are left to the
        // byte tail so that the second load never leaves [src, src + n).
        nvec = n >= 32 ? (n - 16) >> 4 : 0;
        const uint8_t *al = src - sh;                      // 16-byte aligned
        const uint32_t ws = sh >> 2, bs = 8u * (sh & 3u);
        auto realign = [&](const uint4 &a, const uint4 &b) -> uint4 {
            uint32_t w0, w1, w2, w3, w4;
            switch (ws) {
                case 0: w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; break;
                case 1: w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; break;
                case 2: w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; break;
                default: w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; break;
            }
            return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs),
                              __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
        };
        uint32_t i = lane;
        for (; i + kWarp < nvec; i += 2 * kWarp) {
            const uint4 a0 = ldg128(al + 16ull * i), b0 = ldg128(al + 16ull * i + 16);
            const uint4 a1 = ldg128(al + 16ull * (i + kWarp)), b1 = ldg128(al + 16ull * (i + kWarp) + 16);
            stg128(dst + 16ull * i, realign(a0, b0));
            stg128(dst + 16ull * (i + kWarp), realign(a1, b1));
        }
        for (; i < nvec; i += kWarp)
            stg128(dst + 16ull * i, realign(ldg128(al + 16ull * i), ldg128(al + 16ull * i + 16)));
    }
    for (uint32_t i = (nvec << 4) + lane; i < n; i += kWarp) dst[i] = src[i];
}

// CTA-wide variant: every warp of the CTA takes an equal slice (slices are multiples of 16
// bytes from an aligned start, so only the first warp sees the unaligned head).
__device__ __forceinline__ void cta_copy(uint8_t *dst, const uint8_t *src, uint64_t n) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > n) head = n;
    if (warp == 0) for (uint32_t i = lane; i < head; i += kWarp) dst[i] = src[i];
    dst += head; src += head; n -= head;
    // slices of 4 KiB handed out round-robin
    const uint64_t slice = 4096;
    for (uint64_t s = (uint64_t)warp * slice; s < n; s += (uint64_t)nwarps * slice) {
        uint64_t m = n - s < slice ? n - s : slice;
        warp_copy(dst + s, src + s, (uint32_t)m, lane);
    }
}

}  // namespace b2b
