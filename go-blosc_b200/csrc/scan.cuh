// scan.cuh -- K5: single-pass exclusive scan (decoupled look-back) of per-frame sizes.
//
// Builds the packed-offsets table of a batch -- the "bstarts" analogue north_star asks for.
// The reference has no counterpart (one frame per call).  u32 in, u64 out.
// Each CTA takes the next tile from an atomic ticket (so a tile's predecessors are always
// already running), publishes its aggregate, then walks back over predecessor descriptors
// until it meets an inclusive prefix.  Descriptor = 2 flag bits | 62-bit value.
#pragma once
#include "common.cuh"
#include "lz4_encode.cuh"

namespace b2b {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;                                // items per thread
constexpr int kScanTile = kScanThreads * kScanItems;         // 1024 items per tile
constexpr uint64_t kScanFlagAgg = 1ull << 62, kScanFlagIncl = 2ull << 62;
constexpr uint64_t kScanValMask = (1ull << 62) - 1;

enum ScanOp : int {
    kScanIdentity = 0,   // x
    kScanAlign16 = 1,    // (x + 15) & ~15     packed frames start on 16-byte boundaries
    kScanSegSlot = 2,    // scratch bytes of the frame's segment slots (lz4_encode.cuh)
    kScanSegCount = 3,   // number of 64 KiB segments of the frame
    kScanSeqSlots = 4,   // sequence records of the split decoder for a frame of capacity x
    kScanStream = 5,     // x ? x + 4 : 0     a stream of a Blosc-1 block behind its int32 size (blocks.cuh)
    kScanChunks = 6,     // ceil(x / 8192)     parse chunks of an LZ4 block of x bytes (lz4_decode2.cuh)
    kScanChunksSmall = 7 // ceil(x / 1024)     the same with 1 KiB chunks (a handful of frames: b2b.cu)
};

__device__ __forceinline__ uint64_t scan_apply(int op, uint32_t x) {
    if (op == kScanAlign16) return ((uint64_t)x + 15ull) & ~15ull;
    if (op == kScanSegSlot) return frame_slot_bytes(x);
    if (op == kScanSegCount) return seg_count(x);
    if (op == kScanSeqSlots) return (uint64_t)(x / 4u) + kSeqSlack;
    if (op == kScanStream) return x ? (uint64_t)x + 4ull : 0ull;
    if (op == kScanChunks) return ((uint64_t)x + 8191ull) / 8192ull;
    if (op == kScanChunksSmall) return ((uint64_t)x + 1023ull) / 1024ull;
    return x;
}

struct ScanWork {
    uint64_t *tile_state;  // ntiles entries, zero before launch
    uint32_t *ticket;      // one counter, zero before launch
};

__global__ void __launch_bounds__(kScanThreads)
scan_offsets_kernel(const uint32_t *__restrict__ in, uint32_t n, uint64_t *__restrict__ out,
                    uint64_t *__restrict__ total, ScanWork w, int op) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp[kScanThreads / 32];
    __shared__ uint64_t s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(w.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
    uint64_t v[kScanItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? scan_apply(op, in[base + k]) : 0ull;
        sum += v[k];
    }
    // inclusive scan of the per-thread sums across the CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t warp_off = 0, cta_total = 0;
#pragma unroll
    for (int k = 0; k < kScanThreads / 32; k++) {
        if (k < warp) warp_off += s_warp[k];
        cta_total += s_warp[k];
    }
    // publish the aggregate, then look back
    if (threadIdx.x == 0) {
        volatile uint64_t *st = w.tile_state;
        uint64_t prefix = 0;
        if (tile == 0) {
            st[0] = kScanFlagIncl | cta_total;
        } else {
            st[tile] = kScanFlagAgg | cta_total;
            __threadfence();
            int64_t p = (int64_t)tile - 1;
            while (true) {
                uint64_t d = st[p];
                if ((d >> 62) == 0) continue;  // predecessor has not published yet
                prefix += d & kScanValMask;
                if (d & kScanFlagIncl) break;
                p--;
            }
            st[tile] = kScanFlagIncl | ((prefix + cta_total) & kScanValMask);
        }
        s_prefix = prefix;
        if ((uint64_t)(tile + 1) * kScanTile >= n && total) *total = prefix + cta_total;
    }
    __syncthreads();
    uint64_t run = s_prefix + warp_off + (incl - sum);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

}  // namespace b2b
