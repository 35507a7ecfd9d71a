// b2b.cu -- C ABI (include/b2b.h) over the sm_100a kernels.  No CPU fallback anywhere: every
// transform, LZ4 block and frame is produced by the kernels in this directory.
#include "../../include/b2b.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges around every stage, free when no tool listens

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "filter_kernels.cuh"
#include "lz4_encode.cuh"
#include "lz4_kernels.cuh"
#include "scan.cuh"
#include "lz4_decode2.cuh"
#include "lz4_decode3.cuh"
#include "lz4_decode4.cuh"
#include "blocks.cuh"
#include "host_staging.hpp"

#include <sched.h>

using namespace b2b;

enum KernelId { K_FILTER = 0, K_ENCODE, K_DECODE, K_SCAN, K_PACK, K_INFO, K_FINALIZE, K_PARSE,
                K_BLOCKS_META, K_BLOCKS_PACK, K_BLOCKS_DECODE, K_PREP2, K_PARSE2, K_STITCH2, K_COPY2, K_LANE,
                K_JUMP_MAP, K_JUMP_ROUND, K_JUMP_GATHER, K_JUMP_LONG, K_COUNT };
static const char *const kKernelNames[K_COUNT] = {"filter_batch_kernel", "lz4_encode_kernel", "lz4_decode_kernel",
                                                  "scan_offsets_kernel", "pack_frames_kernel", "frame_info_kernel",
                                                  "finalize_frames_kernel", "lz4_parse_kernel",
                                                  "blocks_meta_kernels", "blocks_pack_kernel", "blocks_decode_kernel",
                                                  "frame_prep_kernel", "lz4_chunk_parse_kernel", "lz4_stitch_kernel",
                                                  "lz4_copy2_kernel", "lz4_lane_decode_kernel",
                                                  "lz4_jump_map_kernel", "lz4_jump_round_kernel", "lz4_jump_gather_kernel",
                                                  "lz4_jump_long_kernel"};

struct TimedLaunch { int id; cudaEvent_t a, b; };

struct b2b_ctx {
    int opt_timing = 0;                 // record CUDA events around every kernel launch
    std::vector<TimedLaunch> pending;   // not yet folded into the sums
    std::vector<cudaEvent_t> event_pool;
    double kernel_ms[K_COUNT] = {};
    uint64_t kernel_launches[K_COUNT] = {};
    int device = 0;
    int sm_count = 148;
    std::mutex mu;
    cudaStream_t stream = nullptr;     // for the host-pointer entry points
    uint8_t *arena = nullptr;          // device scratch in use, grow-only (= arenas[cur_arena])
    size_t arena_cap = 0;
    // arena 0 serves the device-pointer entry points; 1..kSlots the chunks of the host pipeline, whose
    // kernels run on one stream per slot so that small chunks overlap on the device
    uint8_t *arenas[5] = {};
    size_t arena_caps[5] = {};
    int cur_arena = 0;
    cudaStream_t s_k[4] = {};
    // device staging of the host-pointer paths: {in, out, tables} x kSlots pipeline slots, and one
    // pinned host table block per slot (offsets / lengths / status travel through it, so that no
    // copy of the pipeline ever touches pageable memory and blocks the host)
    static constexpr int kSlots = 4;
    uint8_t *hbuf[3 * kSlots] = {};
    size_t hcap[3 * kSlots] = {};
    uint8_t *ptab[kSlots] = {};
    size_t ptab_cap[kSlots] = {};
    cudaStream_t s_in = nullptr, s_out = nullptr, s_tab = nullptr;   // H2D / D2H / table streams
    cudaEvent_t ev_in_ready[kSlots] = {}, ev_in_free[kSlots] = {};
    cudaEvent_t ev_done[kSlots] = {}, ev_out_free[kSlots] = {}, ev_tab[kSlots] = {};
    int opt_quirk = 0;
    int opt_filter_ctas_per_sm = 0;
    int opt_hash_log = 0;              // 0: automatic (launch_encode)
    int opt_hash_bytes = 0;            // 0: automatic
    uint32_t opt_tune[4] = {0, 0, 0, 0};   // encoder experiment knobs (0: built-in default)
    int opt_fused_decode = -1;         // K4 variant: -1 automatic (by the batch's shape, see decompress_batch_dev_locked), 3 one lane per
                                       // frame (lz4_decode3.cuh), 0 chunk-parallel
                                       // decoder (lz4_decode2.cuh: a frame is spread over many threads), 1 the first design's fused
                                       // kernel (one warp per frame), 2 the first design's parse kernel + copy kernel (one warp per frame)
    uint64_t opt_stage_bytes = 128ull << 20;
    int opt_encode_ctas = 0;           // (option 106) persistent encoder CTAs per SM, 0 = as many as fit
    int opt_persistent_decode = 0;     // (option 105) one-warp-per-frame decoders as persistent warps that take frames from a ticket: measured
                                       // 3 % slower on one stream and neutral on two (the gain of the two streams is not a tail effect), so off
    int opt_parse_ctas = 0;            // (option 109) chunk-parse CTAs per SM (0: as many as fit)
    int opt_no_small_chunks = 0;       // (option 108) 1: 8 KiB parse chunks also for a handful of small frames
    uint32_t opt_jump_min_bytes = 0;   // (option 107) batches of at most 4 frames: frames over this size take the chunk-parallel parse (0: 192 KiB; one 256 KiB frame 1.55 -> 1.22 ms, one 64 KiB frame 0.43 -> 0.81 ms)
    int opt_decode_streams = 0;        // streams a large decompress batch is split over: 0 automatic (2), 1 none, 2..4
    static constexpr int kSide = 3;
    cudaStream_t s_side[kSide] = {};   // the extra streams of large device-pointer decompress batches
    cudaEvent_t ev_fork = nullptr, ev_join[kSide] = {};
    int opt_fuse_unshuffle = 0;        // K4 (one warp per frame): 1 = the decoding warp also un-shuffles its frame (typesize 2 / 4).
                                       // Measured neutral on C3 (decode +2.3 ms, separate pass -2.3 ms per 8 GiB: DESIGN.md section 4), so off
    int opt_host_threads = 0;          // host threads that move pageable caller memory into / out of the pinned ring (0: automatic)
    int opt_no_staging = 0;            // 1: pageable buffers go to cudaMemcpyAsync directly (synchronous, driver-staged), as in round 1
    HostStaging *staging = nullptr;    // created when the first pageable buffer arrives
    uint64_t launches = 0;
    std::string last_err;
    // arena 0 is shared by every device-pointer call: a call on another stream than the previous one waits for it
    cudaEvent_t ev_arena = nullptr;
    cudaStream_t arena_stream = nullptr;
    bool arena_busy = false;
};

namespace {

#define CU(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->last_err = std::string(#call) + ": " + cudaGetErrorString(e__);            \
            return B2B_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// Brackets one kernel launch with CUDA events on the launching stream when timing is on.
struct LaunchTimer {
    b2b_ctx *ctx; int id; cudaStream_t s; TimedLaunch t{};
    bool on;
    LaunchTimer(b2b_ctx *c, int kid, cudaStream_t st) : ctx(c), id(kid), s(st), on(c->opt_timing != 0) {
        ctx->launches++;
        ctx->kernel_launches[id]++;
        nvtxRangePushA(kKernelNames[id]);
        if (!on) return;
        auto get = [&]() { cudaEvent_t e; if (!ctx->event_pool.empty()) { e = ctx->event_pool.back(); ctx->event_pool.pop_back(); } else cudaEventCreate(&e); return e; };
        t.id = id; t.a = get(); t.b = get();
        cudaEventRecord(t.a, s);
    }
    ~LaunchTimer() {
        nvtxRangePop();
        if (!on) return;
        cudaEventRecord(t.b, s);
        ctx->pending.push_back(t);
    }
};

struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

void fold_timings(b2b_ctx *ctx) {
    for (auto &t : ctx->pending) {
        float ms = 0;
        if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess)
            ctx->kernel_ms[t.id] += ms;
        ctx->event_pool.push_back(t.a); ctx->event_pool.push_back(t.b);
    }
    ctx->pending.clear();
}

// Device-pointer calls carve their scratch from offset 0 of arena 0 and only enqueue work.  Two calls on the SAME
// stream are ordered by the stream; a call on a DIFFERENT stream than the previous one first waits for the event the
// previous call left behind, so that two streams never work in the same scratch at once (ctx->mu held).
struct ArenaGuard {
    b2b_ctx *ctx; cudaStream_t s;
    ArenaGuard(b2b_ctx *c, cudaStream_t st) : ctx(c), s(st) {
        if (ctx->cur_arena != 0) return;                   // pipeline slots own their arena and stream
        if (ctx->arena_busy && ctx->arena_stream != s) cudaStreamWaitEvent(s, ctx->ev_arena, 0);
    }
    ~ArenaGuard() {
        if (ctx->cur_arena != 0) return;
        cudaEventRecord(ctx->ev_arena, s);
        ctx->arena_stream = s; ctx->arena_busy = true;
    }
};

// bump allocator over the ctx arena
struct Arena {
    b2b_ctx *ctx;
    uint64_t used = 0;
    explicit Arena(b2b_ctx *c) : ctx(c) {}
    template <typename T> T *take(uint64_t count) {
        used = align_up(used, 256);
        T *p = reinterpret_cast<T *>(ctx->arena + used);
        used += count * sizeof(T);
        return p;
    }
};

int ensure_arena(b2b_ctx *ctx, uint64_t bytes) {
    bytes = align_up(bytes + 4096, 1 << 20);
    if (bytes <= ctx->arena_cap) return B2B_OK;
    CU(ctx, cudaDeviceSynchronize());
    if (ctx->arena) CU(ctx, cudaFree(ctx->arena));
    ctx->arena = nullptr; ctx->arena_cap = 0;
    ctx->arenas[ctx->cur_arena] = nullptr; ctx->arena_caps[ctx->cur_arena] = 0;
    CU(ctx, cudaMalloc(&ctx->arena, bytes));
    ctx->arena_cap = bytes;
    ctx->arenas[ctx->cur_arena] = ctx->arena; ctx->arena_caps[ctx->cur_arena] = bytes;
    return B2B_OK;
}

void select_arena(b2b_ctx *ctx, int i) {
    ctx->arenas[ctx->cur_arena] = ctx->arena; ctx->arena_caps[ctx->cur_arena] = ctx->arena_cap;
    ctx->cur_arena = i;
    ctx->arena = ctx->arenas[i]; ctx->arena_cap = ctx->arena_caps[i];
}

// pinned staging of pageable caller buffers (host_staging.hpp); created on first use
HostStaging *staging_of(b2b_ctx *ctx) {
    if (!ctx->staging) {
        int n = ctx->opt_host_threads;
        if (n <= 0) {
            cpu_set_t set;
            int avail = (int)std::thread::hardware_concurrency();
            if (sched_getaffinity(0, sizeof set, &set) == 0) avail = CPU_COUNT(&set);
            n = std::max(1, std::min(8, avail / 2));
        }
        ctx->staging = new (std::nothrow) HostStaging(ctx->device, n - 1, ctx->s_in, ctx->s_out);
    }
    return ctx->staging;
}

// grow-only device staging buffer i of the host-pointer entry points
int ensure_hbuf(b2b_ctx *ctx, int i, uint64_t bytes, uint8_t **out) {
    bytes = align_up(bytes + 256, 1 << 20);
    if (bytes > ctx->hcap[i]) {
        CU(ctx, cudaDeviceSynchronize());
        if (ctx->hbuf[i]) CU(ctx, cudaFree(ctx->hbuf[i]));
        ctx->hbuf[i] = nullptr; ctx->hcap[i] = 0;
        CU(ctx, cudaMalloc(&ctx->hbuf[i], bytes));
        ctx->hcap[i] = bytes;
    }
    *out = ctx->hbuf[i];
    return B2B_OK;
}

// grow-only pinned host table block of pipeline slot i
int ensure_ptab(b2b_ctx *ctx, int i, uint64_t bytes, uint8_t **out) {
    bytes = align_up(bytes + 256, 1 << 16);
    if (bytes > ctx->ptab_cap[i]) {
        CU(ctx, cudaDeviceSynchronize());
        if (ctx->ptab[i]) CU(ctx, cudaFreeHost(ctx->ptab[i]));
        ctx->ptab[i] = nullptr; ctx->ptab_cap[i] = 0;
        CU(ctx, cudaHostAlloc((void **)&ctx->ptab[i], bytes, cudaHostAllocDefault));
        ctx->ptab_cap[i] = bytes;
    }
    *out = ctx->ptab[i];
    return B2B_OK;
}

uint32_t tiles_for(uint64_t max_len, uint32_t nframes, const b2b_ctx *ctx) {
    uint64_t t = (max_len + kTileBytes - 1) / kTileBytes;
    if (t < 1) t = 1;
    if (ctx->opt_filter_ctas_per_sm > 0) {  // persistent-style grid: CTAs stride over tiles
        uint64_t want = (uint64_t)ctx->sm_count * ctx->opt_filter_ctas_per_sm;
        uint64_t per = (want + nframes - 1) / nframes;
        if (per < 1) per = 1;
        t = std::min(t, per);
    }
    const uint64_t cap = (1ull << 30) / std::max<uint32_t>(nframes, 1);
    t = std::min<uint64_t>(t, std::max<uint64_t>(cap, 1));
    return (uint32_t)t;
}

uint64_t scan_scratch_bytes(uint32_t n) {
    const uint64_t tiles = ((uint64_t)n + kScanTile - 1) / kScanTile;
    return align_up(tiles * 8 + 8, 256) + 256;
}

// exclusive scan on `s`; `work` points at scan_scratch_bytes(n) of scratch
int launch_scan(b2b_ctx *ctx, const uint32_t *d_len, uint32_t n, uint64_t *d_off, uint64_t *d_total,
                int op, uint8_t *work, cudaStream_t s) {
    if (n == 0) {
        if (d_total) CU(ctx, cudaMemsetAsync(d_total, 0, 8, s));
        return B2B_OK;
    }
    const uint32_t tiles = (n + kScanTile - 1) / kScanTile;
    const uint64_t wbytes = align_up((uint64_t)tiles * 8 + 8, 256);
    CU(ctx, cudaMemsetAsync(work, 0, wbytes, s));
    ScanWork w;
    w.tile_state = reinterpret_cast<uint64_t *>(work);
    w.ticket = reinterpret_cast<uint32_t *>(work + (uint64_t)tiles * 8);
    { LaunchTimer lt(ctx, K_SCAN, s); scan_offsets_kernel<<<tiles, kScanThreads, 0, s>>>(d_len, n, d_off, d_total, w, op); }
    CU(ctx, cudaGetLastError());
    return B2B_OK;
}

int launch_filter(b2b_ctx *ctx, const uint8_t *src, uint8_t *dst, const uint64_t *off,
                  const uint32_t *len, uint64_t uniform_len, uint32_t nframes, uint64_t max_len,
                  const FrameMeta *meta, FrameMeta uniform, const uint32_t *status, int inverse,
                  cudaStream_t s, uint64_t limit = 0) {
    if (nframes == 0) return B2B_OK;
    FilterArgs a;
    a.src = src; a.dst = dst;
    a.ft.off = off; a.ft.len = len; a.ft.uniform_len = uniform_len; a.ft.nframes = nframes;
    a.ft.tiles_per_frame = tiles_for(max_len, nframes, ctx);
    a.meta = meta; a.uniform = uniform; a.status = status; a.inverse = inverse;
    a.copy_inactive = 1; a.limit = limit;
    const uint64_t grid = (uint64_t)nframes * a.ft.tiles_per_frame;
    { LaunchTimer lt(ctx, K_FILTER, s); filter_batch_kernel<<<(unsigned)grid, kFilterThreads, 0, s>>>(a); }
    CU(ctx, cudaGetLastError());
    return B2B_OK;
}

// unshuffled: the input did not go through a byte / bit shuffle (B2B_NOSHUFFLE, typesize <= 1, the raw-block API)
int launch_encode(b2b_ctx *ctx, const EncodeArgs &e, bool unshuffled, cudaStream_t s) {
    // automatic policy: typed arrays behind a shuffle keep the small table and the 4-byte hash (short matches
    // count there and resident warps are what the kernel lives on); unshuffled input gets 5 hashed bytes and a
    // table that reaches four times as far (text 1.17 -> 1.08, low-entropy int16 1.42 -> 1.05 of the oracle's size)
    int hb = ctx->opt_hash_bytes ? ctx->opt_hash_bytes : (unshuffled ? 5 : 4);
    int hl = ctx->opt_hash_log ? ctx->opt_hash_log : (unshuffled ? 12 : kHashLogDefault);
    if (hb == 5 && hl < 11) hl = 11;
    if (hb == 6 && hl < 12) hl = 12;
    const uint64_t warps = (uint64_t)e.nframes * e.segs_grid;
    const size_t smem = (size_t)kEncWarps * sizeof(uint32_t) << hl;
    // persistent CTAs: as many as fit on the device (warps pull items from the ticket)
    uint64_t per_sm = hl <= 10 ? 7 : hl == 11 ? 5 : hl == 12 ? 2 : 1;   // __launch_bounds__ of the kernel
    if (ctx->opt_encode_ctas > 0 && (uint64_t)ctx->opt_encode_ctas < per_sm) per_sm = (uint64_t)ctx->opt_encode_ctas;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((warps + kEncWarps - 1) / kEncWarps,
                                                                     (uint64_t)ctx->sm_count * per_sm));
    CU(ctx, cudaMemsetAsync(e.ticket, 0, 8, s));
    auto go = [&](auto kernel) -> int {
        if (smem > 48 * 1024)
            CU(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { LaunchTimer lt(ctx, K_ENCODE, s); kernel<<<grid, kEncThreads, smem, s>>>(e); }
        CU(ctx, cudaGetLastError());
        return B2B_OK;
    };
    if (e.phase_mask && hb == 4) {   // bit-shuffled input: the place inside the group is part of the key
        switch (hl) {
            case 10: return go(lz4_encode_kernel<10, 4, true>);
            case 11: return go(lz4_encode_kernel<11, 4, true>);
            case 12: return go(lz4_encode_kernel<12, 4, true>);
            default: return go(lz4_encode_kernel<13, 4, true>);
        }
    }
    switch (hb * 100 + hl) {
        case 410: return go(lz4_encode_kernel<10, 4, false>);
        case 411: return go(lz4_encode_kernel<11, 4, false>);
        case 412: return go(lz4_encode_kernel<12, 4, false>);
        case 413: return go(lz4_encode_kernel<13, 4, false>);
        case 511: return go(lz4_encode_kernel<11, 5, false>);
        case 512: return go(lz4_encode_kernel<12, 5, false>);
        case 513: return go(lz4_encode_kernel<13, 5, false>);
        case 612: return go(lz4_encode_kernel<12, 6, false>);
        case 613: return go(lz4_encode_kernel<13, 6, false>);
        default: return B2B_EINVAL;
    }
}

FrameMeta uniform_meta(int mode, int64_t typesize) {
    FrameMeta m; m.mode = 0; m.typesize = 0;
    if ((mode == B2B_SHUFFLE || mode == B2B_BITSHUFFLE) && typesize > 1 &&
        typesize <= 0xFFFFFFFFll) {
        m.mode = (uint32_t)mode; m.typesize = (uint32_t)typesize;
    }
    return m;
}

// ---- device-pointer cores (ctx->mu held by the caller) ----------------------------------
uint64_t comp_scratch_bytes(uint64_t total_src, uint32_t nframes) {
    // every segment slot holds the LZ4 worst case of 64 KiB; a frame's last, partial segment
    // still takes a whole slot when the frame has more than one
    const uint64_t multi = std::min<uint64_t>(nframes, total_src / kSegBytes + 1);
    return total_src / kSegBytes * kSegSlot + (uint64_t)kSegSlot + multi * kSegSlot + 64ull * nframes +
           total_src / 255 + 4096;
}

// raw_block: b2b_lz4_block_compress -- no filter, no header, no memcpy substitution
int compress_batch_dev_locked(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                              const uint32_t *d_src_len, uint32_t nframes, uint64_t total_src,
                              uint32_t max_len, int shuffle, int64_t typesize, void *d_dst,
                              uint64_t dst_cap, uint64_t *d_frame_off, uint32_t *d_frame_len,
                              uint32_t *d_status, uint64_t *d_total_out, cudaStream_t s,
                              bool raw_block = false, uint64_t *d_index = nullptr, uint32_t segs_per_frame = 0) {
    NvtxRange nvtx_range("b2b.compress_batch_dev");
    ArenaGuard arena_guard(ctx, s);
    if (nframes == 0) {
        if (d_total_out) CU(ctx, cudaMemsetAsync(d_total_out, 0, 8, s));
        return B2B_OK;
    }
    if (!d_src || !d_src_off || !d_src_len || !d_dst || !d_frame_off || !d_frame_len || !d_status)
        return B2B_EINVAL;
    if (((uintptr_t)d_dst & 15u) != 0) return B2B_EINVAL;
    if (!raw_block && dst_cap < total_src + 31ull * nframes) return B2B_EDST_TOO_SMALL;
    if (typesize <= 0) typesize = 1;                                   // blosc.go:274-276
    const FrameMeta fm = raw_block ? FrameMeta{0, 0} : uniform_meta(shuffle, typesize);
    const bool filtered = fm.mode != 0;
    const uint32_t shuffle_flag = raw_block ? 0 : shuffle == B2B_SHUFFLE ? B2B_FLAG_SHUFFLE
                                : shuffle == B2B_BITSHUFFLE ? B2B_FLAG_BITSHUFFLE : 0;  // blosc.go:348-353

    // scratch layout
    const uint64_t comp_bytes = comp_scratch_bytes(total_src, nframes);
    const uint64_t max_segs_total = total_src / kSegBytes + nframes + 1;
    const uint64_t need = (filtered ? align_up(total_src + 64, 256) : 0) + align_up(comp_bytes, 256) +
                          align_up(16 * max_segs_total, 256) + align_up(16 * max_segs_total, 256) +
                          8 * align_up(8ull * nframes, 256) + 3 * scan_scratch_bytes(nframes) + 8192;
    int rc = ensure_arena(ctx, need);
    if (rc) return rc;
    Arena ar(ctx);
    uint8_t *d_shuf = filtered ? ar.take<uint8_t>(total_src + 64) : nullptr;
    uint8_t *d_comp = ar.take<uint8_t>(comp_bytes);
    SegMeta *d_meta = ar.take<SegMeta>(max_segs_total);
    SegPlace *d_place = ar.take<SegPlace>(max_segs_total);
    uint64_t *d_comp_off = ar.take<uint64_t>(nframes);
    uint64_t *d_seg_base = ar.take<uint64_t>(nframes);
    uint32_t *d_comp_len = ar.take<uint32_t>(nframes);
    uint32_t *d_flags = ar.take<uint32_t>(nframes);
    uint32_t *d_final_ll = ar.take<uint32_t>(nframes);
    uint32_t *d_final_off = ar.take<uint32_t>(nframes);
    uint8_t *scan_a = ar.take<uint8_t>(scan_scratch_bytes(nframes));
    uint8_t *scan_b = ar.take<uint8_t>(scan_scratch_bytes(nframes));
    uint8_t *scan_c = ar.take<uint8_t>(scan_scratch_bytes(nframes));
    unsigned long long *d_ticket = ar.take<unsigned long long>(4);

    const uint8_t *in = static_cast<const uint8_t *>(d_src);
    if (filtered) {
        rc = launch_filter(ctx, in, d_shuf, d_src_off, d_src_len, 0, nframes, max_len, nullptr, fm,
                           nullptr, 0, s, total_src);
        if (rc) return rc;
        in = d_shuf;
    }
    rc = launch_scan(ctx, d_src_len, nframes, d_comp_off, nullptr, kScanSegSlot, scan_a, s);
    if (rc) return rc;
    rc = launch_scan(ctx, d_src_len, nframes, d_seg_base, nullptr, kScanSegCount, scan_b, s);
    if (rc) return rc;

    uint64_t segs_grid = std::max<uint64_t>(1, ((uint64_t)max_len + kSegBytes - 1) / kSegBytes);
    segs_grid = std::min<uint64_t>(segs_grid, std::max<uint64_t>(1, (1ull << 30) / nframes));

    EncodeArgs e;
    e.in = in; e.src_off = d_src_off; e.src_len = d_src_len; e.nframes = nframes;
    e.segs_grid = (uint32_t)segs_grid; e.comp = d_comp; e.comp_off = d_comp_off;
    e.seg_base = d_seg_base; e.meta = d_meta; e.ticket = d_ticket;
    for (int i = 0; i < 4; i++) e.tune[i] = ctx->opt_tune[i];
    e.independent = d_index ? 1u : 0u;
    e.planes = fm.mode == 1 ? fm.typesize : 0u;
    // bit shuffle: the place inside the 8 * typesize group is part of the hash key (lz4_encode.cuh, enc_hash)
    e.phase_mask = (fm.mode == 2 && (fm.typesize & (fm.typesize - 1)) == 0 && fm.typesize <= 512) ? 8u * fm.typesize - 1u : 0u;
    e.comp_cap = comp_bytes; e.seg_cap = max_segs_total; e.src_cap = filtered ? total_src : ~0ull;
    rc = launch_encode(ctx, e, !filtered, s);
    if (rc) return rc;

    FinalizeArgs fa;
    fa.src_len = d_src_len; fa.seg_base = d_seg_base; fa.meta = d_meta; fa.place = d_place;
    fa.nframes = nframes; fa.shuffle_flag = shuffle_flag; fa.keep_raw = raw_block ? 1 : 0;
    fa.comp_len = d_comp_len; fa.frame_len = d_frame_len; fa.flags = d_flags;
    fa.final_ll = d_final_ll; fa.final_off = d_final_off; fa.status = d_status;
    fa.index = d_index; fa.segs_per_frame = segs_per_frame;
    fa.comp_off = d_comp_off; fa.comp_cap = comp_bytes; fa.seg_cap = max_segs_total;
    fa.src_off = d_src_off; fa.src_cap = filtered ? total_src : ~0ull;
    { LaunchTimer lt(ctx, K_FINALIZE, s); finalize_frames_kernel<<<(unsigned)(((uint64_t)nframes * 32 + 127) / 128), 128, 0, s>>>(fa); }
    CU(ctx, cudaGetLastError());

    rc = launch_scan(ctx, d_frame_len, nframes, d_frame_off, d_total_out, kScanAlign16, scan_c, s);
    if (rc) return rc;

    PackArgs p;
    p.in = in;
    p.raw = ctx->opt_quirk ? static_cast<const uint8_t *>(d_src) : in;   // SURVEY F4 policy
    p.src_off = d_src_off; p.src_len = d_src_len; p.comp = d_comp; p.comp_off = d_comp_off;
    p.seg_base = d_seg_base; p.meta = d_meta; p.place = d_place; p.comp_len = d_comp_len;
    p.flags = d_flags; p.final_ll = d_final_ll; p.final_off = d_final_off; p.status = d_status;
    p.frame_off = d_frame_off; p.dst = static_cast<uint8_t *>(d_dst);
    p.nframes = nframes; p.segs_grid = (uint32_t)segs_grid;
    p.codec = B2B_LZ4; p.typesize_u8 = (uint32_t)(uint8_t)typesize;     // blosc.go:362
    p.header = raw_block ? 0 : 1; p.dst_cap = dst_cap;
    { LaunchTimer lt(ctx, K_PACK, s);
      pack_frames_kernel<<<(unsigned)((uint64_t)nframes * segs_grid), kFilterThreads, 0, s>>>(p); }
    CU(ctx, cudaGetLastError());
    return B2B_OK;
}

int decompress_batch_dev_locked(b2b_ctx *ctx, const void *d_frames, const uint64_t *d_frame_off,
                                const uint32_t *d_frame_len, uint32_t nframes,
                                int64_t typesize_override, void *d_dst, const uint64_t *d_dst_off,
                                const uint32_t *d_dst_cap, uint64_t total_dst, uint32_t max_orig,
                                uint32_t *d_out_len, uint32_t *d_status, cudaStream_t s,
                                const uint64_t *d_index = nullptr, uint32_t segs_per_frame = 0) {
    NvtxRange nvtx_range("b2b.decompress_batch_dev");
    ArenaGuard arena_guard(ctx, s);
    if (nframes == 0) return B2B_OK;
    if (!d_frames || !d_frame_off || !d_frame_len || !d_dst || !d_dst_off || !d_dst_cap ||
        !d_out_len || !d_status)
        return B2B_EINVAL;
    const bool indexed = d_index && segs_per_frame;
    // Automatic choice: the first design gives every frame ONE warp, which is the faster arrangement when a batch
    // has thousands of small frames (C3: 32 768 x 256 KiB) and a poor one when frames are large or few (a 2 MiB
    // frame keeps one warp busy for milliseconds while the rest of the device idles; C5 mixed: 138 -> 298 GB/s
    // with the chunk-parallel decoder, whose parse runs on one thread per 8 KiB of stream and whose copy stage
    // gives a frame a whole CTA).
    int variant = ctx->opt_fused_decode;
    // A third arrangement gives every frame ONE LANE (lz4_decode3.cuh): 32 scalar decoders in lockstep per warp.  It
    // needs tens of thousands of frames to fill the device and each of them is decoded at the latency of one
    // thread, so it only pays for very many very small frames (2^20 frames of 1 KiB: 222 against 158 GB/s; 4 KiB: 367
    // against 348; 16 KiB: 297 against 590; C3's 256 KiB frames: 58 ms against 13.7 ms per 8 GiB).
    // A fourth arrangement (lz4_decode4.cuh) takes the ORDER out of a frame: every output byte gets a source index and
    // pointer jumping resolves the match chains, so one large frame is spread over the whole device instead of one CTA
    // (one 256 MiB frame: 0.35 GB/s in stream order).  It costs 4 bytes of scratch and tens of bytes of traffic per output
    // byte, so it is only chosen when the frames are too few to fill the device in stream order.
    const bool want_jump = variant == 4;
    if (variant == 4) variant = 0;
    // (a handful of frames is latency, not throughput: there the chunk-parallel parse + pointer jumping already pays for
    // frames over opt_jump_min_bytes)
    const uint32_t v2_min = nframes <= 4 ? (ctx->opt_jump_min_bytes ? ctx->opt_jump_min_bytes : (192u << 10)) : (512u << 10);
    if (variant < 0) variant = max_orig > v2_min ? 0 : (max_orig <= 4096u && nframes >= 65536u) ? 3 : 2;
    const bool v2 = !indexed && variant == 0;
    // (tools/few_frames_probe.py: 64 frames of 16 MiB 47.2 -> 13.5 ms, 64 x 1 MiB 3.95 -> 2.42 ms: the tile engine gives a frame
    // 0.35 GB/s, this engine gives the batch 30-80 GB/s, so it wins while the output is under ~170 largest frames)
    const bool jump = v2 && nframes <= 256 && total_dst < 0xFFFF0000ull &&
                      (want_jump || (ctx->opt_fused_decode < 0 && total_dst <= 128ull * max_orig));
    const uint32_t jump_bpf = max_orig / kJumpBlock + 1;
    const uint32_t jump_long_cap = (uint32_t)(total_dst / kJumpLong) + nframes + 16;
    const bool split = !indexed && variant == 2;
    const bool lanes = !indexed && variant == 3;
    const uint64_t nrec_max = total_dst / 4 + (uint64_t)(kSeqSlack + 1) * nframes + 64;   // sum of dst_cap / 4 + slack
    // chunk-parallel decoder: an LZ4 block that decodes to n bytes has at most n + n / 255 + 16 bytes (longer ones
    // are refused by the prep kernel), so the chunks of a batch are bounded by its output size
    // (a handful of frames of a few MiB: 1 KiB chunks -- the parse is one thread per chunk and its slowest chunk is the
    // latency of the call: one 1 MiB frame 1.35 -> see DESIGN section 4; large frames keep 8 KiB, the stitch walks the chunks)
    const uint32_t cshift = (jump && nframes <= 4 && max_orig <= (32u << 20) && !ctx->opt_no_small_chunks) ? kChunkShiftSmall : kChunkShift;
    const uint64_t cbytes = 1ull << cshift, cslot = chunk_slot_records(cshift);
    const uint64_t table_chunks = total_dst / cbytes + total_dst / (255ull * cbytes) + 2ull * nframes + 16;
    const uint64_t need = align_up(total_dst + 64, 256) + align_up(8ull * nframes, 256) + align_up(4ull * nframes, 256) + 8192 +
                          (split ? align_up(8 * nrec_max, 256) + align_up(8ull * nframes, 256) +
                                   align_up(4ull * nframes, 256) + scan_scratch_bytes(nframes) + 1024 : 0) +
                          (v2 ? align_up(sizeof(FrameDec) * (uint64_t)nframes, 256) + 3 * align_up(4ull * nframes, 256) +
                                align_up(8ull * nframes, 256) + scan_scratch_bytes(nframes) +
                                align_up(8ull * cslot * (table_chunks + nframes), 256) + 2 * align_up(32ull * table_chunks, 256) +
                                align_up(4ull * table_chunks + 16, 256) + 2048 : 0) +
                          (jump ? align_up(4ull * total_dst + 64, 256) + 2 * align_up(4ull * nframes, 256) +
                                  align_up(sizeof(JumpLong) * (uint64_t)jump_long_cap, 256) + 512 +
                                  align_up((uint64_t)nframes * jump_bpf, 256) + 1024 : 0);
    int rc = ensure_arena(ctx, need);
    if (rc) return rc;
    Arena ar(ctx);
    uint8_t *d_stage = ar.take<uint8_t>(total_dst + 64);
    FrameMeta *d_meta = ar.take<FrameMeta>(nframes);
    uint32_t *d_cap_eff = ar.take<uint32_t>(nframes);
    unsigned long long *d_ticket = ar.take<unsigned long long>(4);
    unsigned long long *d_tickets = ar.take<unsigned long long>(2 * (1 + b2b_ctx::kSide));   // parse / copy tickets of every part
    clip_caps_kernel<<<(nframes + 255) / 256, 256, 0, s>>>(d_dst_off, d_dst_cap, total_dst, nframes, d_cap_eff);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    d_dst_cap = d_cap_eff;
    uint64_t *d_table = nullptr, *d_table_off = nullptr;
    uint32_t *d_nrec = nullptr;
    uint8_t *scan_t = nullptr;
    if (split) {
        d_table = ar.take<uint64_t>(nrec_max);
        d_table_off = ar.take<uint64_t>(nframes);
        d_nrec = ar.take<uint32_t>(nframes);
        scan_t = ar.take<uint8_t>(scan_scratch_bytes(nframes));
    }

    DecodeArgs a;
    a.frames = static_cast<const uint8_t *>(d_frames); a.frame_off = d_frame_off;
    a.frame_len = d_frame_len; a.nframes = nframes; a.typesize_override = typesize_override;
    a.dst = static_cast<uint8_t *>(d_dst); a.scratch = d_stage; a.dst_off = d_dst_off;
    a.dst_cap = d_dst_cap; a.out_len = d_out_len; a.status = d_status; a.meta = d_meta;
    a.table = nullptr; a.table_off = nullptr; a.nrec = nullptr; a.only = nullptr; a.fuse_unshuffle = ctx->opt_fuse_unshuffle ? 1u : 0u; a.ticket = nullptr;
    if (v2) {
        // prep -> K5 (chunks per frame) -> one thread per chunk parses -> one thread per frame stitches ->
        // one CTA per frame copies (lz4_decode2.cuh); frames the table has no room for (output slots that
        // overlap in dst) go through the first design's fused kernel afterwards
        FrameDec *d_fd = ar.take<FrameDec>(nframes);
        uint32_t *d_plen = ar.take<uint32_t>(nframes);
        uint32_t *d_last = ar.take<uint32_t>(nframes);
        uint32_t *d_fallback = ar.take<uint32_t>(nframes);
        uint64_t *d_chunk_base = ar.take<uint64_t>(nframes);
        uint64_t *d_total_chunks = ar.take<uint64_t>(1);
        uint8_t *scan_c = ar.take<uint8_t>(scan_scratch_bytes(nframes));
        uint2 *d_rec = ar.take<uint2>(cslot * (table_chunks + nframes));   // + one spare slot per frame (warp stitch)
        ChunkMeta *d_cmeta = ar.take<ChunkMeta>(table_chunks);
        ChunkDesc *d_cdesc = ar.take<ChunkDesc>(table_chunks);
        uint32_t *d_dead = ar.take<uint32_t>(table_chunks + 4);      // + the parse ticket behind it
        Prep2Args pa;
        pa.frames = a.frames; pa.frame_off = d_frame_off; pa.frame_len = d_frame_len; pa.dst_cap = d_dst_cap;
        pa.nframes = nframes; pa.typesize_override = typesize_override; pa.fd = d_fd; pa.plen_eff = d_plen;
        pa.out_len = d_out_len; pa.status = d_status; pa.meta = d_meta;
        pa.keep_sparse = jump ? 1u : 0u;
        { LaunchTimer lt(ctx, K_PREP2, s); frame_prep_kernel<<<(nframes + 127) / 128, 128, 0, s>>>(pa); }
        CU(ctx, cudaGetLastError());
        rc = launch_scan(ctx, d_plen, nframes, d_chunk_base, d_total_chunks, cshift == kChunkShift ? kScanChunks : kScanChunksSmall, scan_c, s);
        if (rc) return rc;
        Parse2Args pp;
        pp.frames = a.frames; pp.frame_off = d_frame_off; pp.fd = d_fd; pp.nframes = nframes;
        pp.chunk_base = d_chunk_base; pp.total_chunks = d_total_chunks; pp.table = d_rec; pp.meta = d_cmeta;
        pp.table_chunks = table_chunks; pp.chunk_shift = cshift;
        pp.dead = d_dead; pp.ticket = reinterpret_cast<unsigned long long *>(d_dead + ((table_chunks + 1) & ~1ull));
        CU(ctx, cudaMemsetAsync(d_dead, 0, 4ull * (table_chunks + 4), s));
        const unsigned pgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((table_chunks + kParse2Threads - 1) / kParse2Threads,
                                                                          (uint64_t)ctx->sm_count * (ctx->opt_parse_ctas > 0 ? std::min(ctx->opt_parse_ctas, B2B_PARSE2_CTAS) : B2B_PARSE2_CTAS)));
        { LaunchTimer lt(ctx, K_PARSE2, s); lz4_chunk_parse_kernel<<<pgrid, kParse2Threads, 0, s>>>(pp); }
        CU(ctx, cudaGetLastError());
        {   // mis-speculated chunks are repaired all at once, on the exit of the chunk before them (lz4_decode2.cuh)
            Repair2Args ra;
            ra.frames = a.frames; ra.frame_off = d_frame_off; ra.fd = d_fd; ra.nframes = nframes; ra.chunk_base = d_chunk_base;
            ra.total_chunks = d_total_chunks; ra.table = d_rec; ra.meta = d_cmeta; ra.table_chunks = table_chunks; ra.chunk_shift = cshift;
            LaunchTimer lt(ctx, K_STITCH2, s);
            lz4_chunk_repair_kernel<<<(unsigned)((table_chunks + 127) / 128), 128, 0, s>>>(ra);
        }
        CU(ctx, cudaGetLastError());
        Stitch2Args sa;
        sa.frames = a.frames; sa.frame_off = d_frame_off; sa.fd = d_fd; sa.nframes = nframes; sa.chunk_base = d_chunk_base;
        sa.table = d_rec; sa.meta = d_cmeta; sa.desc = d_cdesc; sa.last_chunk = d_last; sa.fallback = d_fallback;
        sa.table_chunks = table_chunks; sa.chunk_shift = cshift; sa.scratch = d_rec + cslot * table_chunks;
        // few frames: a warp per frame that adopts 32 chunks per step (one 1 GiB frame: 40 ms with one thread)
        { LaunchTimer lt(ctx, K_STITCH2, s);
          if (jump || nframes <= 256) lz4_stitch_warp_kernel<<<(nframes * 32 + 63) / 64, 64, 0, s>>>(sa);
          else lz4_stitch_kernel<<<(nframes + 63) / 64, 64, 0, s>>>(sa); }
        CU(ctx, cudaGetLastError());
        uint32_t *d_jump_state = nullptr;
        if (jump) {
            // select -> map (+ long runs) -> check -> jump rounds -> gather; what it refuses (state 2) is the tile engine's
            JumpArgs ja;
            ja.frames = a.frames; ja.frame_off = d_frame_off; ja.fd = d_fd; ja.nframes = nframes; ja.dst = a.dst; ja.scratch = d_stage;
            ja.dst_off = d_dst_off; ja.chunk_base = d_chunk_base; ja.total_chunks = d_total_chunks; ja.table_chunks = table_chunks;
            ja.desc = d_cdesc; ja.last_chunk = d_last; ja.table = d_rec; ja.fallback = d_fallback;
            ja.S = ar.take<uint32_t>(total_dst + 16);
            ja.state = d_jump_state = ar.take<uint32_t>(nframes);
            ja.total = ar.take<uint32_t>(nframes);
            ja.longq = ar.take<JumpLong>(jump_long_cap); ja.long_cap = jump_long_cap;
            ja.nlong = ar.take<uint32_t>(1 + kJumpRounds); ja.changed = ja.nlong + 1;
            ja.blockdone = ar.take<uint8_t>((uint64_t)nframes * jump_bpf); ja.blocks_per_frame = jump_bpf;
            ja.blocks_grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(jump_bpf, (uint64_t)ctx->sm_count * 8 / nframes));
            ja.out_len = d_out_len; ja.status = d_status; ja.meta = d_meta; ja.chunk_shift = cshift;
            CU(ctx, cudaMemsetAsync(ja.blockdone, 0, (uint64_t)nframes * jump_bpf, s));
            lz4_jump_select_kernel<<<(nframes + 63) / 64, 64, 0, s>>>(ja);
            ctx->launches++;
            CU(ctx, cudaGetLastError());
            const unsigned mgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(table_chunks, (uint64_t)ctx->sm_count * 8));
            { LaunchTimer lt(ctx, K_JUMP_MAP, s); lz4_jump_map_kernel<<<mgrid, kJumpThreads, 0, s>>>(ja); }
            { LaunchTimer lt(ctx, K_JUMP_LONG, s);
              lz4_jump_long_kernel<<<(unsigned)ctx->sm_count * 4, kJumpThreads, 0, s>>>(ja);
              lz4_jump_check_kernel<<<(nframes + 63) / 64, 64, 0, s>>>(ja); }
            ctx->launches += 1;
            CU(ctx, cudaGetLastError());
            { LaunchTimer lt(ctx, K_JUMP_ROUND, s);
              // a chain is shorter than the frame: ceil(log2(largest frame)) rounds resolve every one (the unused ones are not launched)
              uint32_t rounds = 1;
              while (rounds < kJumpRounds && (1ull << rounds) < (uint64_t)max_orig) rounds++;
              for (uint32_t r = 0; r < rounds; r++) lz4_jump_round_kernel<<<nframes * ja.blocks_grid, kJumpThreads, 0, s>>>(ja, r);
              ctx->launches += rounds - 1; }
            CU(ctx, cudaGetLastError());
            { LaunchTimer lt(ctx, K_JUMP_GATHER, s); lz4_jump_gather_kernel<<<nframes * ja.blocks_grid, kJumpThreads, 0, s>>>(ja); }
            CU(ctx, cudaGetLastError());
        }
        Copy2Args ca;
        ca.jump_state = d_jump_state; ca.chunk_shift = cshift;
        ca.frames = a.frames; ca.frame_off = d_frame_off; ca.fd = d_fd; ca.nframes = nframes; ca.dst = a.dst; ca.scratch = d_stage;
        ca.dst_off = d_dst_off; ca.chunk_base = d_chunk_base; ca.desc = d_cdesc; ca.last_chunk = d_last; ca.table = d_rec;
        ca.fallback = d_fallback; ca.out_len = d_out_len; ca.status = d_status; ca.meta = d_meta;
        // stored (memcpy-flag) frames are split over gridDim.y CTAs: about one per MiB of the largest frame
        const unsigned ny = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(64, ((uint64_t)max_orig + (1u << 20) - 1) >> 20));
        { LaunchTimer lt(ctx, K_COPY2, s); lz4_copy2_kernel<<<dim3(nframes, ny), kCopy2Threads, 0, s>>>(ca); }
        CU(ctx, cudaGetLastError());
        a.only = d_fallback;
        { LaunchTimer lt(ctx, K_DECODE, s);
          lz4_decode_kernel<false><<<(nframes + kCodecWarps - 1) / kCodecWarps, kCodecThreads, 0, s>>>(a); }
    } else if (d_index && segs_per_frame) {
        // one warp per (frame, segment): sub-streams between the index entries decode independently
        IndexedDecodeArgs ia; ia.d = a; ia.index = d_index; ia.segs_per_frame = segs_per_frame; ia.ticket = d_ticket;
        const uint64_t items = (uint64_t)nframes * segs_per_frame;
        CU(ctx, cudaMemsetAsync(d_status, 0, 4ull * nframes, s));
        CU(ctx, cudaMemsetAsync(d_ticket, 0, 8, s));
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((items + kCodecWarps - 1) / kCodecWarps,
                                                                         (uint64_t)ctx->sm_count * 8));
        { LaunchTimer lt(ctx, K_DECODE, s);
          lz4_decode_indexed_kernel<<<grid, kCodecThreads, 0, s>>>(ia); }
        CU(ctx, cudaGetLastError());
        index_finish_kernel<<<(nframes + 127) / 128, 128, 0, s>>>(a);
        ctx->launches++;
    } else if (split) {
        // parse kernel (28 registers, 48 warps per SM) -> one 8-byte record per sequence -> copy kernel
        rc = launch_scan(ctx, d_dst_cap, nframes, d_table_off, nullptr, kScanSeqSlots, scan_t, s);
        if (rc) return rc;
        a.table = d_table; a.table_off = d_table_off; a.nrec = d_nrec;
        // Large batches run as TWO halves on two streams (the caller's and an internal one, forked and joined by
        // events): every kernel here ends on a tail of warps that still walk their frames while most of the device
        // idles, and the other half's kernels fill it (C3, 8 GiB: decompress 16.7 -> 15.3 ms).  The halves share the
        // stage buffer and the record table (their frames' slots are disjoint); all tables are per frame.
        const int want = ctx->opt_decode_streams == 0 ? 2 : ctx->opt_decode_streams;
        const int parts = (nframes >= 2048 && ctx->cur_arena == 0) ? std::min(want, 1 + b2b_ctx::kSide) : 1;
        const bool two = parts > 1;
        if (two) {
            CU(ctx, cudaEventRecord(ctx->ev_fork, s));
            for (int i = 0; i + 1 < parts; i++) CU(ctx, cudaStreamWaitEvent(ctx->s_side[i], ctx->ev_fork, 0));
        }
        for (int part = 0; part < parts; part++) {
            const uint32_t f0 = (uint32_t)(((uint64_t)nframes * part / parts + 3u) & ~3ull);
            const uint32_t f1 = part + 1 == parts ? nframes : (uint32_t)(((uint64_t)nframes * (part + 1) / parts + 3u) & ~3ull);
            const uint32_t n = f1 - f0;
            cudaStream_t sp = part ? ctx->s_side[part - 1] : s;
            ParseArgs pa;
            pa.frames = a.frames; pa.frame_off = d_frame_off + f0; pa.frame_len = d_frame_len + f0; pa.dst_cap = d_dst_cap + f0;
            pa.nframes = n; pa.table = d_table; pa.table_off = d_table_off + f0; pa.nrec = d_nrec + f0; pa.table_cap = nrec_max;
            DecodeArgs ap = a;
            ap.frame_off = d_frame_off + f0; ap.frame_len = d_frame_len + f0; ap.nframes = n; ap.dst_off = d_dst_off + f0;
            ap.dst_cap = d_dst_cap + f0; ap.out_len = d_out_len + f0; ap.status = d_status + f0; ap.meta = d_meta + f0;
            ap.table_off = d_table_off + f0; ap.nrec = d_nrec + f0;
            // persistent warps that take frames from a ticket (no warp idles behind the slowest frame of its CTA, no
            // tail of half-empty SMs at the end of a launch) once there are more frames than resident warps
            const unsigned full = (n + kCodecWarps - 1) / kCodecWarps;
            const bool persistent = ctx->opt_persistent_decode && full > (unsigned)ctx->sm_count * 12u;
            unsigned long long *tk = d_tickets + 2 * part;
            if (persistent) CU(ctx, cudaMemsetAsync(tk, 0, 16, sp));
            pa.ticket = persistent ? tk : nullptr; ap.ticket = persistent ? tk + 1 : nullptr;
            { LaunchTimer lt(ctx, K_PARSE, sp);
              lz4_parse_kernel<<<persistent ? (unsigned)ctx->sm_count * 12u : full, kCodecThreads, 0, sp>>>(pa); }
            CU(ctx, cudaGetLastError());
            { LaunchTimer lt(ctx, K_DECODE, sp);
              lz4_decode_kernel<true><<<persistent ? (unsigned)ctx->sm_count * 8u : full, kCodecThreads, 0, sp>>>(ap); }
            CU(ctx, cudaGetLastError());
            if (two) {   // this half's un-shuffle follows on its own stream (the common tail below handles the one-stream case)
                FilterArgs fh;
                fh.src = d_stage; fh.dst = static_cast<uint8_t *>(d_dst);
                fh.ft.off = d_dst_off + f0; fh.ft.len = d_out_len + f0; fh.ft.uniform_len = 0; fh.ft.nframes = n;
                fh.ft.tiles_per_frame = tiles_for(max_orig, n, ctx);
                fh.meta = d_meta + f0; fh.uniform = FrameMeta{0, 0}; fh.status = d_status + f0; fh.inverse = 1;
                fh.copy_inactive = 0;
                { LaunchTimer lt(ctx, K_FILTER, sp);
                  filter_batch_kernel<<<(unsigned)((uint64_t)n * fh.ft.tiles_per_frame), kFilterThreads, 0, sp>>>(fh); }
                CU(ctx, cudaGetLastError());
            }
        }
        if (two) {
            for (int i = 0; i + 1 < parts; i++) {
                CU(ctx, cudaEventRecord(ctx->ev_join[i], ctx->s_side[i]));
                CU(ctx, cudaStreamWaitEvent(s, ctx->ev_join[i], 0));
            }
            return B2B_OK;
        }
    } else if (lanes) {
        // one lane per frame (lz4_decode3.cuh): no table, no parse kernel
        LaunchTimer lt(ctx, K_LANE, s);
        lz4_lane_decode_kernel<<<(nframes + kLaneThreads - 1) / kLaneThreads, kLaneThreads, 0, s>>>(a);
    } else {
        LaunchTimer lt(ctx, K_DECODE, s);
        lz4_decode_kernel<false><<<(nframes + kCodecWarps - 1) / kCodecWarps, kCodecThreads, 0, s>>>(a);
    }
    CU(ctx, cudaGetLastError());

    // unshuffle the frames that asked for it: stage -> dst (frames with mode 0 were decoded
    // straight into dst and are skipped because src != dst is only copied for mode != 0)
    FilterArgs fa;
    fa.src = d_stage; fa.dst = static_cast<uint8_t *>(d_dst);
    fa.ft.off = d_dst_off; fa.ft.len = d_out_len; fa.ft.uniform_len = 0; fa.ft.nframes = nframes;
    fa.ft.tiles_per_frame = tiles_for(max_orig, nframes, ctx);
    fa.meta = d_meta; fa.uniform = FrameMeta{0, 0}; fa.status = d_status; fa.inverse = 1;
    fa.copy_inactive = 0;  // mode 0 frames are already in dst
    { LaunchTimer lt(ctx, K_FILTER, s); filter_batch_kernel<<<(unsigned)((uint64_t)nframes * fa.ft.tiles_per_frame), kFilterThreads, 0, s>>>(fa); }
    CU(ctx, cudaGetLastError());
    return B2B_OK;
}

// ---- Blosc-1 multi-block frames (blocks.cuh): every block is a frame of its own to K1..K4 ----------
int compress_blocks_dev_locked(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                               const uint32_t *d_src_len, uint32_t nframes, uint64_t total_src,
                               uint32_t max_len, int shuffle, int64_t typesize, uint32_t blocksize,
                               void *d_dst, uint64_t dst_cap, uint64_t *d_frame_off, uint32_t *d_frame_len,
                               uint32_t *d_status, uint64_t *d_total_out, cudaStream_t s) {
    NvtxRange nvtx_range("b2b.compress_blocks_dev");
    ArenaGuard arena_guard(ctx, s);
    if (nframes == 0) {
        if (d_total_out) CU(ctx, cudaMemsetAsync(d_total_out, 0, 8, s));
        return B2B_OK;
    }
    if (!d_src || !d_src_off || !d_src_len || !d_dst || !d_frame_off || !d_frame_len || !d_status)
        return B2B_EINVAL;
    if (((uintptr_t)d_dst & 15u) != 0) return B2B_EINVAL;
    if (blocksize != 0 && blocksize < kB1MinBuffer) return B2B_EINVAL;
    if (dst_cap < total_src + 31ull * nframes) return B2B_EDST_TOO_SMALL;
    const uint32_t T = b1_typesize(typesize);
    const uint32_t b_full = b1_blocksize(~0ull, T, blocksize);          // what every frame of >= blocksize bytes uses
    const uint64_t nslots64 = total_src / b_full + 2ull * nframes;      // shorter frames: one block + a partial one
    if (nslots64 >= (1ull << 31)) return B2B_EINVAL;
    const uint32_t nslots = (uint32_t)nslots64;
    const uint32_t max_blk = std::min<uint32_t>(max_len, b_full);
    const FrameMeta fm = uniform_meta(shuffle, T);
    const bool filtered = fm.mode != 0;
    const uint32_t base_flags = (kB1Lz4Format << 5) | kB1DontSplit |
                                (shuffle == B2B_SHUFFLE ? B2B_FLAG_SHUFFLE : shuffle == B2B_BITSHUFFLE ? B2B_FLAG_BITSHUFFLE : 0);

    const uint64_t comp_bytes = comp_scratch_bytes(total_src, nslots);
    const uint64_t max_segs_total = total_src / kSegBytes + nslots + 1;
    const uint64_t need = (filtered ? align_up(total_src + 64, 256) : 0) + align_up(comp_bytes, 256) +
                          align_up(16 * max_segs_total, 256) + align_up(16 * max_segs_total, 256) +
                          12 * align_up(8ull * (nslots + 1), 256) + 4 * align_up(8ull * nframes, 256) +
                          3 * scan_scratch_bytes(nslots + 1) + 2 * scan_scratch_bytes(nframes) + 16384;
    int rc = ensure_arena(ctx, need);
    if (rc) return rc;
    Arena ar(ctx);
    uint8_t *d_shuf = filtered ? ar.take<uint8_t>(total_src + 64) : nullptr;
    uint8_t *d_comp = ar.take<uint8_t>(comp_bytes);
    SegMeta *d_meta = ar.take<SegMeta>(max_segs_total);
    SegPlace *d_place = ar.take<SegPlace>(max_segs_total);
    uint64_t *blk_off = ar.take<uint64_t>(nslots);
    uint64_t *comp_off = ar.take<uint64_t>(nslots);
    uint64_t *seg_base = ar.take<uint64_t>(nslots);
    uint64_t *blk_pos = ar.take<uint64_t>(nslots + 1);
    uint32_t *blk_len = ar.take<uint32_t>(nslots);
    uint32_t *owner = ar.take<uint32_t>(nslots);
    uint32_t *comp_len = ar.take<uint32_t>(nslots + 1);
    uint32_t *blk_flen = ar.take<uint32_t>(nslots);
    uint32_t *blk_flags = ar.take<uint32_t>(nslots);
    uint32_t *final_ll = ar.take<uint32_t>(nslots);
    uint32_t *final_off = ar.take<uint32_t>(nslots);
    uint32_t *blk_status = ar.take<uint32_t>(nslots);
    uint32_t *frm_bs = ar.take<uint32_t>(nframes);
    uint32_t *frm_nblk = ar.take<uint32_t>(nframes);
    uint64_t *blk_base = ar.take<uint64_t>(nframes);
    uint32_t *frm_flags = ar.take<uint32_t>(nframes);
    uint8_t *scan_a = ar.take<uint8_t>(scan_scratch_bytes(nslots + 1));
    uint8_t *scan_b = ar.take<uint8_t>(scan_scratch_bytes(nslots + 1));
    uint8_t *scan_c = ar.take<uint8_t>(scan_scratch_bytes(nslots + 1));
    uint8_t *scan_d = ar.take<uint8_t>(scan_scratch_bytes(nframes));
    uint8_t *scan_e = ar.take<uint8_t>(scan_scratch_bytes(nframes));
    unsigned long long *d_ticket = ar.take<unsigned long long>(4);

    BlocksGeomArgs ga;
    ga.src_len = d_src_len; ga.nframes = nframes; ga.typesize = T; ga.blocksize_req = blocksize;
    ga.bs = frm_bs; ga.nblk = frm_nblk;
    { LaunchTimer lt(ctx, K_BLOCKS_META, s); blocks_geom_kernel<<<(nframes + 127) / 128, 128, 0, s>>>(ga); }
    CU(ctx, cudaGetLastError());
    rc = launch_scan(ctx, frm_nblk, nframes, blk_base, nullptr, kScanIdentity, scan_d, s);
    if (rc) return rc;
    BlocksExpandArgs xa;
    xa.blk_base = blk_base; xa.nblk = frm_nblk; xa.bs = frm_bs; xa.src_off = d_src_off; xa.src_len = d_src_len;
    xa.nframes = nframes; xa.nslots = nslots; xa.owner = owner; xa.blk_off = blk_off; xa.blk_len = blk_len;
    { LaunchTimer lt(ctx, K_BLOCKS_META, s); blocks_expand_kernel<<<(nslots + 127) / 128, 128, 0, s>>>(xa); }
    CU(ctx, cudaGetLastError());

    const uint8_t *in = static_cast<const uint8_t *>(d_src);
    if (filtered) {
        rc = launch_filter(ctx, in, d_shuf, blk_off, blk_len, 0, nslots, max_blk, nullptr, fm, nullptr, 0, s, total_src);
        if (rc) return rc;
        in = d_shuf;
    }
    rc = launch_scan(ctx, blk_len, nslots, comp_off, nullptr, kScanSegSlot, scan_a, s);
    if (rc) return rc;
    rc = launch_scan(ctx, blk_len, nslots, seg_base, nullptr, kScanSegCount, scan_b, s);
    if (rc) return rc;
    uint64_t segs_grid = std::max<uint64_t>(1, ((uint64_t)max_blk + kSegBytes - 1) / kSegBytes);
    segs_grid = std::min<uint64_t>(segs_grid, std::max<uint64_t>(1, (1ull << 30) / nslots));
    EncodeArgs e;
    e.in = in; e.src_off = blk_off; e.src_len = blk_len; e.nframes = nslots;
    e.segs_grid = (uint32_t)segs_grid; e.comp = d_comp; e.comp_off = comp_off;
    e.seg_base = seg_base; e.meta = d_meta; e.ticket = d_ticket;
    for (int i = 0; i < 4; i++) e.tune[i] = ctx->opt_tune[i];
    e.independent = 0; e.planes = 0;
    e.phase_mask = (fm.mode == 2 && (fm.typesize & (fm.typesize - 1)) == 0 && fm.typesize <= 512) ? 8u * fm.typesize - 1u : 0u;
    e.comp_cap = comp_bytes; e.seg_cap = max_segs_total; e.src_cap = filtered ? total_src : ~0ull;
    rc = launch_encode(ctx, e, !filtered, s);
    if (rc) return rc;
    FinalizeArgs fa;
    fa.src_len = blk_len; fa.seg_base = seg_base; fa.meta = d_meta; fa.place = d_place;
    fa.nframes = nslots; fa.shuffle_flag = 0; fa.keep_raw = 0;
    fa.comp_len = comp_len; fa.frame_len = blk_flen; fa.flags = blk_flags;
    fa.final_ll = final_ll; fa.final_off = final_off; fa.status = blk_status;
    fa.index = nullptr; fa.segs_per_frame = 0;
    fa.comp_off = comp_off; fa.comp_cap = comp_bytes; fa.seg_cap = max_segs_total;
    fa.src_off = blk_off; fa.src_cap = filtered ? total_src : ~0ull;
    { LaunchTimer lt(ctx, K_FINALIZE, s); finalize_frames_kernel<<<(unsigned)(((uint64_t)nslots * 32 + 127) / 128), 128, 0, s>>>(fa); }
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemsetAsync(comp_len + nslots, 0, 4, s));
    rc = launch_scan(ctx, comp_len, nslots + 1, blk_pos, nullptr, kScanStream, scan_c, s);
    if (rc) return rc;
    BlocksFrameArgs ba;
    ba.src_len = d_src_len; ba.nblk = frm_nblk; ba.blk_base = blk_base; ba.blk_pos = blk_pos;
    ba.nframes = nframes; ba.base_flags = base_flags; ba.nslots = nslots; ba.blk_status = blk_status;
    ba.frame_len = d_frame_len; ba.frame_flags = frm_flags; ba.status = d_status;
    { LaunchTimer lt(ctx, K_BLOCKS_META, s); blocks_frame_kernel<<<(nframes + 127) / 128, 128, 0, s>>>(ba); }
    CU(ctx, cudaGetLastError());
    rc = launch_scan(ctx, d_frame_len, nframes, d_frame_off, d_total_out, kScanAlign16, scan_e, s);
    if (rc) return rc;

    BlocksPackArgs pk;
    pk.p.in = in; pk.p.raw = in; pk.p.src_off = blk_off; pk.p.src_len = blk_len; pk.p.comp = d_comp;
    pk.p.comp_off = comp_off; pk.p.seg_base = seg_base; pk.p.meta = d_meta; pk.p.place = d_place;
    pk.p.comp_len = comp_len; pk.p.flags = blk_flags; pk.p.final_ll = final_ll; pk.p.final_off = final_off;
    pk.p.status = blk_status; pk.p.frame_off = nullptr; pk.p.dst = nullptr;
    pk.p.nframes = nslots; pk.p.segs_grid = 1; pk.p.codec = B2B_LZ4; pk.p.typesize_u8 = T; pk.p.header = 0; pk.p.dst_cap = 0;
    pk.orig = static_cast<const uint8_t *>(d_src); pk.owner = owner; pk.blk_base = blk_base;
    pk.nblk = frm_nblk; pk.bs = frm_bs; pk.blk_pos = blk_pos; pk.src_off = d_src_off; pk.src_len = d_src_len;
    pk.frame_off = d_frame_off; pk.frame_len = d_frame_len; pk.frame_flags = frm_flags; pk.frame_status = d_status;
    pk.dst = static_cast<uint8_t *>(d_dst); pk.typesize = T;
    { LaunchTimer lt(ctx, K_BLOCKS_PACK, s); blocks_pack_kernel<<<nslots, kFilterThreads, 0, s>>>(pk); }
    CU(ctx, cudaGetLastError());
    return B2B_OK;
}

int decompress_blocks_dev_locked(b2b_ctx *ctx, const void *d_frames, const uint64_t *d_frame_off,
                                 const uint32_t *d_frame_len, uint32_t nframes, void *d_dst,
                                 const uint64_t *d_dst_off, const uint32_t *d_dst_cap, uint64_t total_dst,
                                 uint32_t max_orig, uint32_t blocksize, uint32_t *d_out_len,
                                 uint32_t *d_status, cudaStream_t s) {
    NvtxRange nvtx_range("b2b.decompress_blocks_dev");
    ArenaGuard arena_guard(ctx, s);
    if (nframes == 0) return B2B_OK;
    if (!d_frames || !d_frame_off || !d_frame_len || !d_dst || !d_dst_off || !d_dst_cap || !d_out_len || !d_status)
        return B2B_EINVAL;
    if (blocksize != 0 && blocksize < kB1MinBuffer) return B2B_EINVAL;
    const uint32_t slice = blocksize ? blocksize : kB1DefaultBlock;
    // a frame written with this block size has blocks of whole elements: at least slice - 254 bytes
    const uint32_t slice_min = slice - std::min<uint32_t>(254u, slice / 2);
    const uint64_t nslots64 = total_dst / slice_min + 2ull * nframes;
    if (nslots64 >= (1ull << 31)) return B2B_EINVAL;
    const uint32_t nslots = (uint32_t)nslots64;
    const uint64_t nrec_max = total_dst / 4 + (uint64_t)(kSeqSlack + 1) * nslots + 64;
    const uint64_t need = align_up(total_dst + 64, 256) + 9 * align_up(8ull * nslots, 256) +
                          4 * align_up(8ull * nframes, 256) + align_up(8 * nrec_max, 256) +
                          scan_scratch_bytes(nframes) + scan_scratch_bytes(nslots) + 8192;
    int rc = ensure_arena(ctx, need);
    if (rc) return rc;
    Arena ar(ctx);
    uint8_t *d_stage = ar.take<uint8_t>(total_dst + 64);
    uint64_t *d_table = ar.take<uint64_t>(nrec_max);
    uint64_t *slot_off = ar.take<uint64_t>(nslots);
    uint64_t *strm_src = ar.take<uint64_t>(nslots);
    uint64_t *table_off = ar.take<uint64_t>(nslots);
    uint32_t *slot_len = ar.take<uint32_t>(nslots);
    uint32_t *strm_clen = ar.take<uint32_t>(nslots);
    uint32_t *strm_kind = ar.take<uint32_t>(nslots);
    uint32_t *strm_nrec = ar.take<uint32_t>(nslots);
    uint32_t *owner = ar.take<uint32_t>(nslots);
    FrameMeta *slot_meta = ar.take<FrameMeta>(nslots);
    uint64_t *blk_base = ar.take<uint64_t>(nframes);
    uint32_t *frm_nblk = ar.take<uint32_t>(nframes);
    uint32_t *frm_bs = ar.take<uint32_t>(nframes);
    uint8_t *scan_a = ar.take<uint8_t>(scan_scratch_bytes(nframes));
    uint8_t *scan_b = ar.take<uint8_t>(scan_scratch_bytes(nslots));
    uint32_t *d_cap_eff = ar.take<uint32_t>(nframes);
    clip_caps_kernel<<<(nframes + 255) / 256, 256, 0, s>>>(d_dst_off, d_dst_cap, total_dst, nframes, d_cap_eff);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    d_dst_cap = d_cap_eff;

    BlocksInfoArgs ia;
    ia.frames = static_cast<const uint8_t *>(d_frames); ia.frame_off = d_frame_off; ia.frame_len = d_frame_len;
    ia.dst_cap = d_dst_cap; ia.nframes = nframes; ia.slice = slice; ia.slice_min = slice_min; ia.nblk = frm_nblk; ia.bs = frm_bs;
    ia.out_len = d_out_len; ia.status = d_status;
    { LaunchTimer lt(ctx, K_BLOCKS_META, s); blocks_info_kernel<<<(nframes + 127) / 128, 128, 0, s>>>(ia); }
    CU(ctx, cudaGetLastError());
    rc = launch_scan(ctx, frm_nblk, nframes, blk_base, nullptr, kScanIdentity, scan_a, s);
    if (rc) return rc;
    BlocksOwnerArgs oa;
    oa.blk_base = blk_base; oa.nblk = frm_nblk; oa.nframes = nframes; oa.nslots = nslots; oa.owner = owner;
    { LaunchTimer lt(ctx, K_BLOCKS_META, s); blocks_owner_kernel<<<(nslots + 127) / 128, 128, 0, s>>>(oa); }
    CU(ctx, cudaGetLastError());
    BlocksDecodeArgs da;
    da.frames = ia.frames; da.frame_off = d_frame_off; da.owner = owner; da.blk_base = blk_base; da.bs = frm_bs;
    da.dst = static_cast<uint8_t *>(d_dst); da.scratch = d_stage; da.dst_off = d_dst_off; da.status = d_status;
    da.nslots = nslots; da.strm_src = strm_src; da.strm_clen = strm_clen; da.strm_kind = strm_kind;
    da.slot_off = slot_off; da.slot_len = slot_len; da.slot_meta = slot_meta;
    { LaunchTimer lt(ctx, K_BLOCKS_META, s); blocks_streams_kernel<<<(nslots + 127) / 128, 128, 0, s>>>(da); }
    CU(ctx, cudaGetLastError());
    // K4's two halves over the block streams: parse kernel -> sequence records -> copy kernel
    rc = launch_scan(ctx, slot_len, nslots, table_off, nullptr, kScanSeqSlots, scan_b, s);
    if (rc) return rc;
    StreamArgs sa;
    sa.src = ia.frames; sa.src_off = strm_src; sa.clen = strm_clen; sa.dst = da.dst; sa.scratch = d_stage;
    sa.dst_off = slot_off; sa.cap = slot_len; sa.kind = strm_kind; sa.owner = owner; sa.status = d_status;
    sa.nstreams = nslots; sa.table = d_table; sa.table_off = table_off; sa.nrec = strm_nrec; sa.table_cap = nrec_max;
    const unsigned sgrid = (nslots + kCodecWarps - 1) / kCodecWarps;
    { LaunchTimer lt(ctx, K_PARSE, s); lz4_parse_streams_kernel<<<sgrid, kCodecThreads, 0, s>>>(sa); }
    CU(ctx, cudaGetLastError());
    { LaunchTimer lt(ctx, K_DECODE, s); lz4_copy_streams_kernel<<<sgrid, kCodecThreads, 0, s>>>(sa); }
    CU(ctx, cudaGetLastError());
    { LaunchTimer lt(ctx, K_BLOCKS_DECODE, s); blocks_decode_kernel<<<sgrid, kCodecThreads, 0, s>>>(da); }   // split blocks
    CU(ctx, cudaGetLastError());
    { LaunchTimer lt(ctx, K_BLOCKS_META, s);
      blocks_finish_kernel<<<(nframes + 127) / 128, 128, 0, s>>>(d_status, d_out_len, blk_base, frm_nblk, nslots, nframes); }
    CU(ctx, cudaGetLastError());

    FilterArgs fa;
    fa.src = d_stage; fa.dst = static_cast<uint8_t *>(d_dst);
    fa.ft.off = slot_off; fa.ft.len = slot_len; fa.ft.uniform_len = 0; fa.ft.nframes = nslots;
    fa.ft.tiles_per_frame = tiles_for(std::min<uint32_t>(max_orig, slice), nslots, ctx);
    fa.meta = slot_meta; fa.uniform = FrameMeta{0, 0}; fa.status = nullptr; fa.inverse = 1;
    fa.copy_inactive = 0;
    { LaunchTimer lt(ctx, K_FILTER, s);
      filter_batch_kernel<<<(unsigned)((uint64_t)nslots * fa.ft.tiles_per_frame), kFilterThreads, 0, s>>>(fa); }
    CU(ctx, cudaGetLastError());
    return B2B_OK;
}

int shuffle_dev_locked(b2b_ctx *ctx, int mode, int inverse, int64_t typesize, const void *d_src,
                       void *d_dst, size_t n, cudaStream_t s) {
    if (n == 0) return B2B_OK;
    if (!d_src || !d_dst) return B2B_EINVAL;
    if (typesize > 0xFFFFFFFFll && (uint64_t)n >= (uint64_t)typesize) return B2B_EUNSUPPORTED;
    const FrameMeta fm = uniform_meta(mode, typesize);
    const uint8_t *src = static_cast<const uint8_t *>(d_src);
    uint8_t *dst = static_cast<uint8_t *>(d_dst);
    const bool identity = fm.mode == 0 || (uint64_t)n < fm.typesize;
    if (identity) {
        if (src != dst) CU(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s));
        return B2B_OK;
    }
    NvtxRange nvtx_range("b2b.shuffle_dev");
    ArenaGuard arena_guard(ctx, s);
    uint8_t *out = dst;
    if (src == dst) {  // in place (ShuffleBuffer semantics): transform into scratch, copy back
        int rc = ensure_arena(ctx, n + 4096);
        if (rc) return rc;
        out = ctx->arena;
    }
    int rc = launch_filter(ctx, src, out, nullptr, nullptr, n, 1, n, nullptr, fm, nullptr, inverse ? 1 : 0, s);
    if (rc) return rc;
    if (out != dst) CU(ctx, cudaMemcpyAsync(dst, out, n, cudaMemcpyDeviceToDevice, s));
    return B2B_OK;
}

}  // namespace

// =========================================================================================
// extern "C"
// =========================================================================================
extern "C" {

const char *b2b_version(void) { return "b2b 0.1.0 (sm_100a; go-blosc v1.0.2 frame format)"; }

const char *b2b_strerror(int st) {
    switch (st) {
        case B2B_OK: return "ok";
        case B2B_EINVALID_DATA: return "blosc: invalid compressed data";
        case B2B_EINVALID_HEADER: return "blosc: invalid header";
        case B2B_EINVALID_VERSION: return "blosc: unsupported format version";
        case B2B_EINVALID_CODEC: return "blosc: unsupported codec";
        case B2B_ESIZE_MISMATCH: return "blosc: decompressed size mismatch";
        case B2B_EDATA_TOO_LARGE: return "blosc: data too large";
        case B2B_ECOMPRESSION_FAILED: return "blosc: compression failed";
        case B2B_EDECOMPRESSION_FAILED: return "blosc: decompression failed";
        case B2B_ECUDA: return "b2b: CUDA error";
        case B2B_EUNSUPPORTED: return "b2b: codec outside the GPU path";
        case B2B_EDST_TOO_SMALL: return "b2b: destination too small";
        case B2B_EINVAL: return "b2b: invalid argument";
        default: return "b2b: unknown status";
    }
}

int b2b_init(int device, b2b_ctx **out) {
    if (!out) return B2B_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
        return B2B_ECUDA;  // no CPU fallback: fail loudly
    b2b_ctx *ctx = new (std::nothrow) b2b_ctx();
    if (!ctx) return B2B_ECUDA;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return B2B_ECUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return B2B_ECUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_tab, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_arena, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < b2b_ctx::kSide && ok; i++)
        ok = cudaStreamCreateWithFlags(&ctx->s_side[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < b2b_ctx::kSlots && ok; i++)
        ok = cudaStreamCreateWithFlags(&ctx->s_k[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_in_ready[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_in_free[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_out_free[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_tab[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { b2b_destroy(ctx); return B2B_ECUDA; }
    *out = ctx;
    return B2B_OK;
}

void b2b_destroy(b2b_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    fold_timings(ctx);
    delete ctx->staging;               // joins its threads before the streams it uses go away
    ctx->staging = nullptr;
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    select_arena(ctx, 0);
    for (int i = 0; i < 5; i++) if (ctx->arenas[i]) cudaFree(ctx->arenas[i]);
    for (int i = 0; i < 4; i++) if (ctx->s_k[i]) cudaStreamDestroy(ctx->s_k[i]);
    for (int i = 0; i < 3 * b2b_ctx::kSlots; i++) if (ctx->hbuf[i]) cudaFree(ctx->hbuf[i]);
    for (int i = 0; i < b2b_ctx::kSlots; i++) if (ctx->ptab[i]) cudaFreeHost(ctx->ptab[i]);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->s_tab) cudaStreamDestroy(ctx->s_tab);
    for (int i = 0; i < b2b_ctx::kSlots; i++) {
        if (ctx->ev_tab[i]) cudaEventDestroy(ctx->ev_tab[i]);
        if (ctx->ev_in_ready[i]) cudaEventDestroy(ctx->ev_in_ready[i]);
        if (ctx->ev_in_free[i]) cudaEventDestroy(ctx->ev_in_free[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
        if (ctx->ev_out_free[i]) cudaEventDestroy(ctx->ev_out_free[i]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->ev_arena) cudaEventDestroy(ctx->ev_arena);
    for (int i = 0; i < b2b_ctx::kSide; i++) {
        if (ctx->s_side[i]) cudaStreamDestroy(ctx->s_side[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    delete ctx;
}

const char *b2b_last_error(b2b_ctx *ctx) { return ctx ? ctx->last_err.c_str() : ""; }

int b2b_set_option(b2b_ctx *ctx, int option, int64_t value) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    switch (option) {
        case B2B_OPT_REF_MEMCPY_QUIRK: ctx->opt_quirk = value != 0; return B2B_OK;
        case B2B_OPT_FILTER_CTAS_PER_SM: ctx->opt_filter_ctas_per_sm = (int)std::max<int64_t>(0, value); return B2B_OK;
        case B2B_OPT_HASH_LOG:
            if (value != 0 && (value < 10 || value > 13)) return B2B_EINVAL;
            ctx->opt_hash_log = (int)value; return B2B_OK;
        case B2B_OPT_KERNEL_TIMING: ctx->opt_timing = value != 0; return B2B_OK;
        case B2B_OPT_HASH_BYTES:
            if (value != 0 && (value < 4 || value > 6)) return B2B_EINVAL;
            ctx->opt_hash_bytes = (int)value; return B2B_OK;
        case 100: case 101: case 102: case 103: ctx->opt_tune[option - 100] = (uint32_t)value; return B2B_OK;
        case 104: if (value < -1 || value > 4) return B2B_EINVAL; ctx->opt_fused_decode = (int)value; return B2B_OK;
        case B2B_OPT_HOST_STAGE_BYTES: ctx->opt_stage_bytes = value > 0 ? (uint64_t)value : (128ull << 20); return B2B_OK;
        case B2B_OPT_HOST_THREADS:
            if (value < 0 || value > 64) return B2B_EINVAL;
            ctx->opt_host_threads = (int)value;
            if (ctx->staging) { cudaDeviceSynchronize(); delete ctx->staging; ctx->staging = nullptr; }
            return B2B_OK;
        case B2B_OPT_NO_HOST_STAGING: ctx->opt_no_staging = value != 0; return B2B_OK;
        case B2B_OPT_FUSE_UNSHUFFLE: ctx->opt_fuse_unshuffle = value != 0; return B2B_OK;
        case 105: ctx->opt_persistent_decode = value != 0; return B2B_OK;
        case 109: ctx->opt_parse_ctas = (int)std::max<int64_t>(0, value); return B2B_OK;
        case 108: ctx->opt_no_small_chunks = value != 0; return B2B_OK;
        case 107: ctx->opt_jump_min_bytes = (uint32_t)std::max<int64_t>(0, value); return B2B_OK;
        case 106: ctx->opt_encode_ctas = (int)std::max<int64_t>(0, value); return B2B_OK;
        case B2B_OPT_DECODE_STREAMS:
            if (value < 0 || value > 1 + b2b_ctx::kSide) return B2B_EINVAL;
            ctx->opt_decode_streams = (int)value; return B2B_OK;
        case B2B_OPT_DECODER:
            if (value < -1 || value > 4) return B2B_EINVAL;
            ctx->opt_fused_decode = (int)value; return B2B_OK;
        default: return B2B_EINVAL;
    }
}

int b2b_reserve(b2b_ctx *ctx, uint64_t total, uint32_t nframes) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t need = align_up(total + 64, 256) + comp_scratch_bytes(total, nframes) +
                          24 * (total / kSegBytes + nframes + 1) + 8 * align_up(8ull * nframes, 256) +
                          3 * scan_scratch_bytes(nframes) + (1 << 20);
    return ensure_arena(ctx, need);
}

uint64_t b2b_launch_count(b2b_ctx *ctx) { return ctx ? ctx->launches : 0; }

int b2b_kernel_stats(b2b_ctx *ctx, int kernel, const char **name, uint64_t *launches, double *total_ms) {
    if (!ctx || kernel < 0 || kernel >= K_COUNT) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    fold_timings(ctx);
    if (name) *name = kKernelNames[kernel];
    if (launches) *launches = ctx->kernel_launches[kernel];
    if (total_ms) *total_ms = ctx->kernel_ms[kernel];
    return B2B_OK;
}

int b2b_kernel_stats_reset(b2b_ctx *ctx) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->device);
    fold_timings(ctx);
    for (int k = 0; k < K_COUNT; k++) { ctx->kernel_ms[k] = 0; ctx->kernel_launches[k] = 0; }
    return B2B_OK;
}

size_t b2b_max_frame_size(size_t n) { return n + B2B_HEADER_SIZE; }
size_t b2b_lz4_bound(size_t n) { return n + n / 255 + 16; }

int b2b_parse_header(const void *frame, size_t len, b2b_header *h) {
    if (!h || (!frame && len)) return B2B_EINVAL;
    if (len < B2B_HEADER_SIZE) return B2B_EINVALID_HEADER;             // blosc.go:166-168
    const uint8_t *p = static_cast<const uint8_t *>(frame);
    auto rd = [&](int o) { return (uint32_t)p[o] | ((uint32_t)p[o + 1] << 8) | ((uint32_t)p[o + 2] << 16) | ((uint32_t)p[o + 3] << 24); };
    h->version = p[0]; h->versionlz = p[1]; h->flags = p[2]; h->typesize = p[3];
    h->nbytes_orig = rd(4); h->blocksize = rd(8); h->nbytes_comp = rd(12);
    if (h->version != B2B_FORMAT_VERSION) return B2B_EINVALID_VERSION; // blosc.go:180-182
    return B2B_OK;
}

void b2b_header_bytes(const b2b_header *h, uint8_t out[16]) {
    out[0] = h->version; out[1] = h->versionlz; out[2] = h->flags; out[3] = h->typesize;
    const uint32_t v[3] = {h->nbytes_orig, h->blocksize, h->nbytes_comp};
    for (int k = 0; k < 3; k++)
        for (int b = 0; b < 4; b++) out[4 + 4 * k + b] = (uint8_t)(v[k] >> (8 * b));
}

// ---- device-pointer entry points --------------------------------------------------------
int b2b_shuffle_dev(b2b_ctx *ctx, int mode, int inverse, int64_t typesize, const void *d_src,
                    void *d_dst, size_t n, void *stream) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return shuffle_dev_locked(ctx, mode, inverse, typesize, d_src, d_dst, n, (cudaStream_t)stream);
}

int b2b_compress_batch_dev(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                           const uint32_t *d_src_len, uint32_t nframes, uint64_t total_src_bytes,
                           uint32_t max_frame_len, int shuffle, int64_t typesize, void *d_dst,
                           uint64_t dst_cap, uint64_t *d_frame_off, uint32_t *d_frame_len,
                           uint32_t *d_status, uint64_t *d_total_out, void *stream) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return compress_batch_dev_locked(ctx, d_src, d_src_off, d_src_len, nframes, total_src_bytes,
                                     max_frame_len, shuffle, typesize, d_dst, dst_cap, d_frame_off,
                                     d_frame_len, d_status, d_total_out, (cudaStream_t)stream);
}

int b2b_compress_batch_dev_indexed(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                                   const uint32_t *d_src_len, uint32_t nframes, uint64_t total_src_bytes,
                                   uint32_t max_frame_len, int shuffle, int64_t typesize, void *d_dst,
                                   uint64_t dst_cap, uint64_t *d_frame_off, uint32_t *d_frame_len,
                                   uint32_t *d_status, uint64_t *d_total_out, uint64_t *d_index,
                                   uint32_t segs_per_frame, void *stream) {
    if (!ctx || !d_index) return B2B_EINVAL;
    if (segs_per_frame < b2b_index_segments(max_frame_len)) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return compress_batch_dev_locked(ctx, d_src, d_src_off, d_src_len, nframes, total_src_bytes,
                                     max_frame_len, shuffle, typesize, d_dst, dst_cap, d_frame_off,
                                     d_frame_len, d_status, d_total_out, (cudaStream_t)stream, false,
                                     d_index, segs_per_frame);
}

int b2b_decompress_batch_dev_indexed(b2b_ctx *ctx, const void *d_frames, const uint64_t *d_frame_off,
                                     const uint32_t *d_frame_len, uint32_t nframes, int64_t typesize_override,
                                     void *d_dst, const uint64_t *d_dst_off, const uint32_t *d_dst_cap,
                                     uint64_t total_dst_bytes, uint32_t max_orig_len, uint32_t *d_out_len,
                                     uint32_t *d_status, const uint64_t *d_index, uint32_t segs_per_frame,
                                     void *stream) {
    if (!ctx || !d_index || segs_per_frame == 0) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return decompress_batch_dev_locked(ctx, d_frames, d_frame_off, d_frame_len, nframes,
                                       typesize_override, d_dst, d_dst_off, d_dst_cap,
                                       total_dst_bytes, max_orig_len, d_out_len, d_status,
                                       (cudaStream_t)stream, d_index, segs_per_frame);
}

uint32_t b2b_index_segments(uint32_t max_frame_len) { return max_frame_len ? (max_frame_len + kSegBytes - 1) / kSegBytes : 1; }

int b2b_frame_info_batch_dev(b2b_ctx *ctx, const void *d_frames, const uint64_t *d_frame_off,
                             const uint32_t *d_frame_len, uint32_t nframes, uint32_t *d_orig_len,
                             uint64_t *d_dst_off, uint64_t *d_total, uint32_t *d_status, void *stream) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (nframes == 0) { if (d_total) CU(ctx, cudaMemsetAsync(d_total, 0, 8, s)); return B2B_OK; }
    if (!d_frames || !d_frame_off || !d_frame_len || !d_orig_len || !d_dst_off || !d_status) return B2B_EINVAL;
    ArenaGuard arena_guard(ctx, s);
    int rc = ensure_arena(ctx, scan_scratch_bytes(nframes) + 4096);
    if (rc) return rc;
    { LaunchTimer lt(ctx, K_INFO, s);
      frame_info_kernel<<<(nframes + 255) / 256, 256, 0, s>>>(static_cast<const uint8_t *>(d_frames), d_frame_off,
                                                              d_frame_len, nframes, d_orig_len, d_status); }
    CU(ctx, cudaGetLastError());
    return launch_scan(ctx, d_orig_len, nframes, d_dst_off, d_total, kScanAlign16, ctx->arena, s);
}

int b2b_decompress_batch_dev(b2b_ctx *ctx, const void *d_frames, const uint64_t *d_frame_off,
                             const uint32_t *d_frame_len, uint32_t nframes, int64_t typesize_override,
                             void *d_dst, const uint64_t *d_dst_off, const uint32_t *d_dst_cap,
                             uint64_t total_dst_bytes, uint32_t max_orig_len, uint32_t *d_out_len,
                             uint32_t *d_status, void *stream) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return decompress_batch_dev_locked(ctx, d_frames, d_frame_off, d_frame_len, nframes,
                                       typesize_override, d_dst, d_dst_off, d_dst_cap,
                                       total_dst_bytes, max_orig_len, d_out_len, d_status,
                                       (cudaStream_t)stream);
}

// ---- Blosc-1 multi-block frames --------------------------------------------------------------------
uint32_t b2b_blocks_blocksize(size_t n, int64_t typesize, uint32_t blocksize) {
    return b1_blocksize(n, b1_typesize(typesize), blocksize);
}

int b2b_compress_blocks_batch_dev(b2b_ctx *ctx, const void *d_src, const uint64_t *d_src_off,
                                  const uint32_t *d_src_len, uint32_t nframes, uint64_t total_src_bytes,
                                  uint32_t max_frame_len, int shuffle, int64_t typesize, uint32_t blocksize,
                                  void *d_dst, uint64_t dst_cap, uint64_t *d_frame_off, uint32_t *d_frame_len,
                                  uint32_t *d_status, uint64_t *d_total_out, void *stream) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return compress_blocks_dev_locked(ctx, d_src, d_src_off, d_src_len, nframes, total_src_bytes, max_frame_len,
                                      shuffle, typesize, blocksize, d_dst, dst_cap, d_frame_off, d_frame_len,
                                      d_status, d_total_out, (cudaStream_t)stream);
}

int b2b_decompress_blocks_batch_dev(b2b_ctx *ctx, const void *d_frames, const uint64_t *d_frame_off,
                                    const uint32_t *d_frame_len, uint32_t nframes, void *d_dst,
                                    const uint64_t *d_dst_off, const uint32_t *d_dst_cap,
                                    uint64_t total_dst_bytes, uint32_t max_orig_len, uint32_t blocksize,
                                    uint32_t *d_out_len, uint32_t *d_status, void *stream) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return decompress_blocks_dev_locked(ctx, d_frames, d_frame_off, d_frame_len, nframes, d_dst, d_dst_off,
                                        d_dst_cap, total_dst_bytes, max_orig_len, blocksize, d_out_len,
                                        d_status, (cudaStream_t)stream);
}

int b2b_compress_blocks(b2b_ctx *ctx, const void *src, size_t n, int shuffle, int64_t typesize,
                        uint32_t blocksize, void *dst, size_t cap, size_t *out_len) {
    if (!ctx || !out_len || !dst) return B2B_EINVAL;
    if (n == 0) return B2B_EINVALID_DATA;
    if (!src) return B2B_EINVAL;
    if (n > kB1MaxBuffer) return B2B_EDATA_TOO_LARGE;
    if (shuffle != B2B_NOSHUFFLE && shuffle != B2B_SHUFFLE && shuffle != B2B_BITSHUFFLE) return B2B_EINVAL;
    if (cap < n + 16) return B2B_EDST_TOO_SMALL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_tab = nullptr;
    int rc = ensure_hbuf(ctx, 0, n + 64, &d_in);
    if (rc == B2B_OK) rc = ensure_hbuf(ctx, 1, n + 256, &d_out);
    if (rc == B2B_OK) rc = ensure_hbuf(ctx, 2, 4096, &d_tab);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    uint64_t *d_off = reinterpret_cast<uint64_t *>(d_tab);            // [0] src_off, [1] frame_off, [2] total
    uint32_t *d_u32 = reinterpret_cast<uint32_t *>(d_tab + 256);      // [0] src_len, [1] frame_len, [2] status
    const uint64_t h_off[3] = {0, 0, 0};
    const uint32_t h_u32[3] = {(uint32_t)n, 0, 0};
    CU(ctx, cudaMemcpyAsync(d_in, src, n, cudaMemcpyHostToDevice, s));
    CU(ctx, cudaMemcpyAsync(d_off, h_off, sizeof h_off, cudaMemcpyHostToDevice, s));
    CU(ctx, cudaMemcpyAsync(d_u32, h_u32, sizeof h_u32, cudaMemcpyHostToDevice, s));
    rc = compress_blocks_dev_locked(ctx, d_in, d_off, d_u32, 1, n, (uint32_t)n, shuffle, typesize, blocksize, d_out,
                                    n + 256, d_off + 1, d_u32 + 1, d_u32 + 2, d_off + 2, s);
    if (rc) return rc;
    uint32_t h_res[3] = {0, 0, 0};
    CU(ctx, cudaMemcpyAsync(h_res, d_u32, sizeof h_res, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    if (h_res[2]) return (int)h_res[2];
    CU(ctx, cudaMemcpyAsync(dst, d_out, h_res[1], cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    *out_len = h_res[1];
    return B2B_OK;
}

int b2b_decompress_blocks(b2b_ctx *ctx, const void *frame, size_t len, void *dst, size_t cap, size_t *out_len) {
    if (!ctx || !out_len || (!frame && len)) return B2B_EINVAL;
    b2b_header h;
    int rc = b2b_parse_header(frame, len, &h);
    if (rc) return rc;
    if (len > 0xFFFFFFFFull) len = 0xFFFFFFFFull;                     // cbytes is a u32: the rest is not the frame
    if (h.nbytes_orig == 0 && h.nbytes_comp >= 16 && h.nbytes_comp <= len) { *out_len = 0; return B2B_OK; }
    if (!dst) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = std::min<uint64_t>(h.nbytes_orig, cap);
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_tab = nullptr;
    rc = ensure_hbuf(ctx, 0, len + 64, &d_in);
    if (rc == B2B_OK) rc = ensure_hbuf(ctx, 1, n + 256, &d_out);
    if (rc == B2B_OK) rc = ensure_hbuf(ctx, 2, 4096, &d_tab);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    uint64_t *d_off = reinterpret_cast<uint64_t *>(d_tab);            // [0] frame_off, [1] dst_off
    uint32_t *d_u32 = reinterpret_cast<uint32_t *>(d_tab + 256);      // [0] frame_len, [1] dst_cap, [2] out_len, [3] status
    const uint64_t h_off[2] = {0, 0};
    const uint32_t h_u32[4] = {(uint32_t)len, (uint32_t)n, 0, 0};
    CU(ctx, cudaMemcpyAsync(d_in, frame, len, cudaMemcpyHostToDevice, s));
    CU(ctx, cudaMemcpyAsync(d_off, h_off, sizeof h_off, cudaMemcpyHostToDevice, s));
    CU(ctx, cudaMemcpyAsync(d_u32, h_u32, sizeof h_u32, cudaMemcpyHostToDevice, s));
    // the table is sized from the header's own block size (a stored frame is copied in 64 KiB slices)
    uint32_t bs = (h.flags & B2B_FLAG_MEMCPY) ? 0u : std::min<uint32_t>(h.blocksize, h.nbytes_orig);
    if (bs != 0 && bs < kB1MinBuffer) bs = kB1MinBuffer;
    rc = decompress_blocks_dev_locked(ctx, d_in, d_off, d_u32, 1, d_out, d_off + 1, d_u32 + 1, n, (uint32_t)n, bs,
                                      d_u32 + 2, d_u32 + 3, s);
    if (rc) return rc;
    uint32_t h_res[4] = {0, 0, 0, 0};
    CU(ctx, cudaMemcpyAsync(h_res, d_u32, sizeof h_res, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    if (h_res[3]) return (int)h_res[3];
    CU(ctx, cudaMemcpyAsync(dst, d_out, h_res[2], cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    *out_len = h_res[2];
    return B2B_OK;
}

int b2b_scan_offsets_dev(b2b_ctx *ctx, const uint32_t *d_len, uint32_t n, uint64_t *d_off,
                         uint64_t *d_total, void *stream) {
    if (!ctx) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    if (n && (!d_len || !d_off)) return B2B_EINVAL;
    ArenaGuard arena_guard(ctx, (cudaStream_t)stream);
    int rc = ensure_arena(ctx, scan_scratch_bytes(n) + 4096);
    if (rc) return rc;
    return launch_scan(ctx, d_len, n, d_off, d_total, kScanIdentity, ctx->arena, (cudaStream_t)stream);
}

// ---- multi-GPU: sizes all-gather + global offsets (the only exchange of the sharded path) ---------------
int b2b_allgather_sizes(b2b_ctx *ctx, void *nccl_comm, const uint32_t *d_local_len, uint32_t n_local, uint32_t world,
                        uint32_t *d_all_len, uint64_t *d_all_off, uint64_t *d_total, int align16, void *stream) {
    if (!ctx || !nccl_comm || world == 0) return B2B_EINVAL;
    if ((uint64_t)n_local * world >= (1ull << 32)) return B2B_EINVAL;
    if (n_local && (!d_local_len || !d_all_len || !d_all_off)) return B2B_EINVAL;
    // ncclResult_t ncclAllGather(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t); ncclUint32 = 3
    typedef int (*allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
    static allgather_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy the process already uses
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) fn = reinterpret_cast<allgather_fn>(dlsym(h, "ncclAllGather"));
        if (!fn) fn = reinterpret_cast<allgather_fn>(dlsym(RTLD_DEFAULT, "ncclAllGather"));
    });
    if (!fn) return B2B_EUNSUPPORTED;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    NvtxRange nvtx_range("b2b.allgather_sizes");
    ArenaGuard arena_guard(ctx, s);
    const uint32_t n = n_local * world;
    int rc = ensure_arena(ctx, scan_scratch_bytes(n) + 4096);
    if (rc) return rc;
    if (n_local) {
        const int nrc = fn(d_local_len, d_all_len, n_local, 3, nccl_comm, s);
        if (nrc != 0) { ctx->last_err = "ncclAllGather failed with ncclResult_t " + std::to_string(nrc); return B2B_ECUDA; }
    }
    return launch_scan(ctx, d_all_len, n, d_all_off, d_total, align16 ? kScanAlign16 : kScanIdentity, ctx->arena, s);
}

// ---- host-pointer entry points ----------------------------------------------------------
int b2b_shuffle(b2b_ctx *ctx, int mode, int inverse, int64_t typesize, const void *src, void *dst,
                size_t n) {
    if (!ctx) return B2B_EINVAL;
    if (n == 0) return B2B_OK;
    if (!src || !dst) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    // stage buffers live outside the arena (shuffle_dev may use the arena for in-place)
    uint8_t *d_in = nullptr, *d_out = nullptr;
    int rc = ensure_hbuf(ctx, 0, n + 64, &d_in);
    if (rc) return rc;
    rc = ensure_hbuf(ctx, 1, n + 64, &d_out);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(d_in, src, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = shuffle_dev_locked(ctx, mode, inverse, typesize, d_in, d_out, n, ctx->stream);
    if (e == cudaSuccess && rc == B2B_OK) e = cudaMemcpyAsync(dst, d_out, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->last_err = cudaGetErrorString(e); return B2B_ECUDA; }
    return rc;
}

// Host batches run as a kSlots-deep pipeline over chunks of frames (B2B_OPT_HOST_STAGE_BYTES each):
//   s_in    H2D of the chunk's bytes and of its tables (from the slot's pinned table block)
//   s_k[i]  the kernels of the chunk in slot i, with the slot's own scratch arena: a chunk is far
//           too small to fill the device (one warp works a 64 KiB segment / one frame serially for
//           about a millisecond), so the chunks in flight have to overlap on the device as well
//   s_tab   D2H of the chunk's result tables into the pinned block
//   s_out   D2H of the chunk's bytes
// The host only ever waits for a chunk's TABLES (one chunk behind the one it just launched), never
// for a bulk copy, so H2D of chunk k+1, the kernels of chunk k and D2H of chunk k-1 overlap and
// both PCIe directions stay busy.  Pinned caller buffers are DMA'd directly; pageable ones (what a Go
// caller has) go through the context's pinned ring (host_staging.hpp), so they overlap the same way.
constexpr uint64_t kStagingMinBytes = 16ull << 20;   // pageable batches below this go to cudaMemcpyAsync directly
struct HostChunk { uint32_t f0, f1; uint64_t lo, hi; uint32_t max_len; uint64_t sum; };   // sum: bytes of all frames (> hi - lo when they overlap)

// Chunk sizes ramp up at the start of a batch and down at its end (stage / 4, stage / 2, stage, ..., stage / 2,
// stage / 4): nothing overlaps the first chunk's H2D or the last chunk's kernels and D2H, so those are kept short.
static uint64_t chunk_target(uint64_t done, uint64_t total, uint64_t stage) {
    const uint64_t left = total - done;
    uint64_t t = stage;
    if (done == 0) t = stage / 4;
    else if (done < stage) t = stage / 2;
    if (left <= stage / 4 + stage / 8) return left;
    if (left <= stage) t = std::min(t, left > stage / 2 ? left - stage / 4 : left);
    else if (left <= 2 * stage) t = std::min(t, stage / 2 + (left - stage) / 2);
    return std::max<uint64_t>(t, 1);
}

static std::vector<HostChunk> split_chunks(const uint64_t *off, const uint32_t *len, uint32_t nframes,
                                           uint64_t stage_bytes) {
    std::vector<HostChunk> out;
    uint64_t total = 0, done = 0;
    for (uint32_t i = 0; i < nframes; i++) total += len[i];
    uint32_t f = 0;
    while (f < nframes) {
        HostChunk c{f, f, ~0ull, 0, 0, 0};
        uint64_t bytes = 0;
        const uint64_t target = chunk_target(done, total, stage_bytes);
        while (c.f1 < nframes && (c.f1 == c.f0 || bytes + len[c.f1] <= target)) {
            c.lo = std::min(c.lo, off[c.f1]); c.hi = std::max(c.hi, off[c.f1] + len[c.f1]);
            c.max_len = std::max(c.max_len, len[c.f1]);
            bytes += len[c.f1]; c.f1++;
        }
        c.sum = bytes;
        done += bytes;
        out.push_back(c);
        f = c.f1;
    }
    return out;
}

// The filter writes frame f's transformed bytes at f's own source offset in scratch, so source ranges that
// overlap would clobber each other: such a batch is refused (frames in offset order are checked in one pass).
static bool ranges_disjoint(const uint64_t *off, const uint32_t *len, uint32_t n) {
    bool ordered = true;
    for (uint32_t f = 1; f < n && ordered; f++) ordered = off[f] >= off[f - 1] + len[f - 1];
    if (ordered) return true;
    std::vector<std::pair<uint64_t, uint32_t>> v(n);
    for (uint32_t f = 0; f < n; f++) v[f] = {off[f], len[f]};
    std::sort(v.begin(), v.end());
    for (uint32_t f = 1; f < n; f++)
        if (v[f].second != 0 && v[f].first < v[f - 1].first + v[f - 1].second) return false;
    return true;
}

static int sync_pipeline(b2b_ctx *ctx, cudaError_t e, int rc) {
    if (ctx->staging) { const cudaError_t es = ctx->staging->flush(); if (e == cudaSuccess) e = es; }
    const cudaError_t e1 = cudaStreamSynchronize(ctx->s_out), e2 = cudaStreamSynchronize(ctx->s_tab);
    cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < b2b_ctx::kSlots; i++) { const cudaError_t ek = cudaStreamSynchronize(ctx->s_k[i]); if (e3 == cudaSuccess) e3 = ek; }
    const cudaError_t e4 = cudaStreamSynchronize(ctx->s_in);
    if (e == cudaSuccess) e = e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3 != cudaSuccess ? e3 : e4;
    if (e != cudaSuccess) { ctx->last_err = cudaGetErrorString(e); return B2B_ECUDA; }
    return rc;
}

// blocks: Blosc-1 multi-block frames (blocks.cuh) of `blocksize` instead of the reference's one-block frames
static int host_compress_batch(b2b_ctx *ctx, const void *src, const uint64_t *src_off, const uint32_t *src_len,
                               uint32_t nframes, int shuffle, int64_t typesize, void *dst, uint64_t dst_cap,
                               uint64_t *frame_off, uint32_t *frame_len, uint32_t *status, uint64_t *total_out,
                               bool blocks, uint32_t blocksize) {
    NvtxRange nvtx_range("b2b.compress_batch (host pipeline)");
    if (!ctx) return B2B_EINVAL;
    if (nframes == 0) { if (total_out) *total_out = 0; return B2B_OK; }
    if (!src || !src_off || !src_len || !dst || !frame_off || !frame_len || !status) return B2B_EINVAL;
    if (!ranges_disjoint(src_off, src_len, nframes)) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    constexpr int S = b2b_ctx::kSlots;
    const uint8_t *hsrc = static_cast<const uint8_t *>(src);
    uint8_t *hdst = static_cast<uint8_t *>(dst);
    const std::vector<HostChunk> chunks = split_chunks(src_off, src_len, nframes, ctx->opt_stage_bytes);
    uint64_t running = 0;
    int rc = B2B_OK;
    cudaError_t e = cudaSuccess;
    uint8_t *d_out[S] = {};
    // pageable caller memory goes through the context's pinned ring (host_staging.hpp)
    // (small batches are left to the driver: a thread hand-off costs more than it saves below a few MiB)
    uint64_t batch_bytes = 0;
    for (const HostChunk &c : chunks) batch_bytes += c.sum;
    const bool stage_ok = !ctx->opt_no_staging && batch_bytes >= kStagingMinBytes;
    const bool src_pageable = stage_ok && host_pointer_is_pageable(src);
    const bool dst_pageable = stage_ok && host_pointer_is_pageable(dst);
    HostStaging *stg = (src_pageable || dst_pageable) ? staging_of(ctx) : nullptr;
    if ((src_pageable || dst_pageable) && !stg) return B2B_ECUDA;
    uint64_t rec_ticket[S] = {};
    // table block of a chunk of n frames (same layout on the device and in the pinned block):
    //   [src_off u64 n][frame_off u64 n][src_len u32 n][frame_len u32 n][status u32 n][total u64]
    auto layout = [](uint32_t n, uint64_t &a8, uint64_t &a4) { a8 = align_up(8ull * n, 256); a4 = align_up(4ull * n, 256); };

    // retire: the chunk's tables are on the host -> ship its packed bytes, publish its tables
    auto retire = [&](size_t k) -> int {
        const HostChunk &c = chunks[k];
        const int slot = (int)(k % S);
        const uint32_t n = c.f1 - c.f0;
        uint64_t a8, a4; layout(n, a8, a4);
        CU(ctx, cudaEventSynchronize(ctx->ev_tab[slot]));
        const uint8_t *t = ctx->ptab[slot];
        const uint64_t *h_frame_off = (const uint64_t *)(t + a8);
        const uint32_t *h_frame_len = (const uint32_t *)(t + 2 * a8 + a4);
        const uint32_t *h_status = (const uint32_t *)(t + 2 * a8 + 2 * a4);
        const uint64_t h_total = *(const uint64_t *)(t + 2 * a8 + 3 * a4);
        if (running + h_total > dst_cap) return B2B_EDST_TOO_SMALL;
        if (dst_pageable) {
            if (h_total) stg->d2h(hdst + running, d_out[slot], h_total);
            rec_ticket[slot] = stg->record(ctx->ev_out_free[slot]);
        } else {
            if (h_total) CU(ctx, cudaMemcpyAsync(hdst + running, d_out[slot], h_total, cudaMemcpyDeviceToHost, ctx->s_out));
            CU(ctx, cudaEventRecord(ctx->ev_out_free[slot], ctx->s_out));
        }
        for (uint32_t i = 0; i < n; i++) frame_off[c.f0 + i] = h_frame_off[i] + running;
        memcpy(frame_len + c.f0, h_frame_len, 4ull * n);
        memcpy(status + c.f0, h_status, 4ull * n);
        running += h_total;
        return B2B_OK;
    };

    size_t launched = 0, retired = 0;
    for (size_t k = 0; k < chunks.size() && rc == B2B_OK; k++) {
        const HostChunk &c = chunks[k];
        const int slot = (int)(k % S);
        const uint32_t n = c.f1 - c.f0;
        // scratch and output are sized for the bytes of all frames, which is more than the span when frames overlap
        const uint64_t span = c.hi - c.lo, work = std::max<uint64_t>(span, c.sum), out_cap = work + 31ull * n + 64;
        uint64_t a8, a4; layout(n, a8, a4);
        const uint64_t tab_bytes = 2 * a8 + 3 * a4 + 256;
        uint8_t *d_in = nullptr, *d_tab = nullptr, *h_tab = nullptr;
        // a slot is reused kSlots chunks later: that chunk must have been retired (its tables read)
        while (retired + S <= k && rc == B2B_OK) rc = retire(retired++);
        if (rc) break;
        rc = ensure_hbuf(ctx, 3 * slot + 0, span + 64, &d_in);
        if (rc == B2B_OK) rc = ensure_hbuf(ctx, 3 * slot + 1, out_cap, &d_out[slot]);
        if (rc == B2B_OK) rc = ensure_hbuf(ctx, 3 * slot + 2, tab_bytes, &d_tab);
        if (rc == B2B_OK) rc = ensure_ptab(ctx, slot, tab_bytes, &h_tab);
        if (rc) break;
        uint64_t *d_src_off = (uint64_t *)d_tab, *d_frame_off = (uint64_t *)(d_tab + a8);
        uint32_t *d_src_len = (uint32_t *)(d_tab + 2 * a8), *d_frame_len = (uint32_t *)(d_tab + 2 * a8 + a4);
        uint32_t *d_status = (uint32_t *)(d_tab + 2 * a8 + 2 * a4);
        uint64_t *d_total = (uint64_t *)(d_tab + 2 * a8 + 3 * a4);
        uint64_t *h_src_off = (uint64_t *)h_tab;
        uint32_t *h_src_len = (uint32_t *)(h_tab + 2 * a8);
        for (uint32_t i = 0; i < n; i++) h_src_off[i] = src_off[c.f0 + i] - c.lo;
        memcpy(h_src_len, src_len + c.f0, 4ull * n);
        // H2D (the kernels that last read this slot's input must be done)
        if (k >= (size_t)S) e = cudaStreamWaitEvent(ctx->s_in, ctx->ev_in_free[slot], 0);
        if (e == cudaSuccess) e = src_pageable ? stg->h2d(d_in, hsrc + c.lo, span)
                                               : cudaMemcpyAsync(d_in, hsrc + c.lo, span, cudaMemcpyHostToDevice, ctx->s_in);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_src_off, h_src_off, 8ull * n, cudaMemcpyHostToDevice, ctx->s_in);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_src_len, h_src_len, 4ull * n, cudaMemcpyHostToDevice, ctx->s_in);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_in_ready[slot], ctx->s_in);
        // kernels (the D2H that last read this slot's output must be done)
        cudaStream_t sk = ctx->s_k[slot];
        if (e == cudaSuccess) e = cudaStreamWaitEvent(sk, ctx->ev_in_ready[slot], 0);
        if (k >= (size_t)S && dst_pageable) stg->wait_recorded(rec_ticket[slot]);   // the drain thread has queued that chunk's copies
        if (e == cudaSuccess && k >= (size_t)S) e = cudaStreamWaitEvent(sk, ctx->ev_out_free[slot], 0);
        if (e != cudaSuccess) break;
        select_arena(ctx, 1 + slot);
        rc = blocks ? compress_blocks_dev_locked(ctx, d_in, d_src_off, d_src_len, n, work, c.max_len, shuffle, typesize,
                                                 blocksize, d_out[slot], out_cap, d_frame_off, d_frame_len, d_status,
                                                 d_total, sk)
                    : compress_batch_dev_locked(ctx, d_in, d_src_off, d_src_len, n, work, c.max_len, shuffle, typesize,
                                                d_out[slot], out_cap, d_frame_off, d_frame_len, d_status, d_total, sk);
        select_arena(ctx, 0);
        if (rc) break;
        e = cudaEventRecord(ctx->ev_done[slot], sk);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_in_free[slot], sk);
        // result tables -> pinned block
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_tab, ctx->ev_done[slot], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_tab + a8, d_tab + a8, a8 + 3 * a4 + 8 - 0, cudaMemcpyDeviceToHost, ctx->s_tab);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_tab[slot], ctx->s_tab);
        if (e != cudaSuccess) break;
        launched = k + 1;
        // while this chunk is in flight, ship the one before it
        while (retired + 1 < launched && rc == B2B_OK) rc = retire(retired++);
    }
    while (rc == B2B_OK && e == cudaSuccess && retired < launched) rc = retire(retired++);
    rc = sync_pipeline(ctx, e, rc);
    if (rc == B2B_OK && total_out) *total_out = running;
    return rc;
}

int b2b_compress_batch(b2b_ctx *ctx, const void *src, const uint64_t *src_off, const uint32_t *src_len,
                       uint32_t nframes, int shuffle, int64_t typesize, void *dst, uint64_t dst_cap,
                       uint64_t *frame_off, uint32_t *frame_len, uint32_t *status, uint64_t *total_out) {
    return host_compress_batch(ctx, src, src_off, src_len, nframes, shuffle, typesize, dst, dst_cap, frame_off,
                               frame_len, status, total_out, false, 0);
}

int b2b_compress_blocks_batch(b2b_ctx *ctx, const void *src, const uint64_t *src_off, const uint32_t *src_len,
                              uint32_t nframes, int shuffle, int64_t typesize, uint32_t blocksize, void *dst,
                              uint64_t dst_cap, uint64_t *frame_off, uint32_t *frame_len, uint32_t *status,
                              uint64_t *total_out) {
    if (blocksize != 0 && blocksize < kB1MinBuffer) return B2B_EINVAL;
    return host_compress_batch(ctx, src, src_off, src_len, nframes, shuffle, typesize, dst, dst_cap, frame_off,
                               frame_len, status, total_out, true, blocksize);
}

// Output slots may come in any order and with gaps between them: frames are processed in the order of their
// output offsets, slots that overlap are refused (B2B_EINVAL), and only the bytes a frame produced are ever
// written to dst (a gap between two slots, or the slot of a failed frame, keeps what the caller had there).
// Adjacent slots form a RUN that is laid out on the device exactly as in dst, so that a run of good frames
// goes back in one copy; the input side is gathered the same way (runs of nearby frames, one H2D each).
struct HostRun { uint64_t host, dev, len; uint32_t i0, i1; };

static int host_decompress_batch(b2b_ctx *ctx, const void *frames, const uint64_t *frame_off,
                                 const uint32_t *frame_len, uint32_t nframes, int64_t typesize_override,
                                 void *dst, uint64_t dst_cap, const uint64_t *dst_off, uint32_t *out_len,
                                 uint32_t *status, bool blocks, uint32_t blocksize) {
    NvtxRange nvtx_range("b2b.decompress_batch (host pipeline)");
    if (!ctx) return B2B_EINVAL;
    if (nframes == 0) return B2B_OK;
    if (!frames || !frame_off || !frame_len || !dst_off || !out_len || !status) return B2B_EINVAL;
    if (!dst && dst_cap) return B2B_EINVAL;
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    constexpr int S = b2b_ctx::kSlots;
    const uint8_t *hf = static_cast<const uint8_t *>(frames);
    uint8_t *hdst = static_cast<uint8_t *>(dst);
    // capacity of slot f: what the header announces, clipped to the caller's buffer and to what an
    // LZ4 block of that size can possibly produce (never allocate more than 255x the input)
    std::vector<uint32_t> cap(nframes);
    for (uint32_t f = 0; f < nframes; f++) {
        uint32_t norig = 0;
        if (frame_len[f] >= 16) {
            const uint8_t *p = hf + frame_off[f];
            norig = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
        }
        const uint64_t room = dst_off[f] <= dst_cap ? dst_cap - dst_off[f] : 0;
        const uint64_t reach = 255ull * frame_len[f] + 64;
        cap[f] = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(norig, room), reach);
    }
    // processing order = order of the output slots
    std::vector<uint32_t> ord;
    bool mono = true;
    for (uint32_t f = 1; f < nframes && mono; f++) mono = dst_off[f] >= dst_off[f - 1] + cap[f - 1];
    if (!mono) {
        ord.resize(nframes);
        for (uint32_t f = 0; f < nframes; f++) ord[f] = f;
        std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return dst_off[x] < dst_off[y]; });
        uint64_t end = 0;
        for (uint32_t i = 0; i < nframes; i++) {
            const uint32_t f = ord[i];
            if (cap[f] == 0) continue;
            if (dst_off[f] < end) return B2B_EINVAL;          // two output slots overlap
            end = dst_off[f] + cap[f];
        }
    }
    auto F = [&](uint32_t i) { return mono ? i : ord[i]; };
    // chunks by OUTPUT bytes (the larger side), in processing order
    struct Chunk { uint32_t i0, i1, max_cap; uint64_t out_bytes; };
    std::vector<Chunk> chunks;
    uint64_t cap_total = 0, cap_done = 0;
    for (uint32_t i = 0; i < nframes; i++) cap_total += cap[i];
    for (uint32_t i = 0; i < nframes;) {
        Chunk c{i, i, 0, 0};
        const uint64_t target = chunk_target(cap_done, cap_total, ctx->opt_stage_bytes);   // ramp up / down like split_chunks
        while (c.i1 < nframes && (c.i1 == c.i0 || c.out_bytes + cap[F(c.i1)] <= target)) {
            c.out_bytes += cap[F(c.i1)]; c.max_cap = std::max(c.max_cap, cap[F(c.i1)]); c.i1++;
        }
        cap_done += c.out_bytes;
        chunks.push_back(c);
        i = c.i1;
    }
    int rc = B2B_OK;
    cudaError_t e = cudaSuccess;
    std::vector<HostRun> out_runs[S];
    std::vector<uint64_t> dev_dst_off[S];
    uint8_t *d_out_slot[S] = {};
    uint64_t batch_bytes = 0;
    for (const Chunk &c : chunks) batch_bytes += c.out_bytes;
    const bool stage_ok = !ctx->opt_no_staging && batch_bytes >= kStagingMinBytes;
    const bool src_pageable = stage_ok && host_pointer_is_pageable(frames);
    const bool dst_pageable = stage_ok && dst && host_pointer_is_pageable(dst);
    HostStaging *stg = (src_pageable || dst_pageable) ? staging_of(ctx) : nullptr;
    if ((src_pageable || dst_pageable) && !stg) return B2B_ECUDA;
    uint64_t rec_ticket[S] = {};
    auto d2h = [&](void *h, const void *dv, uint64_t len) -> cudaError_t {
        if (dst_pageable) { stg->d2h(h, dv, len); return cudaSuccess; }
        return cudaMemcpyAsync(h, dv, len, cudaMemcpyDeviceToHost, ctx->s_out);
    };
    // table block: [frame_off u64 n][dst_off u64 n][frame_len u32 n][cap u32 n][out_len u32 n][status u32 n]
    auto retire = [&](size_t k) -> int {
        const Chunk &c = chunks[k];
        const int slot = (int)(k % S);
        const uint32_t n = c.i1 - c.i0;
        const uint64_t a8 = align_up(8ull * n, 256), a4 = align_up(4ull * n, 256);
        CU(ctx, cudaEventSynchronize(ctx->ev_tab[slot]));
        const uint8_t *t = ctx->ptab[slot];
        const uint32_t *h_out_len = (const uint32_t *)(t + 2 * a8 + 2 * a4);
        const uint32_t *h_status = (const uint32_t *)(t + 2 * a8 + 3 * a4);
        for (uint32_t j = 0; j < n; j++) { out_len[F(c.i0 + j)] = h_out_len[j]; status[F(c.i0 + j)] = h_status[j]; }
        // bytes back: a run whose frames all filled their slots goes in one copy, else frame by frame, and
        // only what a frame produced (status OK, or the short output of a size mismatch)
        for (const HostRun &r : out_runs[slot]) {
            bool whole = true;
            for (uint32_t i = r.i0; i < r.i1 && whole; i++)
                whole = h_status[i - c.i0] == B2B_OK && h_out_len[i - c.i0] == cap[F(i)];
            if (whole) {
                if (r.len) CU(ctx, d2h(hdst + r.host, d_out_slot[slot] + r.dev, r.len));
                continue;
            }
            for (uint32_t i = r.i0; i < r.i1; i++) {
                const uint32_t st = h_status[i - c.i0], got = std::min(h_out_len[i - c.i0], cap[F(i)]);
                if ((st == B2B_OK || st == B2B_ESIZE_MISMATCH) && got)
                    CU(ctx, d2h(hdst + dst_off[F(i)], d_out_slot[slot] + dev_dst_off[slot][i - c.i0], got));
            }
        }
        if (dst_pageable) rec_ticket[slot] = stg->record(ctx->ev_out_free[slot]);
        else CU(ctx, cudaEventRecord(ctx->ev_out_free[slot], ctx->s_out));
        return B2B_OK;
    };
    size_t launched = 0, retired = 0;
    std::vector<HostRun> in_runs;
    for (size_t k = 0; k < chunks.size() && rc == B2B_OK; k++) {
        const Chunk &c = chunks[k];
        const int slot = (int)(k % S);
        const uint32_t n = c.i1 - c.i0;
        while (retired + S <= k && rc == B2B_OK) rc = retire(retired++);
        if (rc) break;
        // input runs: frames that follow each other in `frames` (or nearly: a gap of up to 4 KiB is carried along)
        in_runs.clear();
        uint64_t in_bytes = 0;
        std::vector<uint64_t> dev_frame_off(n);
        for (uint32_t i = c.i0; i < c.i1; i++) {
            const uint64_t o = frame_off[F(i)], l = frame_len[F(i)];
            if (!in_runs.empty()) {
                HostRun &r = in_runs.back();
                if (o >= r.host && o <= r.host + r.len + 4096) {
                    r.len = std::max(r.len, o + l - r.host); r.i1 = i + 1;
                    dev_frame_off[i - c.i0] = r.dev + (o - r.host);
                    continue;
                }
                in_bytes = align_up(r.dev + r.len, 16);
            }
            in_runs.push_back(HostRun{o, in_bytes, l, i, i + 1});
            dev_frame_off[i - c.i0] = in_bytes;
        }
        if (!in_runs.empty()) in_bytes = in_runs.back().dev + in_runs.back().len;
        // output runs: slots that touch; the device image of a run keeps the layout it has in dst
        out_runs[slot].clear();
        dev_dst_off[slot].assign(n, 0);
        uint64_t out_bytes = 0;
        for (uint32_t i = c.i0; i < c.i1; i++) {
            const uint64_t o = dst_off[F(i)], l = cap[F(i)];
            if (!out_runs[slot].empty()) {
                HostRun &r = out_runs[slot].back();
                if (o == r.host + r.len) {
                    dev_dst_off[slot][i - c.i0] = r.dev + r.len; r.len += l; r.i1 = i + 1;
                    continue;
                }
                out_bytes = align_up(r.dev + r.len, 16);
            }
            out_runs[slot].push_back(HostRun{o, out_bytes, l, i, i + 1});
            dev_dst_off[slot][i - c.i0] = out_bytes;
        }
        if (!out_runs[slot].empty()) out_bytes = out_runs[slot].back().dev + out_runs[slot].back().len;
        const uint64_t a8 = align_up(8ull * n, 256), a4 = align_up(4ull * n, 256);
        const uint64_t tab_bytes = 2 * a8 + 4 * a4 + 256;
        uint8_t *d_in = nullptr, *d_out = nullptr, *d_tab = nullptr, *h_tab = nullptr;
        rc = ensure_hbuf(ctx, 3 * slot + 0, in_bytes + 64, &d_in);
        if (rc == B2B_OK) rc = ensure_hbuf(ctx, 3 * slot + 1, out_bytes + 64, &d_out);
        if (rc == B2B_OK) rc = ensure_hbuf(ctx, 3 * slot + 2, tab_bytes, &d_tab);
        if (rc == B2B_OK) rc = ensure_ptab(ctx, slot, tab_bytes, &h_tab);
        if (rc) break;
        d_out_slot[slot] = d_out;
        uint64_t *d_frame_off = (uint64_t *)d_tab, *d_dst_off = (uint64_t *)(d_tab + a8);
        uint32_t *d_frame_len = (uint32_t *)(d_tab + 2 * a8), *d_cap = (uint32_t *)(d_tab + 2 * a8 + a4);
        uint32_t *d_out_len = (uint32_t *)(d_tab + 2 * a8 + 2 * a4), *d_status = (uint32_t *)(d_tab + 2 * a8 + 3 * a4);
        uint64_t *h_frame_off = (uint64_t *)h_tab, *h_dst_off = (uint64_t *)(h_tab + a8);
        uint32_t *h_frame_len = (uint32_t *)(h_tab + 2 * a8), *h_cap = (uint32_t *)(h_tab + 2 * a8 + a4);
        for (uint32_t j = 0; j < n; j++) {
            h_frame_off[j] = dev_frame_off[j];
            h_dst_off[j] = dev_dst_off[slot][j];
            h_frame_len[j] = frame_len[F(c.i0 + j)];
            h_cap[j] = cap[F(c.i0 + j)];
        }
        if (k >= (size_t)S) e = cudaStreamWaitEvent(ctx->s_in, ctx->ev_in_free[slot], 0);
        for (const HostRun &r : in_runs)
            if (e == cudaSuccess && r.len) e = src_pageable ? stg->h2d(d_in + r.dev, hf + r.host, r.len)
                                                            : cudaMemcpyAsync(d_in + r.dev, hf + r.host, r.len, cudaMemcpyHostToDevice, ctx->s_in);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab, h_tab, 2 * a8 + 2 * a4, cudaMemcpyHostToDevice, ctx->s_in);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_in_ready[slot], ctx->s_in);
        cudaStream_t sk = ctx->s_k[slot];
        if (e == cudaSuccess) e = cudaStreamWaitEvent(sk, ctx->ev_in_ready[slot], 0);
        if (k >= (size_t)S && dst_pageable) stg->wait_recorded(rec_ticket[slot]);   // the drain thread has queued that chunk's copies
        if (e == cudaSuccess && k >= (size_t)S) e = cudaStreamWaitEvent(sk, ctx->ev_out_free[slot], 0);
        if (e != cudaSuccess) break;
        select_arena(ctx, 1 + slot);
        rc = blocks ? decompress_blocks_dev_locked(ctx, d_in, d_frame_off, d_frame_len, n, d_out, d_dst_off, d_cap,
                                                   out_bytes, c.max_cap, blocksize, d_out_len, d_status, sk)
                    : decompress_batch_dev_locked(ctx, d_in, d_frame_off, d_frame_len, n, typesize_override, d_out,
                                                  d_dst_off, d_cap, out_bytes, c.max_cap, d_out_len, d_status, sk);
        select_arena(ctx, 0);
        if (rc) break;
        e = cudaEventRecord(ctx->ev_done[slot], sk);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_in_free[slot], sk);
        // result tables -> pinned block; the bytes follow when the tables are on the host (retire)
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_tab, ctx->ev_done[slot], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_tab + 2 * a8 + 2 * a4, d_tab + 2 * a8 + 2 * a4, 2 * a4, cudaMemcpyDeviceToHost, ctx->s_tab);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_tab[slot], ctx->s_tab);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[slot], 0);
        if (e != cudaSuccess) break;
        launched = k + 1;
        // while this chunk is in flight, ship the one before it
        while (retired + 1 < launched && rc == B2B_OK) rc = retire(retired++);
    }
    while (rc == B2B_OK && e == cudaSuccess && retired < launched) rc = retire(retired++);
    return sync_pipeline(ctx, e, rc);
}

int b2b_decompress_batch(b2b_ctx *ctx, const void *frames, const uint64_t *frame_off,
                         const uint32_t *frame_len, uint32_t nframes, int64_t typesize_override,
                         void *dst, uint64_t dst_cap, const uint64_t *dst_off, uint32_t *out_len,
                         uint32_t *status) {
    return host_decompress_batch(ctx, frames, frame_off, frame_len, nframes, typesize_override, dst, dst_cap,
                                 dst_off, out_len, status, false, 0);
}

int b2b_decompress_blocks_batch(b2b_ctx *ctx, const void *frames, const uint64_t *frame_off,
                                const uint32_t *frame_len, uint32_t nframes, uint32_t blocksize, void *dst,
                                uint64_t dst_cap, const uint64_t *dst_off, uint32_t *out_len, uint32_t *status) {
    if (blocksize != 0 && blocksize < kB1MinBuffer) return B2B_EINVAL;
    return host_decompress_batch(ctx, frames, frame_off, frame_len, nframes, 0, dst, dst_cap, dst_off, out_len,
                                 status, true, blocksize);
}

int b2b_compress(b2b_ctx *ctx, const void *src, size_t n, int codec, int level, int shuffle,
                 int64_t typesize, void *dst, size_t cap, size_t *out_len) {
    (void)level;  // clamped at blosc.go:277-282, never read by the LZ4 adapter (codec.go:63)
    if (!ctx || !out_len) return B2B_EINVAL;
    if (n == 0) return B2B_EINVALID_DATA;                                  // blosc.go:269-271
    if (codec <= B2B_BLOSCLZ || codec > B2B_ZSTD) return B2B_EINVALID_CODEC;  // blosc.go:322-325
    if (codec != B2B_LZ4) return B2B_EUNSUPPORTED;
    if (n > 0xFFFFFFFFull - 16ull) return B2B_EDATA_TOO_LARGE;
    if (!src || !dst) return B2B_EINVAL;
    const uint64_t off = 0; const uint32_t len = (uint32_t)n;
    uint64_t foff = 0, total = 0; uint32_t flen = 0, st = 0;
    std::vector<uint8_t> tmp;
    void *out = dst; uint64_t out_cap = cap;
    if (cap < n + 16 + 64) { tmp.resize(n + 16 + 64); out = tmp.data(); out_cap = tmp.size(); }
    int rc = b2b_compress_batch(ctx, src, &off, &len, 1, shuffle, typesize, out, out_cap, &foff, &flen, &st, &total);
    if (rc) return rc;
    if (st) return (int)st;
    if (flen > cap) return B2B_EDST_TOO_SMALL;
    if (out != dst) memcpy(dst, tmp.data(), flen);
    *out_len = flen;
    return B2B_OK;
}

int b2b_decompress(b2b_ctx *ctx, const void *frame, size_t len, int64_t typesize_override, void *dst,
                   size_t cap, size_t *out_len) {
    if (!ctx || !out_len) return B2B_EINVAL;
    b2b_header h;
    if (len < B2B_HEADER_SIZE) return B2B_EINVALID_HEADER;                 // blosc.go:297-299
    if (!frame) return B2B_EINVAL;
    int rc = b2b_parse_header(frame, len, &h);
    if (rc) return rc;
    if ((uint64_t)h.nbytes_comp > len || h.nbytes_comp < B2B_HEADER_SIZE) return B2B_EINVALID_DATA;
    if (len > 0xFFFFFFFFull) len = h.nbytes_comp;  // bytes past NBytesComp are ignored anyway
    const uint64_t foff = 0, doff = 0; const uint32_t flen = (uint32_t)len;
    uint32_t got = 0, st = 0;
    // cap < NBytesOrig surfaces as EDST_TOO_SMALL from the kernel, after the reference's own checks
    rc = b2b_decompress_batch(ctx, frame, &foff, &flen, 1, typesize_override, dst, cap, &doff, &got, &st);
    if (rc) return rc;
    if (st) return (int)st;
    *out_len = got;
    return B2B_OK;
}

int b2b_lz4_block_compress(b2b_ctx *ctx, const void *src, size_t n, void *dst, size_t cap, size_t *out_len) {
    if (!ctx || !out_len || (!src && n) || !dst) return B2B_EINVAL;
    if (n > 0xFFFFFFFFull - 16ull) return B2B_EDATA_TOO_LARGE;
    if (cap < b2b_lz4_bound(n)) return B2B_EDST_TOO_SMALL;
    if (n == 0) {  // an empty block is the single token 0x00
        static_cast<uint8_t *>(dst)[0] = 0; *out_len = 1; return B2B_OK;
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t out_cap = align_up(b2b_lz4_bound(n) + 64, 256);
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_tab = nullptr;
    int rc = ensure_hbuf(ctx, 0, n + 64, &d_in);
    if (rc == B2B_OK) rc = ensure_hbuf(ctx, 1, out_cap, &d_out);
    if (rc == B2B_OK) rc = ensure_hbuf(ctx, 2, 4096, &d_tab);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    uint64_t *d_off = reinterpret_cast<uint64_t *>(d_tab);            // [0] src_off, [1] frame_off, [2] total
    uint32_t *d_u32 = reinterpret_cast<uint32_t *>(d_tab + 256);      // [0] src_len, [1] frame_len, [2] status
    const uint64_t h_off[3] = {0, 0, 0};
    const uint32_t h_u32[3] = {(uint32_t)n, 0, 0};
    CU(ctx, cudaMemcpyAsync(d_in, src, n, cudaMemcpyHostToDevice, s));
    CU(ctx, cudaMemcpyAsync(d_off, h_off, sizeof h_off, cudaMemcpyHostToDevice, s));
    CU(ctx, cudaMemcpyAsync(d_u32, h_u32, sizeof h_u32, cudaMemcpyHostToDevice, s));
    rc = compress_batch_dev_locked(ctx, d_in, d_off, d_u32, 1, n, (uint32_t)n, B2B_NOSHUFFLE, 1, d_out, out_cap,
                                   d_off + 1, d_u32 + 1, d_u32 + 2, d_off + 2, s, /*raw_block=*/true);
    if (rc) return rc;
    uint32_t h_res[3] = {0, 0, 0};
    CU(ctx, cudaMemcpyAsync(h_res, d_u32, sizeof h_res, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    if (h_res[2]) return (int)h_res[2];
    const uint32_t c = h_res[1] - 16;   // frame_len counts a header that the raw API does not write
    CU(ctx, cudaMemcpyAsync(dst, d_out, c, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    *out_len = c;
    return B2B_OK;
}

int b2b_lz4_block_decompress(b2b_ctx *ctx, const void *src, size_t n, void *dst, size_t expected,
                             size_t *out_len) {
    if (!ctx || !out_len || (!src && n) || (!dst && expected)) return B2B_EINVAL;
    if (n > 0xFFFFFFFFull - 32ull || expected > 0xFFFFFFFFull) return B2B_EDATA_TOO_LARGE;
    // wrap the block in a frame header and reuse the frame decoder
    std::vector<uint8_t> fr(16 + n);
    b2b_header h = {2, B2B_LZ4, 0, 1, (uint32_t)expected, (uint32_t)expected, (uint32_t)(16 + n)};
    b2b_header_bytes(&h, fr.data());
    if (n) memcpy(fr.data() + 16, src, n);
    const uint64_t foff = 0, doff = 0; const uint32_t flen = (uint32_t)fr.size();
    uint32_t got = 0, st = 0;
    int rc = b2b_decompress_batch(ctx, fr.data(), &foff, &flen, 1, 0, dst, expected, &doff, &got, &st);
    if (rc) return rc;
    // codec.go:77-84 returns buf[:n]: a short decode is not an error at this level
    if (st && st != B2B_ESIZE_MISMATCH) return (int)st;
    *out_len = got;
    return B2B_OK;
}

}  // extern "C"
