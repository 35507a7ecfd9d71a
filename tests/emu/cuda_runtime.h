// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.  A shim that lets g++ compile the device code
// under go-blosc_b200/csrc/*.cuh unchanged and run it on the CPU, one coroutine per CUDA thread, so
// that kernel LOGIC (match finding, emission, token parsing, status rules) can be exercised and
// instrumented without a GPU.  It is found instead of the real <cuda_runtime.h> only when a test tool
// is built with -I tests/emu; nothing in the product includes, links or executes it, and the product
// library has no CPU path (b2b_init fails without a device).
//
// Execution model: a kernel launch runs its CTAs one after the other; the threads of a CTA are
// coroutines scheduled round-robin.  A thread runs until it reaches a warp / CTA collective
// (__shfl_sync, __ballot_sync, __syncwarp, __syncthreads, ...), where it waits for the other live
// threads of its warp / CTA.  Between two collectives the lanes of a warp therefore run one after the
// other in lane order instead of instruction by instruction; code that relies on lockstep between two
// collectives (none of ours does on purpose; benign hash-table races resolve in lane order) can differ.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define B2B_EMU 1

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
typedef void *cudaStream_t;

namespace emu {

struct Warp {
    uint64_t in[2][32];
    uint32_t arrived = 0, active = 0, gen = 0;
};
struct Cta {
    uint32_t arrived = 0, live = 0, gen = 0;
    uint32_t red[2] = {0, 0};   // predicate counts of __syncthreads_count / _or, by generation parity
};
struct Thread {
    ucontext_t ctx;
    uint3 tidx;
    int lane = 0;
    Warp *warp = nullptr;
    Cta *cta = nullptr;
    bool done = false;
    void *stack = nullptr;
};

extern Thread *cur;           // the running CUDA thread
extern ucontext_t sched_ctx;  // the scheduler
extern uint3 g_blockIdx, g_blockDim, g_gridDim;
extern uint64_t n_collectives, n_switches;

inline void yield() { n_switches++; swapcontext(&cur->ctx, &sched_ctx); }

// every live lane of the warp contributes v; returns the contributions of this collective
inline const uint64_t *warp_collective(uint64_t v) {
    Warp *w = cur->warp;
    const uint32_t g = w->gen;
    w->in[g & 1][cur->lane] = v;
    w->arrived |= 1u << cur->lane;
    n_collectives++;
    while (w->gen == g) {
        if (w->arrived == w->active) { w->arrived = 0; w->gen = g + 1; break; }
        yield();
    }
    return w->in[g & 1];
}
inline uint32_t live_mask() { return cur->warp->active; }

void launch(uint32_t grid, uint32_t block, const std::function<void()> &body);

}  // namespace emu

#define threadIdx (emu::cur->tidx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

// ---- collectives (full masks only: that is all the kernels use) ------------------------------------
static inline void __syncwarp(uint32_t = 0xffffffffu) { emu::warp_collective(0); }
static inline void __syncthreads() {
    emu::Cta *c = emu::cur->cta;
    const uint32_t g = c->gen;
    c->arrived++;
    while (c->gen == g) {
        if (c->arrived == c->live) { c->arrived = 0; c->red[(g + 1) & 1] = 0; c->gen = g + 1; emu::n_collectives++; break; }
        emu::yield();
    }
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
// CTA-wide barrier + reduction of a predicate
static inline int emu_syncthreads_reduce(int pred, int op) {
    emu::Cta *c = emu::cur->cta;
    const uint32_t g = c->gen;
    if (pred) c->red[g & 1]++;
    c->arrived++;
    while (c->gen == g) {
        if (c->arrived == c->live) { c->arrived = 0; c->red[(g + 1) & 1] = 0; c->gen = g + 1; emu::n_collectives++; break; }
        emu::yield();
    }
    const int n = (int)c->red[g & 1];
    return op == 0 ? n : (n != 0);
}
static inline int __syncthreads_count(int pred) { return emu_syncthreads_reduce(pred, 0); }
static inline int __syncthreads_or(int pred) { return emu_syncthreads_reduce(pred, 1); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline uint32_t __ballot_sync(uint32_t, int pred) {
    const uint32_t live = emu::live_mask();
    const uint64_t *in = emu::warp_collective(pred ? 1 : 0);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && in[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(uint32_t m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(uint32_t m, int pred) { return __ballot_sync(m, !pred) == 0; }
template <typename T> static inline T __shfl_sync(uint32_t, T v, int src, int = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const uint64_t *in = emu::warp_collective(raw);
    T out; memcpy(&out, &in[src & 31], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_up_sync(uint32_t, T v, unsigned d, int = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const int lane = emu::cur->lane;
    const uint64_t *in = emu::warp_collective(raw);
    T out = v;
    if (lane >= (int)d) memcpy(&out, &in[lane - d], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_down_sync(uint32_t, T v, unsigned d, int = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const int lane = emu::cur->lane;
    const uint64_t *in = emu::warp_collective(raw);
    T out = v;
    if (lane + (int)d < 32) memcpy(&out, &in[lane + d], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_xor_sync(uint32_t, T v, int x, int = 32) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    const int lane = emu::cur->lane;
    const uint64_t *in = emu::warp_collective(raw);
    T out; memcpy(&out, &in[(lane ^ x) & 31], sizeof(T));
    return out;
}
static inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) {
    const uint32_t live = emu::live_mask();
    const uint64_t *in = emu::warp_collective(v);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = (uint32_t)in[l] > r ? (uint32_t)in[l] : r;
    return r;
}
static inline uint32_t __reduce_min_sync(uint32_t, uint32_t v) {
    const uint32_t live = emu::live_mask();
    const uint64_t *in = emu::warp_collective(v);
    uint32_t r = 0xFFFFFFFFu;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = (uint32_t)in[l] < r ? (uint32_t)in[l] : r;
    return r;
}
static inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) {
    const uint32_t live = emu::live_mask();
    const uint64_t *in = emu::warp_collective(v);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r += (uint32_t)in[l];
    return r;
}
static inline uint32_t __reduce_or_sync(uint32_t, uint32_t v) {
    const uint32_t live = emu::live_mask();
    const uint64_t *in = emu::warp_collective(v);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r |= (uint32_t)in[l];
    return r;
}
static inline uint32_t __match_any_sync(uint32_t, uint32_t v) {
    const uint32_t live = emu::live_mask();
    const uint64_t *in = emu::warp_collective(v);
    uint32_t r = 0;
    for (int l = 0; l < 32; l++) if (((live >> l) & 1u) && (uint32_t)in[l] == v) r |= 1u << l;
    return r;
}

// ---- scalar intrinsics -----------------------------------------------------------------------------
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s & 31u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
    return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (s & 31u)) >> 32);
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t __brev(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t ab = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) {
        const uint32_t s = (sel >> (4 * k)) & 0xF;
        uint32_t byte = (uint32_t)((ab >> (8 * (s & 7))) & 0xFF);
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * k);
    }
    return r;
}
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
