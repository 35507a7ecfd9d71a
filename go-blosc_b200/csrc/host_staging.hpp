// host_staging.hpp -- pinned staging for PAGEABLE caller buffers of the host-pointer batch calls.
//
// The reference's API hands over ordinary Go slices (blosc.go:257-265, 291-303), i.e. pageable memory.
// cudaMemcpyAsync on pageable memory is synchronous and staged by the driver one small buffer at a time, so
// the H2D / kernels / D2H pipeline of b2b.cu would not overlap for the callers the drop-in actually has.
// Here the library owns the staging:
//   H2D   the calling thread copies 8 MiB blocks into a ring of pinned buffers (the copy itself is spread
//         over a small pool of host threads) and queues one cudaMemcpyAsync per block;
//   D2H   a drain thread queues one cudaMemcpyAsync per block into its own pinned ring and moves a block
//         to the caller's buffer as soon as its event completes, while the next block is in flight.
// One cudaMemcpyAsync per block in each direction; no batched-copy APIs.  Host only (no kernels here).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace b2b {

// Large copies between pageable and pinned memory: both sides are touched once, so the destination is written
// with streaming stores (no read-for-ownership of lines that are overwritten whole: a third less DRAM traffic,
// and the pinned side is read by the DMA engine from DRAM anyway).
#if defined(__x86_64__)
__attribute__((target("avx2"))) static inline void stream_copy_avx2(uint8_t *d, const uint8_t *s, size_t n) {
    size_t head = (32 - ((uintptr_t)d & 31)) & 31;
    if (head > n) head = n;
    memcpy(d, s, head); d += head; s += head; n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(s + i)), b = _mm256_loadu_si256((const __m256i *)(s + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(s + i + 64)), e = _mm256_loadu_si256((const __m256i *)(s + i + 96));
        _mm256_stream_si256((__m256i *)(d + i), a); _mm256_stream_si256((__m256i *)(d + i + 32), b);
        _mm256_stream_si256((__m256i *)(d + i + 64), c); _mm256_stream_si256((__m256i *)(d + i + 96), e);
    }
    _mm_sfence();
    memcpy(d + i, s + i, n - i);
}
#endif
static inline void stream_copy(void *d, const void *s, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("B2B_STAGE_MEMCPY");
    if (avx2 && n >= 4096) { stream_copy_avx2((uint8_t *)d, (const uint8_t *)s, n); return; }
#endif
    memcpy(d, s, n);
}

// a few host threads that split large memcpy calls between them (callers may come from several threads)
class CopyPool {
public:
    explicit CopyPool(int nthreads) {
        for (int i = 0; i < nthreads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int threads() const { return (int)workers_.size(); }
    void copy(void *dst, const void *src, size_t n) {
        const size_t kMin = 512u << 10;
        const size_t parts = workers_.empty() ? 1 : std::min<size_t>(workers_.size() + 1, (n + kMin - 1) / kMin);
        if (parts <= 1) { stream_copy(dst, src, n); return; }
        const size_t piece = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
        std::atomic<int> left{0};
        std::mutex dm; std::condition_variable dcv;
        size_t mine_off = 0, mine_len = 0;
        {
            std::lock_guard<std::mutex> g(mu_);
            for (size_t off = 0; off < n; off += piece) {
                const size_t len = std::min(piece, n - off);
                if (off == 0) { mine_off = off; mine_len = len; continue; }
                left++;
                tasks_.push_back([=, &left, &dm, &dcv] {
                    stream_copy((uint8_t *)dst + off, (const uint8_t *)src + off, len);
                    if (--left == 0) { std::lock_guard<std::mutex> g2(dm); dcv.notify_one(); }
                });
            }
        }
        cv_.notify_all();
        stream_copy((uint8_t *)dst + mine_off, (const uint8_t *)src + mine_off, mine_len);
        std::unique_lock<std::mutex> lk(dm);
        dcv.wait(lk, [&] { return left.load() == 0; });
    }
private:
    void loop() {
        for (;;) {
            std::function<void()> t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !tasks_.empty(); });
                if (tasks_.empty()) return;
                t = std::move(tasks_.front()); tasks_.pop_front();
            }
            t();
        }
    }
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> tasks_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

class HostStaging {
public:
    size_t kBlock = 8u << 20;          // bytes of one staging block (B2B_STAGE_BLOCK_MB overrides: experiments)
    static constexpr int kRing = 3;

    HostStaging(int device, int nthreads, cudaStream_t s_in, cudaStream_t s_out)
        : device_(device), pool_(nthreads), s_in_(s_in), s_out_(s_out) {
        if (const char *e = getenv("B2B_STAGE_BLOCK_MB")) { const long v = atol(e); if (v >= 1 && v <= 256) kBlock = (size_t)v << 20; }
    }
    ~HostStaging() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        if (drain_.joinable()) drain_.join();
        for (int i = 0; i < kRing; i++) {
            if (in_[i]) cudaFreeHost(in_[i]);
            if (out_[i]) cudaFreeHost(out_[i]);
            if (in_ev_[i]) cudaEventDestroy(in_ev_[i]);
            if (out_ev_[i]) cudaEventDestroy(out_ev_[i]);
        }
    }
    int threads() const { return pool_.threads(); }

    // pageable host -> device through the pinned ring, on s_in (calling thread)
    cudaError_t h2d(void *dev, const void *host, size_t n) {
        cudaError_t e = ensure(in_, in_ev_);
        for (size_t off = 0; off < n && e == cudaSuccess; off += kBlock) {
            const size_t len = std::min(kBlock, n - off);
            const int i = (int)(in_seq_++ % kRing);
            e = cudaEventSynchronize(in_ev_[i]);                       // the DMA that last read this block is done
            if (e != cudaSuccess) break;
            pool_.copy(in_[i], (const uint8_t *)host + off, len);
            e = cudaMemcpyAsync((uint8_t *)dev + off, in_[i], len, cudaMemcpyHostToDevice, s_in_);
            if (e == cudaSuccess) e = cudaEventRecord(in_ev_[i], s_in_);
        }
        return e;
    }

    // device -> pageable host: queued for the drain thread (s_out).  record(): an event on s_out behind
    // everything queued so far; returns a ticket for wait_recorded().
    void d2h(void *host, const void *dev, size_t n) { push(Job{0, host, dev, n, nullptr}); }
    uint64_t record(cudaEvent_t ev) { push(Job{1, nullptr, nullptr, 0, ev}); return ++records_pushed_; }
    void wait_recorded(uint64_t ticket) {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return records_issued_ >= ticket || err_ != cudaSuccess; });
    }
    // everything queued has reached the caller's memory
    cudaError_t flush() {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return (jobs_.empty() && !busy_) || err_ != cudaSuccess; });
        const cudaError_t e = err_;
        if (e != cudaSuccess) { jobs_.clear(); err_ = cudaSuccess; }
        return e;
    }

private:
    struct Job { int kind; void *host; const void *dev; size_t n; cudaEvent_t ev; };

    cudaError_t ensure(uint8_t **ring, cudaEvent_t *ev) {
        for (int i = 0; i < kRing; i++) {
            if (!ring[i]) {
                cudaError_t e = cudaHostAlloc((void **)&ring[i], kBlock, cudaHostAllocDefault);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    }
    void push(const Job &j) {
        {
            std::lock_guard<std::mutex> g(mu_);
            if (!drain_.joinable()) drain_ = std::thread([this] { drain_loop(); });
            jobs_.push_back(j);
        }
        cv_.notify_all();
    }
    void fail(cudaError_t e) {
        std::lock_guard<std::mutex> g(mu_);
        if (err_ == cudaSuccess) err_ = e;
        done_cv_.notify_all();
    }
    void drain_loop() {
        cudaSetDevice(device_);
        cudaError_t e = ensure(out_, out_ev_);
        if (e != cudaSuccess) fail(e);
        struct Pending { void *host; int slot; size_t len; };
        std::deque<Pending> pend;
        auto complete_one = [&]() {
            const Pending p = pend.front(); pend.pop_front();
            const cudaError_t e2 = cudaEventSynchronize(out_ev_[p.slot]);
            if (e2 != cudaSuccess) { fail(e2); return; }
            pool_.copy(p.host, out_[p.slot], p.len);
        };
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (jobs_.empty() && pend.empty()) { busy_ = false; done_cv_.notify_all(); }
                if (jobs_.empty() && !pend.empty()) { lk.unlock(); complete_one(); continue; }   // nothing new: finish what is in flight
                cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (jobs_.empty()) return;
                j = jobs_.front(); jobs_.pop_front();
                busy_ = true;
            }
            if (j.kind == 1) {
                const cudaError_t e2 = cudaEventRecord(j.ev, s_out_);
                if (e2 != cudaSuccess) fail(e2);
                std::lock_guard<std::mutex> g(mu_);
                records_issued_++;
                done_cv_.notify_all();
                continue;
            }
            for (size_t off = 0; off < j.n; off += kBlock) {
                const size_t len = std::min(kBlock, j.n - off);
                if ((int)pend.size() == kRing) complete_one();          // the oldest block leaves its pinned buffer
                const int i = (int)(out_seq_++ % kRing);
                cudaError_t e2 = cudaMemcpyAsync(out_[i], (const uint8_t *)j.dev + off, len, cudaMemcpyDeviceToHost, s_out_);
                if (e2 == cudaSuccess) e2 = cudaEventRecord(out_ev_[i], s_out_);
                if (e2 != cudaSuccess) { fail(e2); break; }
                pend.push_back(Pending{(uint8_t *)j.host + off, i, len});
                if (pend.size() > 1) complete_one();                    // move the block before this one while this one is in flight
            }
        }
    }

    int device_;
    CopyPool pool_;
    cudaStream_t s_in_, s_out_;
    uint8_t *in_[kRing] = {}, *out_[kRing] = {};
    cudaEvent_t in_ev_[kRing] = {}, out_ev_[kRing] = {};
    uint64_t in_seq_ = 0, out_seq_ = 0;
    std::thread drain_;
    std::deque<Job> jobs_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    bool stop_ = false, busy_ = false;
    cudaError_t err_ = cudaSuccess;
    uint64_t records_pushed_ = 0, records_issued_ = 0;
};

inline bool host_pointer_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace b2b
