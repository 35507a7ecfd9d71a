// tests/emu/emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the
// coroutine scheduler behind the CUDA-on-CPU shim.
#include "cuda_runtime.h"

#include <stdio.h>

namespace emu {

Thread *cur = nullptr;
ucontext_t sched_ctx;
uint3 g_blockIdx, g_blockDim, g_gridDim;
uint64_t n_collectives = 0, n_switches = 0;

static const std::function<void()> *g_body = nullptr;
static constexpr size_t kStack = 512 * 1024;

static void trampoline() {
    (*g_body)();
    Thread *t = cur;
    t->done = true;
    t->warp->active &= ~(1u << t->lane);
    t->cta->live--;
    swapcontext(&t->ctx, &sched_ctx);
}

void launch(uint32_t grid, uint32_t block, const std::function<void()> &body) {
    g_body = &body;
    g_gridDim = uint3{grid, 1, 1};
    g_blockDim = uint3{block, 1, 1};
    std::vector<Thread> th(block);
    std::vector<Warp> warps((block + 31) / 32);
    for (uint32_t t = 0; t < block; t++) {
        if (posix_memalign(&th[t].stack, 64, kStack)) abort();
    }
    for (uint32_t b = 0; b < grid; b++) {
        g_blockIdx = uint3{b, 0, 0};
        Cta cta;
        cta.live = block;
        for (auto &w : warps) w = Warp();
        for (uint32_t t = 0; t < block; t++) {
            Thread &T = th[t];
            T.tidx = uint3{t, 0, 0};
            T.lane = (int)(t & 31);
            T.warp = &warps[t >> 5];
            T.cta = &cta;
            T.done = false;
            T.warp->active |= 1u << T.lane;
            getcontext(&T.ctx);
            T.ctx.uc_stack.ss_sp = T.stack;
            T.ctx.uc_stack.ss_size = kStack;
            T.ctx.uc_link = nullptr;
            makecontext(&T.ctx, trampoline, 0);
        }
        uint32_t alive = block;
        uint64_t idle_rounds = 0;
        while (alive) {
            const uint64_t before = n_collectives;
            uint32_t finished = 0;
            for (uint32_t t = 0; t < block; t++) {
                if (th[t].done) continue;
                cur = &th[t];
                swapcontext(&sched_ctx, &th[t].ctx);
                if (th[t].done) finished++;
            }
            alive -= finished;
            // a whole round without any thread reaching a collective or finishing: the kernel waits for
            // something that can never arrive (a divergent collective)
            if (finished == 0 && n_collectives == before) {
                if (++idle_rounds > 4) { fprintf(stderr, "emu: deadlock in CTA %u (divergent collective?)\n", b); abort(); }
            } else idle_rounds = 0;
        }
    }
    for (uint32_t t = 0; t < block; t++) free(th[t].stack);
    cur = nullptr;
}

}  // namespace emu
