// CPU emulation runtime for the melvin-b200 kernels -- DEVELOPMENT / TEST HARNESS ONLY.
//
// Compiles the unmodified kernel sources (-DMLV_EMU) for the host so that their
// index logic can be checked in a container without a GPU.  Each CTA runs to
// completion before the next one starts; the threads of a CTA are ucontext
// fibers on one OS thread and __syncthreads() yields to a round-robin scheduler.
// This library is never loaded by the product package and is not a fallback:
// it exists so that tests/emu/test_emu_kernels.py can exercise the C ABI here.
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#include "../../melvin.py_b200/csrc/mlv_common.cuh"

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
unsigned char* emu_smem_base = nullptr;

namespace {
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
};
ucontext_t g_sched;
std::vector<Fiber> g_fibers;
int g_cur = -1;
const std::function<void()>* g_body = nullptr;
constexpr size_t kStack = 256 * 1024;

void trampoline() {
    (*g_body)();
    g_fibers[g_cur].done = true;
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}
}  // namespace

// ---- barriers: counting, generation based (bar.sync / bar.arrive semantics)
namespace {
constexpr int kMaxBar = 64;     // 0..15 named barriers, 16.. one per warp (warp_sync)
int g_bar_cnt[kMaxBar];
unsigned g_bar_gen[kMaxBar];
void yield_fiber() { swapcontext(&g_fibers[g_cur].ctx, &g_sched); }
}  // namespace

void emu_yield() { yield_fiber(); }

void emu_bar_arrive(int id, int count) {
    if (++g_bar_cnt[id] == count) { g_bar_cnt[id] = 0; ++g_bar_gen[id]; }
}

void emu_bar_sync(int id, int count) {
    const unsigned gen = g_bar_gen[id];
    if (++g_bar_cnt[id] == count) { g_bar_cnt[id] = 0; ++g_bar_gen[id]; return; }
    while (g_bar_gen[id] == gen) yield_fiber();
}

void __syncthreads() { emu_bar_sync(0, (int)blockDim.x); }

namespace mlv {
void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body) {
    g_body = &body;
    gridDim.x = grid;
    blockDim.x = block;
    std::vector<unsigned char> sm(smem + 64);
    if (g_fibers.size() < block) g_fibers.resize(block);
    for (unsigned t = 0; t < block; ++t)
        if (!g_fibers[t].stack) g_fibers[t].stack = (char*)malloc(kStack);
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        // poison shared memory so that reads of never-written slots show up as NaN
        for (size_t i = 0; i + 8 <= sm.size(); i += 8) {
            const unsigned long long nanbits = 0x7ff8dead00000000ULL;
            memcpy(&sm[i], &nanbits, 8);
        }
        emu_smem_base = (unsigned char*)(((uintptr_t)sm.data() + 15) & ~(uintptr_t)15);
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = g_fibers[t];
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
        for (int i = 0; i < kMaxBar; ++i) { g_bar_cnt[i] = 0; g_bar_gen[i] = 0; }
        bool all_done = false;
        while (!all_done) {
            all_done = true;
            for (unsigned t = 0; t < block; ++t) {
                if (g_fibers[t].done) continue;
                g_cur = (int)t;
                threadIdx.x = t;
                swapcontext(&g_sched, &g_fibers[t].ctx);
                if (!g_fibers[t].done) all_done = false;
            }
        }
    }
    g_body = nullptr;
}
}  // namespace mlv
