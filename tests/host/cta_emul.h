// TEST INFRASTRUCTURE: one CUDA thread block emulated on the host, so that a kernel's own source
// (csrc/hqr_mb.cuh: the multi-bulge QR kernel that takes 60 % of the sweep) runs without a GPU.
//
// Every thread of the CTA is a cooperative fiber (ucontext) with its own stack; __syncthreads() is a
// rendezvous of all threads that have not returned, warp collectives (__syncwarp, __ballot_sync,
// __shfl_sync) are rendezvous of the 32 threads of a warp.  Between two rendezvous a thread runs
// undisturbed and the threads run one after the other - in a new pseudo-random order every scheduling
// round when CTA_SEED is set, which turns a missing barrier (a shared-memory race) into a result that
// depends on the seed.  `static` stands in for `__shared__` (one CTA at a time), a plain buffer for the
// dynamic shared memory.  Not a performance model: it checks the logic - indices, barrier placement,
// deflation and convergence tests - against LAPACK (tests/test_host.py::test_hqr_kernel_on_emulated_cta).
#pragma once
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

namespace cta {
constexpr int MAX_THREADS = 1024;
struct Rendezvous { int arrived = 0; unsigned gen = 0; };
struct Uint3 { unsigned x, y, z; };

// Context switch.  x86-64: a dozen instructions (callee-saved registers + stack pointer) - glibc's swapcontext
// makes a system call per switch (signal mask), which dominates an emulation that switches ~10^8 times;
// elsewhere: ucontext.
#if defined(__x86_64__) && !defined(CTA_EMUL_UCONTEXT)
#define CTA_EMUL_FAST_SWITCH 1
extern "C" void cta_emul_switch(void** save_sp, void* load_sp);
asm(".text\n"
    ".globl cta_emul_switch\n"
    ".type cta_emul_switch,@function\n"
    "cta_emul_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size cta_emul_switch,.-cta_emul_switch\n");
struct Context { void* sp; };
#else
struct Context { ucontext_t uc; };
#endif

struct State {
    Context main_ctx;
    std::vector<Context> ctx;
    std::vector<std::vector<char>> stack;
    std::vector<char> done;
    int nthreads = 0, alive = 0, cur = 0, block = 0, block_y = 0, block_z = 0, grid_x = 1, grid_y = 1, grid_z = 1;
    long progress = 0, barriers = 0;
    Rendezvous block_bar, warp_bar[MAX_THREADS / 32];
    double xb[MAX_THREADS][2];
    int pred[MAX_THREADS];
    std::function<void()> body;
    alignas(16) unsigned char dyn_smem[232 * 1024];
};
static State g;

static inline Uint3 thread_idx() { return Uint3{(unsigned)g.cur, 0, 0}; }
static inline Uint3 block_idx() { return Uint3{(unsigned)g.block, (unsigned)g.block_y, (unsigned)g.block_z}; }
static inline Uint3 block_dim() { return Uint3{(unsigned)g.nthreads, 1, 1}; }
static inline Uint3 grid_dim() { return Uint3{(unsigned)g.grid_x, (unsigned)g.grid_y, (unsigned)g.grid_z}; }

static void trampoline(int tid);
#ifdef CTA_EMUL_FAST_SWITCH
static inline void switch_to_main(int me) { cta_emul_switch(&g.ctx[me].sp, g.main_ctx.sp); }
static inline void switch_to_thread(int t) { cta_emul_switch(&g.main_ctx.sp, g.ctx[t].sp); }
static void fiber_entry() { trampoline(g.cur); }          // entered by the `ret` of the first switch; never returns
static void make_thread(int t)
{
    // initial frame: six callee-saved registers (zero) below the entry address; the entry sees a stack that
    // is 8 modulo 16 as after a call
    uintptr_t top = ((uintptr_t)g.stack[t].data() + g.stack[t].size()) & ~(uintptr_t)15;
    void** sp = (void**)(top - 16);                       // slot of the return address: 0 modulo 16
    sp[0] = (void*)fiber_entry;
    sp[1] = nullptr;
    for (int q = 1; q <= 6; ++q) sp[-q] = nullptr;
    g.ctx[t].sp = (void*)(sp - 6);
}
#else
static inline void switch_to_main(int me) { swapcontext(&g.ctx[me].uc, &g.main_ctx.uc); }
static inline void switch_to_thread(int t) { swapcontext(&g.main_ctx.uc, &g.ctx[t].uc); }
static void make_thread(int t)
{
    getcontext(&g.ctx[t].uc);
    g.ctx[t].uc.uc_stack.ss_sp = g.stack[t].data();
    g.ctx[t].uc.uc_stack.ss_size = g.stack[t].size();
    g.ctx[t].uc.uc_link = &g.main_ctx.uc;
    makecontext(&g.ctx[t].uc, (void (*)())trampoline, 1, t);
}
#endif
static void yield_() { const int me = g.cur; switch_to_main(me); g.cur = me; }

// all `expected` participants arrive, then all leave
static void rendezvous(Rendezvous& r, int expected)
{
    const unsigned my = r.gen;
    if (++r.arrived >= expected) { r.arrived = 0; ++r.gen; ++g.progress; return; }
    while (r.gen == my) yield_();
}

static inline void syncthreads() { ++g.barriers; rendezvous(g.block_bar, g.alive); }
static inline void syncwarp() { rendezvous(g.warp_bar[g.cur >> 5], 32); }
static inline unsigned ballot(bool p)
{
    const int w = g.cur >> 5;
    g.pred[g.cur] = p ? 1 : 0;
    rendezvous(g.warp_bar[w], 32);
    unsigned m = 0;
    for (int q = 0; q < 32; ++q) m |= (unsigned)g.pred[32 * w + q] << q;
    rendezvous(g.warp_bar[w], 32);
    return m;
}
static inline double shfl(double v, int src_lane)
{
    const int w = g.cur >> 5;
    g.xb[g.cur][0] = v;
    rendezvous(g.warp_bar[w], 32);
    const double r = g.xb[32 * w + (src_lane & 31)][0];
    rendezvous(g.warp_bar[w], 32);
    return r;
}

static void trampoline(int tid)
{
    g.cur = tid;
    g.body();
    g.done[tid] = 1;
    --g.alive;
    ++g.progress;
    // a thread that returns no longer takes part in block barriers: release one that now is complete
    if (g.block_bar.arrived > 0 && g.block_bar.arrived >= g.alive) { g.block_bar.arrived = 0; ++g.block_bar.gen; }
    switch_to_main(tid);
    abort();                                              // a finished thread is never resumed
}

// run body() on every thread of a CTA of `nthreads` threads (blockIdx = (block, block_y)) to completion
static void run(int nthreads, int block, std::function<void()> body, int block_y = 0, int grid_x = 1, int grid_y = 1,
                int block_z = 0, int grid_z = 1)
{
    g.block_y = block_y; g.grid_x = grid_x; g.grid_y = grid_y; g.block_z = block_z; g.grid_z = grid_z;
    if (nthreads > MAX_THREADS || nthreads % 32) { fprintf(stderr, "cta: bad block size %d\n", nthreads); abort(); }
    g.body = body; g.nthreads = g.alive = nthreads; g.block = block;
    g.block_bar = Rendezvous();
    for (auto& w : g.warp_bar) w = Rendezvous();
    g.ctx.resize(nthreads); g.stack.resize(nthreads); g.done.assign(nthreads, 0);
    for (int t = 0; t < nthreads; ++t) {
        if (g.stack[t].empty()) g.stack[t].resize(96 * 1024);
        make_thread(t);
    }
    static const char* seed_env = getenv("CTA_SEED");
    static unsigned long long rng = seed_env ? 0x9E3779B97F4A7C15ull * (unsigned long long)(atoll(seed_env) + 1) : 0;
    std::vector<int> order(nthreads);
    for (int q = 0; q < nthreads; ++q) order[q] = q;
    for (;;) {
        bool any = false;
        const long before = g.progress;
        if (seed_env)
            for (int q = nthreads - 1; q > 0; --q) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const int t = (int)((rng >> 33) % (unsigned)(q + 1));
                std::swap(order[q], order[t]);
            }
        for (int q = 0; q < nthreads; ++q) {
            const int t = order[q];
            if (!g.done[t]) { any = true; g.cur = t; switch_to_thread(t); }
        }
        if (!any) break;
        if (g.progress == before) { fprintf(stderr, "cta: deadlock (a barrier that not every live thread reaches)\n"); abort(); }
    }
}
// a whole grid, CTA after CTA (kernels without inter-CTA synchronisation)
static void run_grid(int gx, int gy, int gz, int nthreads, std::function<void()> body)
{
    for (int z = 0; z < gz; ++z)
        for (int y = 0; y < gy; ++y)
            for (int x = 0; x < gx; ++x) run(nthreads, x, body, y, gx, gy, z, gz);
}
}  // namespace cta
