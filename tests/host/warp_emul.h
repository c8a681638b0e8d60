// TEST INFRASTRUCTURE: a 32-lane warp emulated on the host with cooperative fibers (ucontext), so that
// warp-synchronous device code written against shuffle / ballot primitives (csrc/aed.cuh, the
// register-resident AED window) runs unchanged - the very same source - without a GPU.
//
// Every lane is a fiber with its own stack; a collective (shuffle, ballot, warp barrier) is a two-phase
// rendezvous: all lanes deposit their operand, then all lanes read.  Lanes are scheduled round-robin by
// run(); a lane that waits in a rendezvous yields to the scheduler.  The emulator checks what the
// hardware only assumes: every lane must reach the SAME collective (call sites are tagged with their
// line number) - divergent control flow around a collective aborts with a message instead of hanging.
// Between two collectives the lanes run one after the other, so a missing barrier (one lane writing what
// another still has to read) shows as a result that depends on the ORDER of the lanes: WEMU_SEED=<n> in the
// environment makes the scheduler visit the lanes in a new pseudo-random order every round; race-free code
// gives bit-identical results for every seed (tests/test_host.py).
#pragma once
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

namespace wemu {
constexpr int W = 32;

struct State {
    ucontext_t main_ctx, ctx[W];
    std::vector<char> stack[W];
    bool done[W];
    int arrived;
    unsigned gen;
    long progress;                      // bumped whenever a rendezvous completes or a lane finishes
    double xb[W][2];
    int pred[W], tag[W];
    long collectives;
    std::function<void(int)> body;
};
static State g;

static void trampoline(int lane)
{
    g.body(lane);
    g.done[lane] = true;
    ++g.progress;
    swapcontext(&g.ctx[lane], &g.main_ctx);
}

static void rendezvous(int lane)
{
    const unsigned my = g.gen;
    if (++g.arrived == W) { g.arrived = 0; ++g.gen; ++g.progress; return; }
    while (g.gen == my) swapcontext(&g.ctx[lane], &g.main_ctx);
}

static void check_tags(int lane, int tag)
{
    for (int q = 0; q < W; ++q)
        if (g.tag[q] != tag) {
            fprintf(stderr, "wemu: lanes %d and %d are in different collectives (lines %d, %d)\n", lane, q, tag, g.tag[q]);
            abort();
        }
}

// value of lane `src` (like __shfl_sync(0xffffffff, v, src)); src must be a valid lane
static inline void shfl2(int lane, double& x, double& y, int src, int tag)
{
    if (src < 0 || src >= W) { fprintf(stderr, "wemu: lane %d shuffles from lane %d (line %d)\n", lane, src, tag); abort(); }
    g.xb[lane][0] = x; g.xb[lane][1] = y; g.tag[lane] = tag;
    rendezvous(lane);
    check_tags(lane, tag);
    x = g.xb[src][0]; y = g.xb[src][1];
    rendezvous(lane);
    if (lane == 0) ++g.collectives;
}

static inline unsigned ballot(int lane, bool p, int tag)
{
    g.pred[lane] = p ? 1 : 0; g.tag[lane] = tag;
    rendezvous(lane);
    check_tags(lane, tag);
    unsigned m = 0;
    for (int q = 0; q < W; ++q) m |= (unsigned)g.pred[q] << q;
    rendezvous(lane);
    if (lane == 0) ++g.collectives;
    return m;
}

static inline void syncwarp(int lane, int tag)
{
    g.tag[lane] = tag;
    rendezvous(lane);
    check_tags(lane, tag);
    rendezvous(lane);
}

// run body(lane) on 32 lanes to completion
static void run(std::function<void(int)> body)
{
    g.body = body;
    g.arrived = 0;
    g.progress = 0;
    for (int lane = 0; lane < W; ++lane) {
        g.done[lane] = false;
        if (g.stack[lane].empty()) g.stack[lane].resize(512 * 1024);
        getcontext(&g.ctx[lane]);
        g.ctx[lane].uc_stack.ss_sp = g.stack[lane].data();
        g.ctx[lane].uc_stack.ss_size = g.stack[lane].size();
        g.ctx[lane].uc_link = &g.main_ctx;
        makecontext(&g.ctx[lane], (void (*)())trampoline, 1, lane);
    }
    static const char* seed_env = getenv("WEMU_SEED");
    static unsigned long long rng = seed_env ? 0x9E3779B97F4A7C15ull * (unsigned long long)(atoll(seed_env) + 1) : 0;
    int order[W];
    for (int q = 0; q < W; ++q) order[q] = q;
    for (;;) {
        bool any = false;
        const long before = g.progress;
        if (seed_env) {                                   // Fisher-Yates with a 64-bit LCG
            for (int q = W - 1; q > 0; --q) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const int t = (int)((rng >> 33) % (unsigned)(q + 1));
                const int tmp = order[q]; order[q] = order[t]; order[t] = tmp;
            }
        }
        for (int q = 0; q < W; ++q) {
            const int lane = order[q];
            if (!g.done[lane]) { any = true; swapcontext(&g.main_ctx, &g.ctx[lane]); }
        }
        if (!any) break;
        if (g.progress == before) {
            // a whole round without a completed rendezvous: some lanes returned while others wait
            fprintf(stderr, "wemu: deadlock - %d lane(s) wait in a collective that the others never reach\n", g.arrived);
            abort();
        }
    }
}
}  // namespace wemu
