// TEST INFRASTRUCTURE: lane-collective model of mma.sync.aligned.m8n8k4.row.col.f64 (SASS DMMA.8x8x4) for the
// emulated warp: D (8 x 8) += A (8 x 4) B (4 x 8) with the fragment layout of the PTX ISA -
//   lane l holds  a = A[l >> 2][l & 3],  b = B[l & 3][l >> 2],  d0 / d1 = D[l >> 2][2 (l & 3) + 0 / 1].
#pragma once
#include "../cta_emul.h"
#define AXS_HOST_DMMA 1
static inline void dmma(double& d0, double& d1, double a, double b)
{
    const int w = cta::g.cur >> 5, l = cta::g.cur & 31, base = 32 * w;
    cta::g.xb[cta::g.cur][0] = a;
    cta::g.xb[cta::g.cur][1] = b;
    cta::rendezvous(cta::g.warp_bar[w], 32);
    const int row = l >> 2, c0 = 2 * (l & 3);
    double s0 = d0, s1 = d1;
    for (int k = 0; k < 4; ++k) {
        const double aik = cta::g.xb[base + 4 * row + k][0];
        s0 = fma(aik, cta::g.xb[base + 4 * c0 + k][1], s0);
        s1 = fma(aik, cta::g.xb[base + 4 * (c0 + 1) + k][1], s1);
    }
    cta::rendezvous(cta::g.warp_bar[w], 32);
    d0 = s0; d1 = s1;
}
static inline void dmma8(double& d0, double& d1, double a, double b) { dmma(d0, d1, a, b); }
static inline void br_dmma(double& d0, double& d1, double a, double b) { dmma(d0, d1, a, b); }
