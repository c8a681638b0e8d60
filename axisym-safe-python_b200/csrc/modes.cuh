// Inverse iteration on the Hessenberg matrix (eigenvectors of the shared-memory eigen path).  A header of
// its own so that tests/host/modes_emul.cpp can compile this very source as host code and run it on an
// emulated thread block (tests/host/cta_emul.h) without a GPU.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
// =========================================================================================
// k_modes
// =========================================================================================
// Inverse iteration on the Hessenberg matrix, one solve (H - k I) x = 1 per mode, done as a
// bottom-up pivoted LU with column operations fused with the back substitution (state per mode:
// one multiplier, one pivot flag and one solution entry per index).
// Mapping: a warp serves 8 modes x 4 lane groups; lane group g owns the rows k with
// (n-1-k) mod 4 == g and the four groups run a skewed software pipeline: the steps of row k at
// column j only need row j+1 to be final, which is tracked by a per-warp progress counter.
// This spreads the n^2/2 sequential steps of a mode over 4 lanes and lets 4 warps (one per SM
// sub-partition) share the shared-memory budget that one 32-mode warp used before.
// grid = (ceil(n / (8*nw)), nf), block = 32*nw.
#include <cuda_pipeline.h>
#define MD_G 4

__global__ void __launch_bounds__(128, 1)
k_modes(int N, int hb, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
        const cplx* __restrict__ eig_all, cplx* __restrict__ psi_all, AxsWork wk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    const int n = 2 * N;
    const int f = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int M = 8 * nw;                                 // modes per CTA
    const int g = lane >> 3, mi = lane & 7, mcol = w * 8 + mi;
    const int mode = blockIdx.x * M + mcol;
    const bool live = mode < n;

    cplx* mul = (cplx*)smem_raw;                          // [n][M]
    cplx* z = mul + (size_t)n * M;                        // [n][M]
    cplx* rbuf = z + (size_t)n * M;                       // [nw][MD_G][2][n]
    unsigned char* swp = (unsigned char*)(rbuf + (size_t)nw * MD_G * 2 * n);   // [n][M]
    cplx* myrow = rbuf + ((size_t)(w * MD_G + g) * 2) * n;

    const cplx* Hr = wk.Hr + (long long)f * hrow_size(n);
    const cplx lam = eig_all[(long long)f * n + (live ? mode : n - 1)];

    double hmax = 0.0;
    for (int i = lane; i < hrow_size(n); i += 32) hmax = fmax(hmax, cabs1(Hr[i]));
    for (int o = 16; o > 0; o >>= 1) hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    const double tiny = fmax(EPS_D * hmax, SMALL_D);

    // the 8 lanes of a group stage one row of H (columns first..n-1, stored at absolute column)
    auto stage_row = [&](int k, int buf) {
        const int first = hrow_first(k), off = hrow_off(k, n);
        cplx* dst = myrow + (size_t)buf * n;
        for (int c = first + mi; c < n; c += 8) __pipeline_memcpy_async(dst + c, Hr + off + c - first, sizeof(cplx));
        __pipeline_commit();
    };
    // Row iteration r: group g reduces row k = n-1-g-4r.  Steps at columns j >= k+3 only need rows
    // >= k+4, which are final (previous row iteration): they run as one tight loop for all four
    // groups at once.  The last three steps + the new column operation of row k need row k+1, i.e.
    // the previous group's result of this very iteration: the four tails run in group order.
    auto step = [&](const cplx* row, int k, int j, cplx& carry, cplx& acc, bool diag) {
        cplx a = row[j];
        if (diag && j == k) a = csub(a, lam);
        const bool sw = swp[j * M + mcol] != 0;
        cplx a2 = sw ? carry : a;
        const cplx c2 = sw ? a : carry;
        a2 = cfms(mul[j * M + mcol], c2, a2);
        acc = cfms(c2, z[(j + 1) * M + mcol], acc);
        carry = a2;
    };
    if (n - 1 - g >= 0) stage_row(n - 1 - g, 0);
    int buf = 0;
    for (int r = 0; n - 1 - MD_G * r >= 0; ++r) {
        const int k = n - 1 - g - MD_G * r;
        const bool active = k >= 0;
        __pipeline_wait_prior(0);
        __syncwarp();
        const cplx* row = myrow + (size_t)buf * n;
        if (active && k - MD_G >= 0) stage_row(k - MD_G, buf ^ 1);      // prefetch the next row
        cplx carry = mk(0, 0), acc = mk(1.0, 0.0);
        if (active) {
            carry = row[n - 1];
            if (k == n - 1) carry = csub(carry, lam);
#pragma unroll 4
            for (int j = n - 2; j >= k + 3; --j) step(row, k, j, carry, acc, false);   // j > k: no diagonal shift
        }
        for (int gg = 0; gg < MD_G; ++gg) {
            if (active && g == gg) {
                const int jtop = (k + 2 < n - 2) ? k + 2 : n - 2;
                for (int j = jtop; j >= k; --j) step(row, k, j, carry, acc, true);
                cplx rkk, a2 = mk(0, 0);
                bool sw = false;
                if (k > 0) {
                    const cplx a = row[k - 1];
                    sw = cabs1(a) > cabs1(carry);
                    a2 = sw ? carry : a;
                    rkk = sw ? a : carry;
                } else rkk = carry;
                if (cabs1(rkk) < tiny) rkk = mk(tiny, 0.0);
                const double rinv = 1.0 / cabs2(rkk);
                const cplx inv = mk(rkk.x * rinv, -rkk.y * rinv);
                if (k > 0) { mul[(k - 1) * M + mcol] = cmul(a2, inv); swp[(k - 1) * M + mcol] = sw ? 1 : 0; }
                z[k * M + mcol] = cmul(acc, inv);
            }
            __syncwarp();
        }
        buf ^= 1;
    }
    __syncwarp();
    // x = W z (group 0 lanes), scale to unit max-norm, hand over to k_backnorm through psi
    if (g == 0) {
        cplx xj = z[mcol];
        for (int jj = 0; jj < n - 1; ++jj) {
            const cplx xj1 = cfms(mul[jj * M + mcol], xj, z[(jj + 1) * M + mcol]);
            if (swp[jj * M + mcol]) { z[jj * M + mcol] = xj1; }
            else { z[jj * M + mcol] = xj; xj = xj1; }
        }
        z[(n - 1) * M + mcol] = xj;
        double mx = 0.0;
        for (int i = 0; i < n; ++i) mx = fmax(mx, cabs1(z[i * M + mcol]));
        const double inv = mx > 0.0 ? 1.0 / mx : 1.0;
        for (int i = 0; i < n; ++i) z[i * M + mcol] = cscale(z[i * M + mcol], inv);
    }
    __syncwarp();
    if (live) {
        cplx* X = psi_all + (long long)f * n * n;
        for (int i = g; i < n; i += MD_G) X[(long long)i * n + mode] = z[i * M + mcol];
    }
}

static size_t modes_smem(int n, int nw) {
    const int M = 8 * nw;
    return (size_t)2 * n * M * sizeof(cplx) + (size_t)nw * MD_G * 2 * n * sizeof(cplx) + (size_t)n * M + 64;
}
