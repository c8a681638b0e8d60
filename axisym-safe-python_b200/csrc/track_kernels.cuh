// k_bpsi, k_bpsi_tiled, k_track_mma, k_apply_order: the tracking kernels (see track.cu for the launchers and
// the overview).  A header of its own so that tests/host/track_emul.cpp can compile this very source as host
// code and run it on an emulated thread block (tests/host/cta_emul.h) without a GPU; the one inline-PTX
// instruction (the FP64 tensor-core MMA) is then replaced by a lane-collective model of its fragment layout.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// Y[f][r][j]:  r <  N : K6[r] * psi[N + r][j]
//              r >= N : K6[r-N] * psi[r-N][j] + sum_c C[r-N][c] psi[N + c][j]
__global__ void __launch_bounds__(256)
k_bpsi(int N, int hb, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
       const cplx* __restrict__ psi_all, cplx* __restrict__ Y_all)
{
    const int n = 2 * N, W = 2 * hb + 1;
    const int f = blockIdx.y;
    const cplx* psi = psi_all + (long long)f * n * n;
    cplx* Y = Y_all + (long long)f * n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        const int r = idx / n, j = idx - r * n;
        cplx acc;
        if (r < N) {
            acc = cmul(K6d[r], psi[(long long)(N + r) * n + j]);
        } else {
            const int rr = r - N;
            acc = cmul(K6d[rr], psi[(long long)rr * n + j]);
            const int c0 = max(0, rr - hb), c1 = min(N - 1, rr + hb);
            for (int c = c0; c <= c1; ++c)
                acc = cfma(Cb[(long long)rr * W + c - rr + hb], psi[(long long)(N + c) * n + j], acc);
        }
        Y[idx] = acc;
    }
}

// Tiled form of k_bpsi: CTA = one frequency x 64 modes.  The lower half of the mode block (63 x 64
// at 2N = 126) is staged in shared memory once and reused by every row of the banded C, which is
// streamed through shared memory in row chunks; thread = (mode, row group).
#define BP_COLS 64
__global__ void __launch_bounds__(256, 2)
k_bpsi_tiled(int N, int hb, int crows, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
             const cplx* __restrict__ psi_all, cplx* __restrict__ Y_all)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    const int n = 2 * N, W = 2 * hb + 1;
    const int f = blockIdx.y, tid = threadIdx.x;
    const int col = tid & (BP_COLS - 1), grp = tid >> 6;
    const int mode = blockIdx.x * BP_COLS + col;
    const bool live = mode < n;
    cplx* U = (cplx*)smem_raw;                        // [N][BP_COLS] lower half of the block
    cplx* Cs = U + (size_t)N * BP_COLS;               // [crows][W]
    const cplx* psi = psi_all + (long long)f * n * n;
    cplx* Y = Y_all + (long long)f * n * n;
    for (int idx = tid; idx < N * BP_COLS; idx += 256) {
        const int r = idx / BP_COLS, c = idx - r * BP_COLS;
        const int m = blockIdx.x * BP_COLS + c;
        U[idx] = (m < n) ? psi[(long long)(N + r) * n + m] : mk(0, 0);
    }
    for (int r0 = 0; r0 < N; r0 += crows) {
        const int r1 = min(N, r0 + crows);
        __syncthreads();                              // U staged / previous chunk consumed
        for (int idx = tid; idx < (r1 - r0) * W; idx += 256) Cs[idx] = Cb[(long long)r0 * W + idx];
        __syncthreads();
        int r = r0 + ((grp - r0) & 3);
        for (; r < r1; r += 4) {
            const cplx k6 = K6d[r];
            cplx acc0 = live ? cmul(k6, psi[(long long)r * n + mode]) : mk(0, 0), acc1 = mk(0, 0);
            const int c0 = max(0, r - hb), c1 = min(N - 1, r + hb);
            const cplx* crow = Cs + (r - r0) * W - r + hb;
            int c = c0;
            for (; c + 1 <= c1; c += 2) {
                acc0 = cfma(crow[c], U[c * BP_COLS + col], acc0);
                acc1 = cfma(crow[c + 1], U[(c + 1) * BP_COLS + col], acc1);
            }
            if (c <= c1) acc0 = cfma(crow[c], U[c * BP_COLS + col], acc0);
            if (live) {
                Y[(long long)(N + r) * n + mode] = cadd(acc0, acc1);
                Y[(long long)r * n + mode] = cmul(k6, U[r * BP_COLS + col]);
            }
        }
    }
}

// -----------------------------------------------------------------------------------------
// k_track_mma: the contraction P[a][b] = sum_r psiP[r][a] * Y[r][b] on the FP64 tensor path
// (mma.sync.aligned.m8n8k4.f64, "DMMA"; tcgen05 has no FP64 kind).  A complex product is four
// real MMAs per 8x8x4 tile: Pre += Are Bre + (-Aim) Bim,  Pim += Are Bim + Aim Bre.
// CTA = 32 rows (a) x all columns (b <= 192), 8 warps; warp w owns the n8 column tiles
// w, w+8, w+16 for all four m8 row tiles.  Shared tiles are padded so that the fragment loads
// (4 k-rows x 8 columns per quarter-warp) are bank-conflict free.
#define TM_KC 16
#define TM_SA 34                       // row stride of sA in cplx (32 + 2 -> 544 B = 32 mod 128)
#define TM_SB 194                      // row stride of sB in cplx (192 + 2)
#ifndef AXS_HOST_DMMA                  /* the host emulation defines dmma() itself (tests/host/shim/dmma.h) */
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

__global__ void __launch_bounds__(256, 2)
k_track_mma(int n, const cplx* __restrict__ psi_all, const cplx* __restrict__ Y_all,
            int* __restrict__ arg_all, double* __restrict__ amax_all)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    cplx* sA = (cplx*)smem_raw;                       // [TM_KC][TM_SA]
    cplx* sB = sA + TM_KC * TM_SA;                    // [TM_KC][TM_SB]
    double* s_best = (double*)(sB + TM_KC * TM_SB);   // [8 warps][32 rows]
    int* s_bidx = (int*)(s_best + 8 * 32);
    const int p = blockIdx.y, a0 = blockIdx.x * 32;
    const cplx* psiP = psi_all + (long long)p * n * n;
    const cplx* Y = Y_all + (long long)(p + 1) * n * n;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;          // groupID, thread-in-group
    double best[4];
    int bidx[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) { best[mt] = -1.0; bidx[mt] = 0; }

    // columns b in chunks of 192 (one chunk for the shared-memory eigen path, 2N <= 166)
    for (int b0 = 0; b0 < n; b0 += 192) {
        const int nbc = (n - b0 < 192) ? n - b0 : 192;
        const int ntiles = (nbc + 7) >> 3;
        double cre[4][3][2], cim[4][3][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int t = 0; t < 3; ++t) { cre[mt][t][0] = cre[mt][t][1] = 0.0; cim[mt][t][0] = cim[mt][t][1] = 0.0; }

        for (int r0 = 0; r0 < n; r0 += TM_KC) {
            for (int idx = tid; idx < TM_KC * 32; idx += 256) {
                const int rr = idx >> 5, aa = idx & 31, r = r0 + rr, a = a0 + aa;
                sA[rr * TM_SA + aa] = (r < n && a < n) ? psiP[(long long)r * n + a] : mk(0, 0);
            }
            for (int idx = tid; idx < TM_KC * 192; idx += 256) {
                const int rr = idx / 192, bb = idx - rr * 192, r = r0 + rr;
                sB[rr * TM_SB + bb] = (r < n && bb < nbc) ? Y[(long long)r * n + b0 + bb] : mk(0, 0);
            }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < TM_KC; ks += 4) {
                cplx af[4];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) af[mt] = sA[(ks + tq) * TM_SA + mt * 8 + gq];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int nt = w + 8 * t;
                    if (nt < ntiles) {
                        const cplx bf = sB[(ks + tq) * TM_SB + nt * 8 + gq];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            dmma(cre[mt][t][0], cre[mt][t][1], af[mt].x, bf.x);
                            dmma(cre[mt][t][0], cre[mt][t][1], -af[mt].y, bf.y);
                            dmma(cim[mt][t][0], cim[mt][t][1], af[mt].x, bf.y);
                            dmma(cim[mt][t][0], cim[mt][t][1], af[mt].y, bf.x);
                        }
                    }
                }
            }
            __syncthreads();
        }
        // running row arg-max of |P|^2 (lowest column index wins ties, like np.argmax)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int nt = w + 8 * t;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int bl = nt * 8 + 2 * tq + e;
                    const double m2 = (nt < ntiles && bl < nbc) ? cre[mt][t][e] * cre[mt][t][e] + cim[mt][t][e] * cim[mt][t][e] : -1.0;
                    const int b = b0 + bl;
                    if (m2 > best[mt] || (m2 == best[mt] && b < bidx[mt])) { best[mt] = m2; bidx[mt] = b; }
                }
            }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        double bb = best[mt];
        int bi = bidx[mt];
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {              // across the 4 lanes of a group
            const double ob = __shfl_xor_sync(0xffffffffu, bb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > bb || (ob == bb && oi < bi)) { bb = ob; bi = oi; }
        }
        if (tq == 0) { s_best[w * 32 + mt * 8 + gq] = bb; s_bidx[w * 32 + mt * 8 + gq] = bi; }
    }
    __syncthreads();
    if (tid < 32) {
        double bb = s_best[tid];
        int bi = s_bidx[tid];
        for (int ww = 1; ww < 8; ++ww) {
            const double ob = s_best[ww * 32 + tid];
            const int oi = s_bidx[ww * 32 + tid];
            if (ob > bb || (ob == bb && oi < bi)) { bb = ob; bi = oi; }
        }
        const int a = a0 + tid;
        if (a < n) { arg_all[(long long)p * n + a] = bi; amax_all[(long long)p * n + a] = sqrt(bb); }
    }
}

__global__ void __launch_bounds__(256)
k_apply_order(int n, const int* __restrict__ order, const cplx* __restrict__ k_in,
              const cplx* __restrict__ psi_in, cplx* __restrict__ k_out, cplx* __restrict__ psi_out)
{
    const int f = blockIdx.y;
    const int* ord = order + (long long)f * n;
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < n; c += blockDim.x) k_out[(long long)f * n + c] = k_in[(long long)f * n + ord[c]];
    if (psi_in && psi_out) {
        const cplx* pi = psi_in + (long long)f * n * n;
        cplx* po = psi_out + (long long)f * n * n;
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
            int r = idx / n, c = idx - r * n;
            po[idx] = pi[(long long)r * n + ord[c]];
        }
    }
}

