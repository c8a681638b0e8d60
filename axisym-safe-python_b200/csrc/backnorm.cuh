// k_backnorm: per-reflector back-transformation + B-normalisation of the shared-memory eigen path (the
// instance for 128 < 2N <= 166 and behind backnorm_wy = 0; 2N <= 128 normally runs the blocked DMMA kernel of
// eig_backx.cu).  A header of its own so that tests/host/sweep_emul.cpp can compile this very source as host
// code and run it on an emulated thread block (tests/host/cta_emul.h) without a GPU.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
// -----------------------------------------------------------------------------------------
// k_backnorm: psi <- normalise( D * Q * X ), Q = H_0 H_1 ... H_{n-3} (Householder reflectors left
// in the work matrix by k_reduce).  CTA = one frequency x 64 modes, 256 threads = 64 modes x 4
// row groups (rows i = g mod 4); the 64-column block of X lives in shared memory.
#define BN_COLS 64
__global__ void __launch_bounds__(256, 1)
k_backnorm(int N, int hb, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
           const cplx* __restrict__ eig_all, cplx* __restrict__ psi_all, AxsWork wk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    const int n = 2 * N;
    const int f = blockIdx.y, tid = threadIdx.x;
    const int col = tid & (BN_COLS - 1), grp = tid >> 6;
    const int mode = blockIdx.x * BN_COLS + col;
    const bool live = mode < n;
    cplx* Xs = (cplx*)smem_raw;                       // [n][BN_COLS]
    cplx* vbuf = Xs + (size_t)n * BN_COLS;            // [3][n]  (triple buffered: one barrier per reflector)
    cplx* red = vbuf + 3 * n;                         // [2][4][BN_COLS]
    const cplx* A = wk.A + (long long)f * n * n;
    const cplx* tau = wk.tau + (long long)f * n;
    const double* dsc = wk.scale + (long long)f * n;
    cplx* psi = psi_all + (long long)f * n * n;

    for (int idx = tid; idx < n * BN_COLS; idx += 256) {
        const int i = idx / BN_COLS, c = idx - i * BN_COLS;
        const int m = blockIdx.x * BN_COLS + c;
        Xs[idx] = (m < n) ? psi[(long long)i * n + m] : mk(0, 0);
    }
    if (n >= 3) {
        const int kk = n - 3;
        for (int i = kk + 2 + tid; i < n; i += 256) vbuf[(kk % 3) * n + i] = A[(long long)kk * n + i];
    }
    __syncthreads();
    for (int kk = n - 3; kk >= 0; --kk) {
        const cplx* vv = vbuf + (size_t)(kk % 3) * n;
        // prefetch the next reflector (column kk-1, rows kk+1..n-1) into registers
        cplx pre = mk(0, 0);
        const int pi = kk + 1 + tid;
        if (kk > 0 && pi < n) pre = A[(long long)(kk - 1) * n + pi];
        const cplx t = tau[kk];
        const bool skip = (t.x == 0.0 && t.y == 0.0);
        cplx part = mk(0, 0);
        if (!skip) {
            int i0 = kk + 1 + ((grp - (kk + 1)) & 3);          // first row >= kk+1 with i mod 4 == grp
            if (i0 == kk + 1) { part = Xs[i0 * BN_COLS + col]; i0 += 4; }
            cplx p0 = mk(0, 0), p1 = mk(0, 0);
            int i = i0;
            for (; i + 4 < n; i += 8) {
                p0 = cfmac(Xs[i * BN_COLS + col], vv[i], p0);
                p1 = cfmac(Xs[(i + 4) * BN_COLS + col], vv[i + 4], p1);
            }
            if (i < n) p0 = cfmac(Xs[i * BN_COLS + col], vv[i], p0);
            part = cadd(part, cadd(p0, p1));
        }
        cplx* rd = red + (size_t)(kk & 1) * 4 * BN_COLS;
        rd[grp * BN_COLS + col] = part;
        if (kk > 0 && pi < n) vbuf[(size_t)((kk - 1) % 3) * n + pi] = pre;
        __syncthreads();
        if (!skip) {
            const cplx dot = cadd(cadd(rd[col], rd[BN_COLS + col]), cadd(rd[2 * BN_COLS + col], rd[3 * BN_COLS + col]));
            const cplx td = cmul(t, dot);
            int i0 = kk + 1 + ((grp - (kk + 1)) & 3);
            if (i0 == kk + 1) { Xs[i0 * BN_COLS + col] = csub(Xs[i0 * BN_COLS + col], td); i0 += 4; }
#pragma unroll 4
            for (int i = i0; i < n; i += 4) Xs[i * BN_COLS + col] = cfms(td, vv[i], Xs[i * BN_COLS + col]);
        }
    }
    __syncthreads();
    // v = D y; B-normalise with u = v[N:]:  s = 2 k u^T K6 u + u^T C u
    const cplx lam = eig_all[(long long)f * n + (live ? mode : n - 1)];
    for (int i = grp; i < n; i += 4) Xs[i * BN_COLS + col] = cscale(Xs[i * BN_COLS + col], dsc[i]);
    __syncthreads();
    const int W = 2 * hb + 1;
    cplx s6 = mk(0, 0), sc = mk(0, 0);
    for (int r = grp; r < N; r += 4) {
        const cplx ur = Xs[(N + r) * BN_COLS + col];
        s6 = cfma(cmul(ur, ur), K6d[r], s6);
        cplx cu0 = mk(0, 0), cu1 = mk(0, 0);
        const int c0 = max(0, r - hb), c1 = min(N - 1, r + hb);
        const cplx* crow = Cb + (long long)r * W - r + hb;
        int c = c0;
        for (; c + 1 <= c1; c += 2) {
            cu0 = cfma(crow[c], Xs[(N + c) * BN_COLS + col], cu0);
            cu1 = cfma(crow[c + 1], Xs[(N + c + 1) * BN_COLS + col], cu1);
        }
        if (c <= c1) cu0 = cfma(crow[c], Xs[(N + c) * BN_COLS + col], cu0);
        sc = cfma(ur, cadd(cu0, cu1), sc);
    }
    red[grp * BN_COLS + col] = cadd(cscale(cmul(lam, s6), 2.0), sc);
    __syncthreads();
    const cplx nrm = cadd(cadd(red[col], red[BN_COLS + col]), cadd(red[2 * BN_COLS + col], red[3 * BN_COLS + col]));
    const cplx inv = cdiv(mk(1.0, 0.0), csqrt_(nrm));
    if (live) {
        for (int r = grp; r < N; r += 4) {
            const cplx u = cmul(Xs[(N + r) * BN_COLS + col], inv);
            psi[(long long)(N + r) * n + mode] = u;
            psi[(long long)r * n + mode] = cmul(lam, u);
        }
    }
}

static size_t backnorm_smem(int n) {
    return ((size_t)n * BN_COLS + 3 * n + 2 * 4 * BN_COLS) * sizeof(cplx);
}
