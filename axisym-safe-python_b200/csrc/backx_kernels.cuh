// k_wy_tfactor + k_backnorm_wy: blocked compact-WY back-transformation on FP64 tensor-core tiles (see
// eig_backx.cu for the launcher and the overview).  A header of its own so that tests/host/sweep_emul.cpp can
// compile this very source as host code and run it on an emulated thread block (tests/host/cta_emul.h) without
// a GPU: the inline-PTX pieces (DMMA, cp.async) sit behind macros that the host shim replaces by models.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define WY_NB 8
#define WY_V1 9                         // V copy for V^H X: odd
#define WY_V2 12                        // V copy for V W': = 4 mod 8
#define WY_WS 10                        // W row stride: = 2 mod 8
#define WY_MAX_N 128

#ifndef AXS_HOST_DMMA                  /* the host emulation models the instruction itself (tests/host/shim/dmma.h) */
__device__ __forceinline__ void dmma8(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif
// 16-byte cp.async (zero fill when bytes = 0), group commit, wait until at most n groups are pending
#ifndef AXS_CP_ASYNC16
#define AXS_CP_ASYNC16(dst, src, bytes) do {                                                               \
        const unsigned dst_ = (unsigned)__cvta_generic_to_shared(dst);                                     \
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst_), "l"(src), "r"(bytes));  \
    } while (0)
#define AXS_CP_ASYNC_COMMIT() asm volatile("cp.async.commit_group;")
#define AXS_CP_ASYNC_WAIT(n) asm volatile("cp.async.wait_group " #n ";" ::: "memory")
#endif

// reflector entry v_k[row] from the work matrix (unit diagonal at row k+1, zeros above)
__device__ __forceinline__ cplx wy_v(const cplx* __restrict__ A, int n, int k, int row) {
    if (k > n - 3 || row >= n || row <= k) return mk(0, 0);
    if (row == k + 1) return mk(1.0, 0.0);
    return A[(long long)k * n + row];
}

__global__ void __launch_bounds__(32)
k_wy_tfactor(int n, AxsWork wk, cplx* __restrict__ Tg, int nblk)
{
    __shared__ cplx Vw[WY_MAX_N * WY_V1];
    __shared__ cplx Gs[WY_NB * WY_NB];
    __shared__ cplx Ts[WY_NB * WY_NB];
    const int blk = blockIdx.x, f = blockIdx.y, lane = threadIdx.x;
    const int k0 = blk * WY_NB;
    const int n8 = (n + 7) & ~7;
    const int rb = (k0 + 1) & ~7;
    const cplx* A = wk.A + (long long)f * n * n;
    const cplx* tau = wk.tau + (long long)f * n;
    for (int idx = lane; idx < WY_NB * (n8 - rb); idx += 32) {
        const int j = idx / (n8 - rb), row = rb + idx - j * (n8 - rb);
        Vw[row * WY_V1 + j] = wy_v(A, n, k0 + j, row);
    }
    for (int i = lane; i < WY_NB * WY_NB; i += 32) { Ts[i] = mk(0, 0); Gs[i] = mk(0, 0); }
    __syncwarp();
    // G[i][j] = v_i^H v_j for i < j
#pragma unroll
    for (int j = 1; j < WY_NB; ++j) {
        cplx acc[WY_NB - 1];
#pragma unroll
        for (int i = 0; i < WY_NB - 1; ++i) acc[i] = mk(0, 0);
        for (int row = rb + lane; row < n8; row += 32) {
            const cplx vj = Vw[row * WY_V1 + j];
#pragma unroll
            for (int i = 0; i < WY_NB - 1; ++i)
                if (i < j) acc[i] = cfmac(vj, Vw[row * WY_V1 + i], acc[i]);     // vj conj(vi)
        }
#pragma unroll
        for (int i = 0; i < WY_NB - 1; ++i)
            if (i < j) {
                const cplx s = warp_csum(acc[i]);
                if (lane == 0) Gs[i * WY_NB + j] = s;
            }
    }
    __syncwarp();
    // T(j,j) = tau_j;  T(0:j, j) = -tau_j T(0:j, 0:j) G(0:j, j)
    for (int j = 0; j < WY_NB; ++j) {
        const int k = k0 + j;
        const cplx tj = (k <= n - 3) ? tau[k] : mk(0, 0);
        if (lane < j) {
            cplx s = mk(0, 0);
            for (int l = lane; l < j; ++l) s = cfma(Ts[lane * WY_NB + l], Gs[l * WY_NB + j], s);
            Ts[lane * WY_NB + j] = cneg(cmul(tj, s));
        } else if (lane == j) Ts[j * WY_NB + j] = tj;
        __syncwarp();
    }
    cplx* out = Tg + ((long long)f * nblk + blk) * (WY_NB * WY_NB);
    for (int i = lane; i < WY_NB * WY_NB; i += 32) out[i] = Ts[i];
}

template <int WY_MODES>
__global__ void __launch_bounds__(WY_MODES * 4, 128 / WY_MODES)
k_backnorm_wy(int N, int hb, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
              const cplx* __restrict__ eig_all, cplx* __restrict__ psi_all, AxsWork wk,
              const cplx* __restrict__ Tg, int nblk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    constexpr int NT = WY_MODES * 4;                  // one warp per 8 modes
    constexpr int WY_XS = WY_MODES + 1;               // X row stride (cplx): odd
    constexpr int NPRE = (WY_NB * WY_MAX_N) / NT;
    const int n = 2 * N;
    const int n8 = (n + 7) & ~7;
    const int f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;            // mma groupID / threadID_in_group
    const int col = tid & (WY_MODES - 1), grp = tid / WY_MODES;
    const int mode = blockIdx.x * WY_MODES + col;
    const bool live = mode < n;
    cplx* Xs = (cplx*)smem_raw;                       // [n8][WY_XS]
    cplx* V1 = Xs + (size_t)n8 * WY_XS;               // [n8][WY_V1]
    cplx* V2 = V1 + (size_t)n8 * WY_V1;               // [n8][WY_V2]
    cplx* Wsm = V2 + (size_t)n8 * WY_V2;              // [warps][8][WY_WS]
    cplx* Ts = Wsm + (WY_MODES / 8) * WY_NB * WY_WS;               // [64]
    cplx* red = Ts + WY_NB * WY_NB;                   // [4][WY_MODES]
    const cplx* A = wk.A + (long long)f * n * n;
    const double* dsc = wk.scale + (long long)f * n;
    cplx* psi = psi_all + (long long)f * n * n;

    // X block: 16-byte cp.async per element (zero fill outside the matrix), committed in groups of
    // 8 rows from the bottom up: block blk of the loop below only needs the rows >= 8 blk, so it
    // waits until at most blk groups are pending and the load hides behind the first blocks.
    const int ngr = n8 >> 3;
    for (int gi = 0; gi < ngr; ++gi) {
        const int r0 = 8 * (ngr - 1 - gi);
        for (int idx = tid; idx < 8 * WY_MODES; idx += NT) {
            const int i = r0 + idx / WY_MODES, c = idx & (WY_MODES - 1);
            const int m = blockIdx.x * WY_MODES + c;
            const bool in = (m < n && i < n);
            const cplx* src = in ? psi + (long long)i * n + m : psi;
            AXS_CP_ASYNC16(Xs + i * WY_XS + c, src, in ? 16 : 0);
        }
        AXS_CP_ASYNC_COMMIT();
    }
    // register prefetch of a block's V (row fastest: coalesced) and T
    cplx pre[NPRE], preT = mk(0, 0);
    auto prefetch = [&](int blk) {
        const int k0 = blk * WY_NB, rb = (k0 + 1) & ~7, span = n8 - rb;
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const int idx = tid + NT * q;
            const int j = idx / span, row = rb + idx - j * span;
            pre[q] = (idx < WY_NB * span) ? wy_v(A, n, k0 + j, row) : mk(0, 0);
        }
        if (tid < WY_NB * WY_NB) preT = Tg[((long long)f * nblk + blk) * (WY_NB * WY_NB) + tid];
    };
    auto commit = [&](int blk) {
        const int k0 = blk * WY_NB, rb = (k0 + 1) & ~7, span = n8 - rb;
#pragma unroll
        for (int q = 0; q < NPRE; ++q) {
            const int idx = tid + NT * q;
            if (idx < WY_NB * span) {
                const int j = idx / span, row = rb + idx - j * span;
                V1[row * WY_V1 + j] = pre[q];
                V2[row * WY_V2 + j] = pre[q];
            }
        }
        if (tid < WY_NB * WY_NB) Ts[tid] = preT;
    };
    prefetch(nblk - 1);
    cplx* Ww = Wsm + warp * (WY_NB * WY_WS);
    const int c0 = 8 * warp;                          // this warp's modes c0 .. c0+7
    for (int blk = nblk - 1; blk >= 0; --blk) {
        switch (blk) {                                // rows >= 8 blk of X have landed (this thread's copies)
            case 0: AXS_CP_ASYNC_WAIT(0); break;
            case 1: AXS_CP_ASYNC_WAIT(1); break;
            case 2: AXS_CP_ASYNC_WAIT(2); break;
            case 3: AXS_CP_ASYNC_WAIT(3); break;
            case 4: AXS_CP_ASYNC_WAIT(4); break;
            case 5: AXS_CP_ASYNC_WAIT(5); break;
            case 6: AXS_CP_ASYNC_WAIT(6); break;
            case 7: AXS_CP_ASYNC_WAIT(7); break;
            case 8: AXS_CP_ASYNC_WAIT(8); break;
            case 9: AXS_CP_ASYNC_WAIT(9); break;
            case 10: AXS_CP_ASYNC_WAIT(10); break;
            case 11: AXS_CP_ASYNC_WAIT(11); break;
            case 12: AXS_CP_ASYNC_WAIT(12); break;
            case 13: AXS_CP_ASYNC_WAIT(13); break;
            case 14: AXS_CP_ASYNC_WAIT(14); break;
            default: AXS_CP_ASYNC_WAIT(15); break;
        }
        __syncthreads();                              // everyone is done with the previous V / T
        commit(blk);
        __syncthreads();
        if (blk > 0) prefetch(blk - 1);
        const int rb = (blk * WY_NB + 1) & ~7;
        // ---- W = V^H X : M = reflector (g), N = mode, K = rows; rows of a chunk of 8 taken as
        //      {0,2,4,6} then {1,3,5,7} so that both fragment loads are conflict free -----------
        double wre[2] = {0, 0}, wie[2] = {0, 0}, wro[2] = {0, 0}, wio[2] = {0, 0};
        for (int r = rb; r < n8; r += 8) {
            const cplx ve = V1[(r + 2 * t) * WY_V1 + g], vo = V1[(r + 2 * t + 1) * WY_V1 + g];
            const cplx xe = Xs[(r + 2 * t) * WY_XS + c0 + g], xo = Xs[(r + 2 * t + 1) * WY_XS + c0 + g];
            dmma8(wre[0], wre[1], ve.x, xe.x); dmma8(wro[0], wro[1], vo.x, xo.x);
            dmma8(wie[0], wie[1], ve.x, xe.y); dmma8(wio[0], wio[1], vo.x, xo.y);
            dmma8(wre[0], wre[1], ve.y, xe.y); dmma8(wro[0], wro[1], vo.y, xo.y);
            dmma8(wie[0], wie[1], -ve.y, xe.x); dmma8(wio[0], wio[1], -vo.y, xo.x);
        }
        // C fragment: W[refl g][mode 2t, 2t+1]
        Ww[g * WY_WS + 2 * t] = mk(wre[0] + wro[0], wie[0] + wio[0]);
        Ww[g * WY_WS + 2 * t + 1] = mk(wre[1] + wro[1], wie[1] + wio[1]);
        __syncwarp();
        // ---- W' = T W (T upper triangular) ------------------------------------------------------
        cplx w0 = mk(0, 0), w1 = mk(0, 0);
        for (int j = g; j < WY_NB; ++j) {
            const cplx tt = Ts[g * WY_NB + j];
            w0 = cfma(tt, Ww[j * WY_WS + 2 * t], w0);
            w1 = cfma(tt, Ww[j * WY_WS + 2 * t + 1], w1);
        }
        __syncwarp();
        Ww[g * WY_WS + 2 * t] = w0;
        Ww[g * WY_WS + 2 * t + 1] = w1;
        __syncwarp();
        // ---- X -= V W' : M = row (g), N = mode, K = reflector ------------------------------------
        const cplx b0 = Ww[t * WY_WS + g], b1 = Ww[(4 + t) * WY_WS + g];     // W'[k][mode g], k = t, 4+t
        for (int r = rb; r < n8; r += 8) {
            cplx* xp = Xs + (r + g) * WY_XS + c0 + 2 * t;
            const cplx a0 = V2[(r + g) * WY_V2 + t], a1 = V2[(r + g) * WY_V2 + 4 + t];
            cplx x0 = xp[0], x1 = xp[1];
            dmma8(x0.x, x1.x, -a0.x, b0.x); dmma8(x0.y, x1.y, -a0.x, b0.y);
            dmma8(x0.x, x1.x, a0.y, b0.y);  dmma8(x0.y, x1.y, -a0.y, b0.x);
            dmma8(x0.x, x1.x, -a1.x, b1.x); dmma8(x0.y, x1.y, -a1.x, b1.y);
            dmma8(x0.x, x1.x, a1.y, b1.y);  dmma8(x0.y, x1.y, -a1.y, b1.x);
            xp[0] = x0; xp[1] = x1;
        }
    }
    __syncthreads();
    // v = D y; B-normalise with u = v[N:]:  s = 2 k u^T K6 u + u^T C u   (as k_backnorm)
    const cplx lam = eig_all[(long long)f * n + (live ? mode : n - 1)];
    for (int i = grp; i < n; i += 4) Xs[i * WY_XS + col] = cscale(Xs[i * WY_XS + col], dsc[i]);
    __syncthreads();
    const int W = 2 * hb + 1;
    cplx s6 = mk(0, 0), sc = mk(0, 0);
    // The band of C is staged through the (now dead) V area in row chunks: coalesced bulk loads
    // instead of one L2 round trip per band entry.
    cplx* Cs = V1;
    const int rows_per_chunk = (n8 * (WY_V1 + WY_V2) + (WY_MODES / 8) * WY_NB * WY_WS) / W;
    for (int rc0 = 0; rc0 < N; rc0 += rows_per_chunk) {
        const int rc1 = min(N, rc0 + rows_per_chunk);
        __syncthreads();
        for (int idx = tid; idx < (rc1 - rc0) * W; idx += NT) Cs[idx] = Cb[(long long)rc0 * W + idx];
        __syncthreads();
        int r = rc0 + ((grp - rc0) & 3);                         // first row >= rc0 with r mod 4 == grp
        for (; r < rc1; r += 4) {
            const cplx ur = Xs[(N + r) * WY_XS + col];
            s6 = cfma(cmul(ur, ur), K6d[r], s6);
            cplx cu0 = mk(0, 0), cu1 = mk(0, 0);
            const int cl = max(0, r - hb), ch = min(N - 1, r + hb);
            const cplx* crow = Cs + (r - rc0) * W - r + hb;
            int c = cl;
            for (; c + 1 <= ch; c += 2) {
                cu0 = cfma(crow[c], Xs[(N + c) * WY_XS + col], cu0);
                cu1 = cfma(crow[c + 1], Xs[(N + c + 1) * WY_XS + col], cu1);
            }
            if (c <= ch) cu0 = cfma(crow[c], Xs[(N + c) * WY_XS + col], cu0);
            sc = cfma(ur, cadd(cu0, cu1), sc);
        }
    }
    red[grp * WY_MODES + col] = cadd(cscale(cmul(lam, s6), 2.0), sc);
    __syncthreads();
    const cplx nrm = cadd(cadd(red[col], red[WY_MODES + col]), cadd(red[2 * WY_MODES + col], red[3 * WY_MODES + col]));
    const cplx inv = cdiv(mk(1.0, 0.0), csqrt_(nrm));
    if (live) {
        for (int r = grp; r < N; r += 4) {
            const cplx u = cmul(Xs[(N + r) * WY_XS + col], inv);
            psi[(long long)(N + r) * n + mode] = u;
            psi[(long long)r * n + mode] = cmul(lam, u);
        }
    }
}

static size_t backnorm_wy_smem(int n, int modes) {
    const int n8 = (n + 7) & ~7;
    return ((size_t)n8 * (modes + 1 + WY_V1 + WY_V2) + (modes / 8) * WY_NB * WY_WS + WY_NB * WY_NB + 4 * modes) * sizeof(cplx);
}

