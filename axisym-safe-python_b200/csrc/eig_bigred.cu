// Kernel 2 of the SAFE sweep, large instance (166 < 2N <= 8192): BLOCKED Householder reduction to
// Hessenberg form with compact-WY panels and FP64 tensor-core (DMMA) trailing updates - the zgehrd
// half of scipy.linalg.eig / LAPACK zgeev (solver.py:136) - using the WHOLE GPU for every matrix.
//
// k_reduce_big (eig_big.cu) keeps one CTA per frequency: level-2, 16 n^3 B of HBM traffic per matrix
// and one SM per matrix however few matrices there are.  Here a batch of nf matrices moves through
// the reduction in lockstep, one kernel launch per step, grid = (work tiles, nf):
//
//   per column c = k0 + j of a panel of nb columns (LAPACK zlahr2):
//     k_br_col    one CTA per matrix: finish Y(:, j-1), T(:, j-1) of the previous column; bring column c
//                 up to date with the panel's reflectors  b <- (I - V T^H V^H)(a_c - Y V(c,:)^H);
//                 generate reflector v_j (zlarfg)
//     k_br_gemv   row blocks x column groups per matrix: A(k0+1:n, c+1:n) v_j, the only level-2 part:
//                 the rows below the panel top are read once per column, 16 (n-k0)(n-c) B, HBM bound
//   per panel (zgehrd's level-3 part), all on DMMA tiles (mma.sync.m8n8k4.f64):
//     k_br_gemm<YTOP>   Y(0:k0, :) = A(0:k0, k0+1:n) V        (then x T and the panel's top rows: k_br_top)
//     k_br_gemm<RIGHT>  A(:, t0:n)      -= Y V(t0:n, :)^H
//     k_br_gemm<LEFTW>  W = V^H A(k0+1:n, t0:n)               (then W <- T^H W: k_br_tw)
//     k_br_gemm<LEFTA>  A(k0+1:n, t0:n) -= V W
//   k_br_pack     packed copies of H for the QR and inverse-iteration kernels, max |H|
//
// HBM traffic per matrix: 16/3 n^3 B for the products A v plus ~2 n^3 / nb x 16 B for the updates
// (nb = 32) instead of 16 n^3 B; the 10/3 n^3 complex flops move from DFMA loops to DMMA tiles except
// for the 2/3 n^3 of the products A v.  All reductions have a fixed summation order, so results are
// bit-reproducible (chunked == one-shot, sharded == single GPU).  The reflectors stay in the work matrix
// below the subdiagonal with tau in wk.tau, exactly as k_reduce_big leaves them (the back-transformation
// kernels do not care which reduction ran).
#include "common.cuh"
#include <stdlib.h>

#include "bigred_kernels.cuh"               // k_br_col, k_br_gemv, k_br_gemm<...>, k_br_top, k_br_tw, k_bx_tw, k_bx_scale, k_br_pack*

extern "C" int axs_launch_balance_big(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                      const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                      const cplx* S_in, AxsWorkBig wk, cudaStream_t st);

extern "C" int axs_launch_reduce_blocked(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                         const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                         const cplx* S_in, AxsWorkBig wk, cudaStream_t st)
{
    const int n = 2 * N;
    // S(w), |a|^2 tables and the zgebal scaling: phases A + B of k_reduce_big (one CTA per matrix)
    int rc = axs_launch_balance_big(N, hb, K0s, Cb, Mb, K1c, K6d, omega, nf, S_in, wk, st);
    if (rc) return rc;
    int nb = axs_tuning()->big_panel;
    if (nb < 4) nb = 4;
    if (nb > BR_NB_MAX) nb = BR_NB_MAX;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const size_t col_smem = (size_t)n * sizeof(cplx);
    AXS_SMEM_ONCE(col_smem, k_br_col);
    // column groups of the products A v: ALWAYS 16 (each group = ceil(columns / 16) consecutive columns), so
    // that the summation order - and with it every bit of the result - does not depend on the batch size
    // (chunked == one-shot, sharded == single GPU); CG is also a layout parameter of the scratch region
    // (Y, W, partial sums, T must fit the yw rows of the workspace).
    const int t_rows = (BR_NB_MAX * BR_NB_MAX + n - 1) / n;
    const int CG = 16;
    if (2 * BR_NB_MAX + CG + t_rows > AXS_BIG_YW) return -1;
    for (int k0 = 0; k0 < n - 2; k0 += nb) {
        const int nbc = (nb < n - 2 - k0) ? nb : n - 2 - k0;
        const int rbn = (n - k0 - 1 + BR_GV_NT - 1) / BR_GV_NT;
        for (int j = 0; j < nbc; ++j) {
            k_br_col<<<nf, BR_COL_NT, col_smem, st>>>(n, k0, j, CG, 0, wk);
            k_br_gemv<<<dim3(rbn * CG, nf), BR_GV_NT, 0, st>>>(n, k0, k0 + j, CG, wk);
        }
        k_br_col<<<nf, BR_COL_NT, col_smem, st>>>(n, k0, nbc, CG, 1, wk);
        const int t0 = k0 + nbc, ntrail = n - t0;
        k_br_gemm<BR_YTOP><<<dim3((k0 + 1 + BR_TM - 1) / BR_TM, 1, nf), 256, 0, st>>>(n, k0, nbc, CG, wk, nullptr);
        k_br_top<<<dim3((k0 + 1 + 255) / 256, nf), 256, 0, st>>>(n, k0, nbc, CG, nb == BR_NB_MAX ? k0 / nb : -1, wk);   // (Tall has room for panels of 32)
        if (ntrail > 0) {
            const int ntn = (ntrail + BR_TN - 1) / BR_TN;
            k_br_gemm<BR_RIGHT><<<dim3((n + BR_TM - 1) / BR_TM, ntn, nf), 256, 0, st>>>(n, k0, nbc, CG, wk, nullptr);
            k_br_gemm<BR_LEFTW><<<dim3((ntrail + BR_TM - 1) / BR_TM, 1, nf), 256, 0, st>>>(n, k0, nbc, CG, wk, nullptr);
            k_br_tw<<<dim3((ntrail + 255) / 256, nf), 256, 0, st>>>(n, k0, nbc, CG, wk);
            k_br_gemm<BR_LEFTA><<<dim3((n - k0 - 1 + BR_TM - 1) / BR_TM, ntn, nf), 256, 0, st>>>(n, k0, nbc, CG, wk, nullptr);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    // packed copies.  The running maximum of |H| goes through a small integer buffer: the per-CTA scratch of
    // k_modes_col (wk.msc) is dead until that kernel runs.
    unsigned long long* hmax = (unsigned long long*)wk.msc;
    AXS_CUDA_OK(cudaMemsetAsync(hmax, 0, (size_t)nf * sizeof(unsigned long long), st));
    int gx = (2 * sms + nf - 1) / nf;
    if (gx < 1) gx = 1;
    if (gx > n) gx = n;
    k_br_pack<<<dim3(gx, nf), 256, 0, st>>>(n, wk, hmax);
    k_br_pack_fin<<<dim3((n + 255) / 256 < 64 ? (n + 255) / 256 : 64, nf), 256, 0, st>>>(n, wk, hmax);
    return (int)cudaGetLastError();
}

// Blocked back-transformation X <- D Q X, Q = product of the panels' (I - V T V^H) in reduction order, i.e.
// the panels are applied to X last to first.  Needs the T factors kept by the blocked reduction of the same
// workspace (k_br_top -> Tall).  psi holds the Hessenberg-basis vectors from k_modes_col.
extern "C" int axs_launch_backx_blocked(int n, int nf, cplx* psi, AxsWorkBig wk, cudaStream_t st)
{
    int nb = axs_tuning()->big_panel;
    if (nb < 4) nb = 4;
    if (nb > BR_NB_MAX) nb = BR_NB_MAX;
    const int CG = 16;
    int last = 0;
    while (last + nb < n - 2) last += nb;
    for (int k0 = last; k0 >= 0; k0 -= nb) {
        const int nbc = (nb < n - 2 - k0) ? nb : n - 2 - k0;
        k_br_gemm<BR_BXW><<<dim3((n + BR_TM - 1) / BR_TM, 1, nf), 256, 0, st>>>(n, k0, nbc, CG, wk, psi);
        k_bx_tw<<<dim3((n + 255) / 256, nf), 256, 0, st>>>(n, nbc, CG, k0 / nb, wk);
        k_br_gemm<BR_BXA><<<dim3((n - k0 - 1 + BR_TM - 1) / BR_TM, (n + BR_TN - 1) / BR_TN, nf), 256, 0, st>>>(n, k0, nbc, CG, wk, psi);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    k_bx_scale<<<dim3(256, nf), 256, 0, st>>>(n, psi, wk);
    return (int)cudaGetLastError();
}
