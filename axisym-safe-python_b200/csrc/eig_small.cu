// Kernels 2 + 3 of the SAFE sweep, shared-memory ("small", 2N <= 166) instance:
// one CTA per frequency point.  Together they replace scipy.linalg.eig (LAPACK zgeev,
// solver.py:136) and the B-normalisation (solver.py:142).
//
//   k_reduce   S(w) built straight from the banded operators (never round-trips HBM as a
//              separate batch), zgebal-style power-of-2 scaling (Gauss-Seidel sweeps on a
//              fp32 |a_ij|^2 table in shared memory), unblocked Householder reduction to
//              Hessenberg form with the two-sided update fused into 2 passes over the
//              L2-resident work matrix (one read pass for w = A v, y = A^H v, one
//              read-modify-write pass for the rank-2 update).
//   k_hqr      explicit single-shift (Wilkinson) QR on the packed Hessenberg matrix held in
//              shared memory: systolic column wavefront for the left rotations (one thread
//              per column keeps its running element in a register), row-parallel right
//              rotations; standard + Ahues-Tisseur deflation; exceptional shifts.
//   k_modes    eigenvectors by one step of inverse iteration on the Hessenberg matrix, done
//              as a bottom-up pivoted LU (column operations) fused with the back
//              substitution, so the per-mode state is O(n): 32 modes per warp, state in
//              shared memory; then back-transformation through the Householder reflectors,
//              un-balancing, B-normalisation psi = v / sqrt(2k u^T K6 u + u^T C u) and the
//              store in the reference psi_hat layout.
#include "common.cuh"
#include "aed.cuh"
// AXS_AED_BUILD: which early-deflation window routine the QR kernels carry: 0 = shared memory (default, the
// faster one on B200), 1 = registers + shuffles, 2 = both behind the tuning switch hqr_aed_reg.  1 and 2 are
// A/B builds (tests/run_ab_aed2.sh): a kernel that carries both routines runs EITHER of them 1.7 - 3.6 x
// slower than a kernel that carries one (measured), so the product build has exactly one.
#ifndef AXS_AED_BUILD
#define AXS_AED_BUILD 0
#endif
#include <stdlib.h>
#include <stdio.h>

#include "reduce_l2.cuh"                    // k_reduce (L2-resident Hessenberg reduction, 2N < 96) + reduce_smem


#include "hqr_mb.cuh"                       // k_hqr_mb (multi-bulge QR, one CTA per matrix) + hqr_mb_smem

#include "modes.cuh"                        // k_modes (inverse iteration on the Hessenberg matrix) + modes_smem

#include "backnorm.cuh"                     // k_backnorm (per-reflector back-transformation + normalisation) + backnorm_smem


// =========================================================================================
// debug accessor: dense H and scale factors out of the workspace
// =========================================================================================
__global__ void k_unpack_H(int n, AxsWork wk, cplx* __restrict__ Hd)
{
    const int f = blockIdx.y;
    const cplx* Hc = wk.Hc + (long long)f * hcol_size(n);
    cplx* out = Hd + (long long)f * n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        int c = idx / n, r = idx - c * n;
        out[idx] = (r <= c + 1) ? Hc[hcol_off(c) + r] : mk(0, 0);
    }
}

// ------------------------------------------------------------------------------ launchers
extern "C" int axs_launch_reduce_oc(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                    const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                    const cplx* S_in, AxsWork wk, cudaStream_t st);

extern "C" int axs_launch_balance(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                  const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                  const cplx* S_in, AxsWork wk, cudaStream_t st);

extern "C" int axs_launch_reduce(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                 const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                 const cplx* S_in, AxsWork wk, cudaStream_t st)
{
    const int n = 2 * N;
    // 2N >= 96: trailing block on chip (eig_reduce.cu); smaller matrices stay L2-resident, 2 CTAs per SM
    const int rc = axs_launch_reduce_oc(N, hb, K0s, Cb, Mb, K1c, K6d, omega, nf, S_in, wk, st);
    if (rc != -1) return rc;
    // small matrices: the short-chain balancing kernel first (tuning reduce_prebal = 0: the fused one-warp loop)
    int prebal = 0;
    if (axs_tuning()->reduce_prebal) {
        const int rb = axs_launch_balance(N, hb, K0s, Cb, Mb, K1c, K6d, omega, nf, S_in, wk, st);
        if (rb) return rb;
        prebal = 1;
    }
    size_t smem = reduce_smem(n);
    AXS_SMEM_ONCE(smem, k_reduce);
    k_reduce<<<nf, RED_THREADS, smem, st>>>(N, hb, K0s, Cb, Mb, K1c, K6d, omega, S_in, wk, prebal);
    return (int)cudaGetLastError();
}

// launch one k_hqr_mb stage
// AED window of the stage with `per_sm` co-resident CTAs (1 -> stage 1, 2, 3, >= 4 -> the last instances):
// tuning hqr_aed, masked per stage by hqr_aed_stages (bit per_sm - 1)
static int aed_window_of_stage(int per_sm)
{
    int aed = axs_tuning()->hqr_aed;
    if (aed > AED_MAX) aed = AED_MAX;
    const int bit = (per_sm >= 4) ? 3 : per_sm - 1;
    if (aed < 2 || !((axs_tuning()->hqr_aed_stages >> bit) & 1)) return 0;
    return aed;
}

template <int NT, int MINB, int NITEMS, bool OPT>
static int launch_mb(int nf, size_t smem, cudaStream_t st, int n, int bmax, int first, int m_stop,
                     const cplx* Hc, cplx* Hp, cplx* eig, int* info, int* ucur, int* stats)
{
    int aed = aed_window_of_stage(MINB);
    if (aed >= 2) {
        if (axs_tuning()->hqr_aed_reg) aed |= 0x100;
        AXS_SMEM_ONCE(smem, k_hqr_mb<NT, MINB, NITEMS, OPT, true>);
        k_hqr_mb<NT, MINB, NITEMS, OPT, true><<<nf, NT, smem, st>>>(n, bmax, first, m_stop, aed, Hc, Hp, eig, info, ucur, stats);
    } else {
        AXS_SMEM_ONCE(smem, k_hqr_mb<NT, MINB, NITEMS, OPT, false>);
        k_hqr_mb<NT, MINB, NITEMS, OPT, false><<<nf, NT, smem, st>>>(n, bmax, first, m_stop, 0, Hc, Hp, eig, info, ucur, stats);
    }
    return (int)cudaGetLastError();
}

// Largest active window whose packed storage fits `per_sm` times into the shared memory of an SM
// (per CTA 2 KB: static shared memory of k_hqr_mb + the 1 KB the system reserves; 5 KB with the AED scratch).
extern "C" int axs_hqr_stage_capacity(int per_sm)
{
    const size_t slack = (aed_window_of_stage(per_sm) >= 2) ? 5120 : 2048;
    int m = 0;
    while (hqr_mb_smem(m + 1) + slack <= (size_t)(227 * 1024 / per_sm)) ++m;
    return m;
}

// The cascade below stage 1: starts with the instance whose occupancy fits the current window
// (512 threads x 2 CTAs/SM, 384 x 3, 256 x 4, 256 x 5, and 128 x 8 for windows <= 35) and hands the
// shrinking window down.  Hc_first != nullptr: the matrices are the un-parked packed Hessenberg
// matrices of size n (small problems that skip stage 1); otherwise the leading blocks parked in Hp
// (window <= m_cur, top index in ucur).  Also the tail of the large path.
static int hqr_cascade(int n, int nf, int m_cur, const cplx* Hc_first, cplx* Hp, cplx* eig, int* info, int* ucur,
                       int* stats, cudaStream_t st)
{
    const int stages = axs_tuning()->hqr_stages;       // 4 stages measured fastest on B200 (36.3 vs 37.7 ms)
    int bulges = axs_tuning()->hqr_bulges;
    if (bulges < 1 || bulges > MB_MAX) bulges = MB_MAX;
    const int m3 = (stages >= 3) ? axs_hqr_stage_capacity(3) : 0;
    const int m4 = (stages >= 4) ? axs_hqr_stage_capacity(4) : 0;
    const int m5 = (stages >= 5 || Hc_first) ? axs_hqr_stage_capacity(5) : 0;
    int first = Hc_first ? 1 : 0;
    const cplx* src = Hc_first ? Hc_first : Hp;
    if (m_cur <= 35 && bulges * (m_cur + 1) <= 3 * (128 - 32) && stages >= 4)
        return launch_mb<128, 8, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, 0, src, Hp, eig, info, ucur, stats);
    if (m3 == 0 || m_cur > m3) {
        const int stop = (m3 >= 8 && m3 < m_cur) ? m3 : 0;
        const int rc = launch_mb<512, 2, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (rc || !stop) return rc;
        first = 0; src = Hp; m_cur = stop;
    }
    if (m4 == 0 || m_cur > m4) {
        const int stop = (m4 >= 8 && m4 < m_cur) ? m4 : 0;
        const int rc = launch_mb<384, 3, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (rc || !stop) return rc;
        first = 0; src = Hp; m_cur = stop;
    }
    if (!(m5 >= 8 && m_cur <= m5)) {
        const int stop = (m5 >= 8 && m5 < m_cur) ? m5 : 0;
        const int rc = launch_mb<256, 4, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (rc || !stop) return rc;
        first = 0; src = Hp; m_cur = stop;
    }
    return launch_mb<256, 5, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, 0, src, Hp, eig, info, ucur, stats);
}

extern "C" int axs_launch_hqr_stages(int n, int nf, int m_first, cplx* Hp, cplx* eig, int* info, int* ucur,
                                     int* stats, cudaStream_t st)
{
    return hqr_cascade(n, nf, m_first, nullptr, Hp, eig, info, ucur, stats, st);
}

extern "C" int axs_launch_hqr(int n, int nf, cplx* eig, int* info, AxsWork wk, cudaStream_t st)
{
    int bulges = axs_tuning()->hqr_bulges;             // 8 concurrent bulges: fastest measured on B200
    if (bulges < 1 || bulges > MB_MAX) bulges = MB_MAX;
    // stage 1: whole matrix, 512 threads, one CTA per SM; stage 2: window fits twice into shared
    // memory (512 threads, 64 registers, 2 CTAs per SM); stage 3: fits three times (384 threads,
    // 56 registers, 3 CTAs per SM); ...  tuning hqr_stages = 1|2|3 limits the cascade (A/B testing).
    const int stages = axs_tuning()->hqr_stages;
    int m2 = (stages >= 2) ? axs_hqr_stage_capacity(2) : 0;
    if (m2 >= n || m2 < 8) m2 = 0;
    size_t smem = hqr_mb_smem(n);
    // problems that fit into the shared memory of an SM at least twice skip stage 1 and enter the
    // cascade at the instance that fits
    if (m2 == 0 && axs_tuning()->hqr_small_direct)
        return hqr_cascade(n, nf, n, wk.Hc, wk.Hp, eig, info, wk.ucur, wk.stats, st);
    int rc1;
    if (bulges * n <= 3 * 384)
        // leader sub-partition kept free in phase C (the 12 phase-C warps x 3 item slots must hold
        // bulges x n column-rotation items: 2N <= 144)
        rc1 = launch_mb<512, 1, 3, true>(nf, smem, st, n, bulges, 1, m2, wk.Hc, wk.Hp, eig, info, wk.ucur, wk.stats);
    else
        rc1 = launch_mb<512, 1, 3, false>(nf, smem, st, n, bulges, 1, m2, wk.Hc, wk.Hp, eig, info, wk.ucur, wk.stats);
    if (rc1 || m2 == 0) return rc1;
    return axs_launch_hqr_stages(n, nf, m2, wk.Hp, eig, info, wk.ucur, wk.stats, st);
}

extern "C" int axs_launch_backnorm_wy(int N, int hb, const cplx* Cb, const cplx* K6d, int nf,
                                      const cplx* eig, cplx* psi, AxsWork wk, cudaStream_t st);

extern "C" int axs_launch_modes(int N, int hb, const cplx* Cb, const cplx* K6d, int nf,
                                const cplx* eig, cplx* psi, AxsWork wk, cudaStream_t st)
{
    const int n = 2 * N;
    int nw = 4;
    while (nw > 1 && modes_smem(n, nw) > 227 * 1024) --nw;
    size_t smem = modes_smem(n, nw);
    AXS_SMEM_ONCE(smem, k_modes);
    k_modes<<<dim3((n + 8 * nw - 1) / (8 * nw), nf), 32 * nw, smem, st>>>(N, hb, Cb, K6d, eig, psi, wk);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    // 2N <= 128: blocked compact-WY back-transformation on DMMA tiles (eig_backx.cu); above that the
    // per-reflector kernel with a 64-mode slab in shared memory
    {
        const int rc = axs_launch_backnorm_wy(N, hb, Cb, K6d, nf, eig, psi, wk, st);
        if (rc != -1) return rc;
    }
    size_t smem2 = backnorm_smem(n);
    AXS_SMEM_ONCE(smem2, k_backnorm);
    k_backnorm<<<dim3((n + BN_COLS - 1) / BN_COLS, nf), 256, smem2, st>>>(N, hb, Cb, K6d, eig, psi, wk);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_unpack(int n, int nf, AxsWork wk, cplx* Hd, cudaStream_t st)
{
    k_unpack_H<<<dim3(32, nf), 256, 0, st>>>(n, wk, Hd);
    return (int)cudaGetLastError();
}
