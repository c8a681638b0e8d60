// On-chip instance of kernel 2 (balancing + Hessenberg reduction) for the shared-memory eigen
// path, 96 <= 2N <= 166.  Replaces the front half of LAPACK zgeev (zgebal 'S' + zgehd2) behind
// scipy.linalg.eig (reference solver.py:136).
//
//   k_balance     S(w) straight from the banded operators, zgebal power-of-2 scaling (one warp
//                 runs LAPACK's Gauss-Seidel loop on a fp32 |a_ij|^2 table), scaled matrix -> HBM.
//                 Small CTAs, 3 per SM: the scaling loop is pure latency, co-residency hides it.
//   k_reduce_oc   Householder reduction with the trailing block ON CHIP: columns jg..n-1 live in
//                 shared memory for the whole reduction, only the first jg columns (which leave the
//                 trailing block first) stay in global memory / L2.  One CTA per SM.  Per reflector
//                 ONE fused pass: rank-2 update of reflector k + the products w = A v', y = A^H v'
//                 of reflector k+1, the per-lane slices of v, w, v' hoisted into registers, four
//                 block barriers per reflector.
// k_reduce (eig_small.cu) keeps the matrix in L2 with two CTAs per SM and is latency bound by L2
// round trips (13.2 ms / 2000 points at 2N = 126); this pair is bound by shared-memory bandwidth
// and the FP64 pipe.
#include "common.cuh"
#include <stdlib.h>
#include <stdio.h>

#include "reduce_oc.cuh"                    // k_balance, k_reduce_oc, balance_smem, reduce_oc_fixed

template <int NWARP, int R, bool PAD, bool PIPE>
static int launch_oc(int n, int nf, AxsWork wk, cudaStream_t st)
{
    const size_t cap = 227 * 1024;
    const size_t fixed = reduce_oc_fixed<NWARP, R>();
    const size_t colb = (size_t)(PAD ? 32 * R : n) * sizeof(cplx);
    int ncs = (int)((cap - fixed) / colb);
    if (ncs > n - 1) ncs = n - 1;                    // column 0 never needs to be on chip
    const int jg = n - ncs;
    const size_t smem = fixed + (size_t)ncs * colb;
    AXS_SMEM_ONCE(smem, k_reduce_oc<NWARP, R, PAD, PIPE>);
    k_reduce_oc<NWARP, R, PAD, PIPE><<<nf, NWARP * 32, smem, st>>>(n, jg, wk);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_balance(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                  const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                  const cplx* S_in, AxsWork wk, cudaStream_t st)
{
    const int n = 2 * N;
    size_t smem = balance_smem(n);
    AXS_SMEM_ONCE(smem, k_balance);
    k_balance<<<nf, BAL_THREADS, smem, st>>>(N, hb, K0s, Cb, Mb, K1c, K6d, omega, S_in, wk);
    return (int)cudaGetLastError();
}

// returns -1 if this instance does not apply (caller falls back to k_reduce)
extern "C" int axs_launch_reduce_oc(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                    const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                    const cplx* S_in, AxsWork wk, cudaStream_t st)
{
    const int n = 2 * N;
    const int nw = axs_tuning()->reduce_oc;          // 0: L2-resident k_reduce; 8 / 12: warps per CTA
    if (nw == 0 || n < 96 || n > AXS_MAX_SMALL_N2) return -1;
    size_t smem = balance_smem(n);
    AXS_SMEM_ONCE(smem, k_balance);
    k_balance<<<nf, BAL_THREADS, smem, st>>>(N, hb, K0s, Cb, Mb, K1c, K6d, omega, S_in, wk);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (n <= 128) {
        if (nw == 8) return launch_oc<8, 4, true, false>(n, nf, wk, st);
        return launch_oc<12, 4, true, false>(n, nf, wk, st);
    }
    return launch_oc<8, 6, false, false>(n, nf, wk, st);
}
