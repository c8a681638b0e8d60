// Blocked (compact WY) back-transformation of the mode shapes on the FP64 tensor path, 2N <= 128.
// Replaces zunghr/zunmhr + the normalisation behind scipy.linalg.eig and solver.py:142:
//     psi <- normalise( D * Q * X ),  Q = H_0 H_1 ... H_{n-3},  H_k = I - tau_k v_k v_k^H
// (reflectors left below the subdiagonal of the work matrix by the Hessenberg reduction).
//
//   k_wy_tfactor   one warp per (frequency, block of 8 reflectors): T of the compact WY form
//                  H_k0 ... H_k0+7 = I - V T V^H  (zlarft, forward / columnwise).
//   k_backnorm_wy  CTA = one frequency x 64 modes, X block in shared memory.  Per block of 8
//                  reflectors two complex GEMMs on mma.sync.m8n8k4.f64 (DMMA; tcgen05 has no FP64
//                  kind):  W = V^H X (8 x 64),  W' = T W,  X -= V W'.  Warp w owns the 8 modes
//                  8w..8w+7 for every row, so X never crosses warps; only the V / T staging is
//                  shared (2 block barriers per 8 reflectors instead of 1 per reflector).
//                  X is read and written once per 8 reflectors instead of once per reflector.
// Fragment loads are bank-conflict free by construction: X row stride 65, V twice (stride 9 for
// the k-major fragments of V^H X with rows taken as {0,2,4,6} / {1,3,5,7}, stride 12 for the
// m-major fragments of V W'), W stride 10.
#include "common.cuh"
#include <stdlib.h>

#include "backx_kernels.cuh"                // k_wy_tfactor, k_backnorm_wy, backnorm_wy_smem

// returns -1 if this instance does not apply (caller falls back to k_backnorm)
extern "C" int axs_launch_backnorm_wy(int N, int hb, const cplx* Cb, const cplx* K6d, int nf,
                                      const cplx* eig, cplx* psi, AxsWork wk, cudaStream_t st)
{
    const int n = 2 * N;
    if (!axs_tuning()->backnorm_wy || n > WY_MAX_N || n < 16) return -1;   // 0: per-reflector slab kernel (k_backnorm)
    const int nblk = (n - 2 + WY_NB - 1) / WY_NB;
    // T factors live in the QR parking buffer (dead once the eigenvalues are out)
    if ((long long)nblk * WY_NB * WY_NB > (long long)hcol_size(n)) return -1;
    cplx* Tg = wk.Hp;
    k_wy_tfactor<<<dim3(nblk, nf), 32, 0, st>>>(n, wk, Tg, nblk);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    // modes per CTA: 64 (default, 12.2 ms per 2000 points with k_modes) or 32 (2 CTAs per SM, 14.3 ms: V staging twice as often)
    if (axs_tuning()->backnorm_wy_modes != 32) {
        const size_t smem = backnorm_wy_smem(n, 64);
        AXS_SMEM_ONCE(smem, k_backnorm_wy<64>);
        k_backnorm_wy<64><<<dim3((n + 63) / 64, nf), 256, smem, st>>>(N, hb, Cb, K6d, eig, psi, wk, Tg, nblk);
    } else {
        const size_t smem = backnorm_wy_smem(n, 32);
        AXS_SMEM_ONCE(smem, k_backnorm_wy<32>);
        k_backnorm_wy<32><<<dim3((n + 31) / 32, nf), 128, smem, st>>>(N, hb, Cb, K6d, eig, psi, wk, Tg, nblk);
    }
    return (int)cudaGetLastError();
}
