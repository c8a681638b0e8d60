// Kernels 2 + 3 of the SAFE sweep, global-memory ("large", 166 < 2N <= 8192) instance.
// Same algorithm and results as eig_small.cu (zgebal-style scaling, Householder Hessenberg,
// multi-bulge single-shift QR, inverse iteration on H, back-transformation, B-normalisation;
// together they replace scipy.linalg.eig / LAPACK zgeev, solver.py:136, and solver.py:142),
// but the matrix of one frequency no longer fits the shared memory of one SM:
//
//   k_reduce_big  S(w), |a|^2 tables and the zgebal scaling (one CTA per frequency); with balance_only
//                 the Hessenberg reduction itself is done by the blocked compact-WY / DMMA kernels of
//                 eig_bigred.cu (default); otherwise (tuning big_blocked = 0) by the unblocked fused
//                 rank-2 update below: 2-D (row block x column group) decomposition over the warps,
//                 y = A^H v and w = A v accumulated with fixed-order reductions.
//   k_hqr_big     a chain of 8 bulges is chased through a 48 x 48 DIAGONAL WINDOW held in shared memory
//                 for 24 steps ("phase D", the latency bound part, identical to k_hqr_mb), the 24 x 8
//                 rotations are recorded, and the off-diagonal strips (columns right of the window, rows
//                 above it) are then updated from 128-wide shared-memory tiles, one thread per
//                 column/row, bulge-major so that the running element stays in a register ("phase O").
//                 One CTA per frequency, or - for batches that do not fill the GPU - a thread-block
//                 cluster of 2 / 4 / 8 CTAs per frequency that shares the strip tiles.  Once the active
//                 block fits the shared-memory kernels the sweep is handed over to the k_hqr_mb stages
//                 of eig_small.cu.
//   k_modes_col   inverse iteration (H - k I) x = 1 as a bottom-up pivoted LU with column operations,
//                 column-oriented: a CTA sweeps the columns of H from right to left, every thread owns
//                 rows (cyclic) for M modes in registers, one barrier per column.  State that has to
//                 survive (multipliers, pivots) is O(n) per mode.
//   k_backx_slab / k_backx_big   per-reflector back-transformation (slab of mode vectors in shared memory
//                 while it fits; registers above that when the blocked form of eig_bigred.cu is off).
//   k_norm_big    un-balanced vectors -> psi = v / sqrt(2k u^T K6 u + u^T C u), psi_hat layout.
#include "common.cuh"
#include <stdlib.h>
#include "big_kernels.cuh"                  // every kernel of this file (cooperative_groups, namespace cg)

extern "C" int axs_launch_hqr_stages(int n, int nf, int m_first, cplx* Hp, cplx* eig, int* info, int* ucur,
                                     int* stats, cudaStream_t st);
extern "C" int axs_hqr_stage_capacity(int per_sm);

extern "C" int axs_launch_backx_blocked(int n, int nf, cplx* psi, AxsWorkBig wk, cudaStream_t st);
extern "C" int axs_launch_reduce_blocked(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                         const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                         const cplx* S_in, AxsWorkBig wk, cudaStream_t st);

// S(w), |a|^2 tables and the zgebal scaling only (phases A + B of k_reduce_big)
extern "C" int axs_launch_balance_big(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                      const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                      const cplx* S_in, AxsWorkBig wk, cudaStream_t st)
{
    const int n = 2 * N;
    const size_t smem = (size_t)n * sizeof(cplx);
    AXS_SMEM_ONCE(smem, k_reduce_big);
    k_reduce_big<<<nf, RB_NT, smem, st>>>(N, hb, K0s, Cb, Mb, K1c, K6d, omega, S_in, wk, 1);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_reduce_big(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                     const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                     const cplx* S_in, AxsWorkBig wk, cudaStream_t st)
{
    const int n = 2 * N;
    // default: blocked compact-WY reduction with DMMA trailing updates over the whole GPU (eig_bigred.cu);
    // tuning big_blocked = 0: one CTA per frequency, unblocked (k_reduce_big)
    if (axs_tuning()->big_blocked)
        return axs_launch_reduce_blocked(N, hb, K0s, Cb, Mb, K1c, K6d, omega, nf, S_in, wk, st);
    const size_t smem = (size_t)n * sizeof(cplx);
    AXS_SMEM_ONCE(smem, k_reduce_big);
    k_reduce_big<<<nf, RB_NT, smem, st>>>(N, hb, K0s, Cb, Mb, K1c, K6d, omega, S_in, wk, 0);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_hqr_big(int n, int nf, cplx* eig, int* info, AxsWorkBig wk, cudaStream_t st)
{
    const int m2 = axs_hqr_stage_capacity(2);            // what the 2-CTA/SM shared-memory stage can hold
    const size_t smem = hqr_big_smem();
    AXS_SMEM_ONCE(smem, k_hqr_big<false>);
    AXS_SMEM_ONCE(smem, k_hqr_big<true>);
    // cluster size: the CTAs of a cluster share the strip tiles of one matrix; a batch that fills the GPU
    // (2 CTAs per SM) keeps one CTA per matrix
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int CL = 1;
    while (CL < 8 && (long long)nf * CL * 2 <= sms && CL * 2 * HB_TILE <= n) CL *= 2;
    if (axs_tuning()->big_qr_cluster > 0) CL = axs_tuning()->big_qr_cluster;       // A/B: 1, 2, 4, 8
    cplx* Hc = wk.Hc;
    int* ucur = wk.ucur;
    if (CL == 1) {
        k_hqr_big<false><<<nf, HB_NT, smem, st>>>(n, m2, Hc, eig, info, ucur);
        cudaError_t e1 = cudaGetLastError();
        if (e1 != cudaSuccess) return (int)e1;
        return axs_launch_hqr_stages(n, nf, m2, wk.Hc, eig, info, wk.ucur, nullptr, st);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nf * CL), 1, 1);
    cfg.blockDim = dim3(HB_NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_hqr_big<true>, n, m2, Hc, eig, info, ucur);
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return axs_launch_hqr_stages(n, nf, m2, wk.Hc, eig, info, wk.ucur, nullptr, st);
}

template <int NT, int R, int M>
static int launch_modes_col(int n, int nf, const cplx* eig, cplx* psi, AxsWorkBig wk, cudaStream_t st)
{
    const int groups = (n + M - 1) / M;
    long long items = (long long)nf * groups;
    int grid = items < AXS_BIG_MODE_GRID ? (int)items : AXS_BIG_MODE_GRID;
    const size_t smem = (size_t)n * sizeof(cplx);
    if (smem > 48 * 1024) AXS_SMEM_ONCE(smem, k_modes_col<NT, R, M>);
    k_modes_col<NT, R, M><<<grid, NT, smem, st>>>(n, nf, eig, psi, wk);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_modes_big(int N, int hb, const cplx* Cb, const cplx* K6d, int nf,
                                    const cplx* eig, cplx* psi, AxsWorkBig wk, cudaStream_t st)
{
    const int n = 2 * N;
    int rc;
    // <threads, rows per thread, modes per CTA>; M must equal axs_big_modes_per_cta(n)
    if (n <= 512) rc = launch_modes_col<256, 2, 4>(n, nf, eig, psi, wk, st);
    else if (n <= 768) rc = launch_modes_col<256, 3, 2>(n, nf, eig, psi, wk, st);
    else if (n <= 1024) rc = launch_modes_col<256, 4, 2>(n, nf, eig, psi, wk, st);
    else if (n <= 1280) rc = launch_modes_col<256, 5, 1>(n, nf, eig, psi, wk, st);
    else if (n <= 1536) rc = launch_modes_col<256, 6, 1>(n, nf, eig, psi, wk, st);
    else if (n <= 2048) rc = launch_modes_col<256, 8, 1>(n, nf, eig, psi, wk, st);
    else if (n <= 4096) rc = launch_modes_col<512, 8, 1>(n, nf, eig, psi, wk, st);
    else rc = launch_modes_col<1024, 8, 1>(n, nf, eig, psi, wk, st);
    if (rc) return rc;
    cudaError_t e;
    int Cs = 8;                                          // slab width that fits next to 3 reflector buffers
    while (Cs >= 1 && (size_t)(Cs + 4) * n * sizeof(cplx) > 224 * 1024) Cs >>= 1;
    if (Cs < 4 && axs_tuning()->big_blocked && axs_tuning()->big_panel >= 32) {
        // sizes whose slab of mode vectors no longer fits an SM: blocked compact-WY back-transformation on
        // DMMA tiles with the T factors the blocked reduction kept (eig_bigred.cu) - level 3 instead of
        // one pass over all reflectors per mode
        rc = axs_launch_backx_blocked(n, nf, psi, wk, st);
        if (rc) return rc;
        k_norm_big<<<dim3((n + 31) / 32, nf), 256, 0, st>>>(N, hb, Cb, K6d, eig, psi);
        return (int)cudaGetLastError();
    }
    if (Cs >= 1) {
        const size_t smem = (size_t)(Cs + 4) * n * sizeof(cplx);
        AXS_SMEM_ONCE(smem, k_backx_slab);
        k_backx_slab<<<dim3((n + Cs - 1) / Cs, nf), 256, smem, st>>>(n, Cs, psi, wk);
        e = cudaGetLastError();
    } else {
        int T = 32, NTb = 256;
        while (T * BX_RS < n) T *= 2;                    // 32 .. 512 threads per mode
        if (T > 256) NTb = 512;
        const int C = NTb / T;
        k_backx_big<<<dim3((n + C - 1) / C, nf), NTb, 0, st>>>(n, T, psi, wk);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return (int)e;
    k_norm_big<<<dim3((n + 31) / 32, nf), 256, 0, st>>>(N, hb, Cb, K6d, eig, psi);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_unpack_big(int n, int nf, AxsWorkBig wk, cplx* Hd, cudaStream_t st)
{
    k_unpack_big<<<dim3(64, nf), 256, 0, st>>>(n, wk, Hd);
    return (int)cudaGetLastError();
}
