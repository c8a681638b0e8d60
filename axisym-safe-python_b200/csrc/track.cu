// Mode tracking (solver.py:151-182) for the SAFE sweep.
//
//   k_bpsi        Y = B psi  with  B = [[0, K6], [K6, C]]   (banded C, diagonal K6)
//   k_track_mma   P = psi_prev^T Y_cur on FP64 tensor-core tiles (DMMA) with a fused row arg-max /
//                 row-max epilogue (the reference forms the full biorth matrix with two dense
//                 n2^3 products, solver.py:157, and then only uses argmax_c |P[r,c]| and its
//                 value, :159-161).  The scalar-DFMA tile kernel it replaced (5.59 vs 2.7 ms per
//                 2000 points) is in the git history.
//   k_apply_order column gather of k and psi once the permutation scan has run.
//   axs_track_scan (host) the sequential integer scan o_i[r] = a_i[o_{i-1}[r]].
#include "common.cuh"
#include <stdlib.h>
#include "track_kernels.cuh"                // k_bpsi, k_bpsi_tiled, k_track_mma, k_apply_order

#include "post_kernels.cuh"                 // k_energy_ratio, k_quad_forms, k_modal_projection, k_rank_modes, k_zero_unselected


extern "C" int axs_launch_quad_forms(int N, int nf, const cplx* psi, const int* tclass, int n_sel, const int* sel,
                                     int n_mat, const int* mat_off, const int* rows, const int* cols,
                                     const cplx* vals, cplx* out, cudaStream_t st)
{
    k_quad_forms<<<dim3((n_sel + 127) / 128, nf), 128, 0, st>>>(N, psi, tclass, n_sel, sel, n_mat, mat_off, rows,
                                                               cols, vals, out);
    return (int)cudaGetLastError();
}


extern "C" int axs_launch_modal_projection(int n, int nf, const cplx* psi, const cplx* fvec, cplx* out, cudaStream_t st)
{
    k_modal_projection<<<dim3((n + 127) / 128, nf), 128, 0, st>>>(n, psi, fvec, out);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_energy_ratio(int N, int nf, const cplx* psi, const int* tclass,
                                       int nnz_t, const int* rt, const int* ct, const cplx* vt,
                                       int nnz_p, const int* rp, const int* cp, const cplx* vp,
                                       double* er, cudaStream_t st)
{
    const int n = 2 * N;
    k_energy_ratio<<<dim3((n + 127) / 128, nf), 128, 0, st>>>(N, psi, tclass, nnz_t, rt, ct, vt, nnz_p, rp, cp, vp, er);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_track(int N, int hb, const cplx* Cb, const cplx* K6d, const cplx* psi,
                                int npairs, cplx* Y, int* arg, double* amax, cudaStream_t st)
{
    const int n = 2 * N;
    cudaError_t e;
    {
        // tiled kernel when the lower half of a 64-mode block + a chunk of the band fit twice per SM
        const int W = 2 * hb + 1;
        const size_t ubytes = (size_t)N * BP_COLS * sizeof(cplx);
        const size_t budget = (size_t)110 * 1024;
        int crows = ubytes < budget ? (int)((budget - ubytes) / ((size_t)W * sizeof(cplx))) : 0;
        if (crows > N) crows = N;
        if (crows >= 8 && axs_tuning()->bpsi_tiled) {
            const size_t smemb = ubytes + (size_t)crows * W * sizeof(cplx);
            AXS_SMEM_ONCE(smemb, k_bpsi_tiled);
            k_bpsi_tiled<<<dim3((n + BP_COLS - 1) / BP_COLS, npairs + 1), 256, smemb, st>>>(N, hb, crows, Cb, K6d, psi, Y);
        } else {
            int bx = (n * n + 255) / 256; if (bx > 64) bx = 64;
            k_bpsi<<<dim3(bx, npairs + 1), 256, 0, st>>>(N, hb, Cb, K6d, psi, Y);
        }
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    size_t smem = (size_t)(TM_KC * TM_SA + TM_KC * TM_SB) * sizeof(cplx) + 8 * 32 * (sizeof(double) + sizeof(int));
    AXS_SMEM_ONCE(smem, k_track_mma);
    k_track_mma<<<dim3((n + 31) / 32, npairs), 256, smem, st>>>(n, psi, Y, arg, amax);
    return (int)cudaGetLastError();
}



extern "C" int axs_launch_select_modes(int n, int nf, const cplx* k, const double* sigma, int m, int* rank,
                                       cplx* psi, cudaStream_t st)
{
    k_rank_modes<<<nf, 256, 0, st>>>(n, k, sigma, rank);
    if (psi && m < n) {
        long long cells = (long long)n * n;
        int bx = (int)((cells + 255) / 256); if (bx > 64) bx = 64;
        k_zero_unselected<<<dim3(bx, nf), 256, 0, st>>>(n, m, rank, psi);
    }
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_apply_order(int n, int nf, const int* order, const cplx* k_in, const cplx* psi_in,
                                      cplx* k_out, cplx* psi_out, cudaStream_t st)
{
    int bx = (n * n + 255) / 256; if (bx > 64) bx = 64;
    k_apply_order<<<dim3(bx, nf), 256, 0, st>>>(n, order, k_in, psi_in, k_out, psi_out);
    return (int)cudaGetLastError();
}

// -----------------------------------------------------------------------------------------
// Small device -> pinned-host copies done by the SMs instead of the copy engine.  The pipelined sweeps
// download the mode shapes of a chunk (0.1 .. 0.3 GB) with cudaMemcpyAsync while the next chunk
// computes; the few hundred KB of tracking tables the host is waiting for would queue behind that bulk
// transfer on the same D2H engine.  Pinned host memory is mapped into the device address space (UVA),
// so a kernel can store to it directly and the stores interleave with the DMA traffic on the link.
__global__ void __launch_bounds__(256)
k_copy_words(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst, long long nwords)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

extern "C" int axs_launch_copy_to_pinned(const void* src, void* dst, long long bytes, cudaStream_t st)
{
    const long long nwords = bytes / 4;
    if (nwords <= 0) return 0;
    long long blocks = (nwords + 255) / 256;
    if (blocks > 296) blocks = 296;
    k_copy_words<<<(int)blocks, 256, 0, st>>>((const unsigned int*)src, (unsigned int*)dst, nwords);
    return (int)cudaGetLastError();
}
