// Kernel 1 of the SAFE sweep: element quadrature, operator scatter, dense S(w) batch.
//
//   k_element_matrices   one warp per element; quadrature tables (ksi, w, D) and the
//                        per-point geometry (J, 1/r, 1/gamma, weight) staged in shared
//                        memory; replaces the per-point Python loops of
//                        elements.py:157-217, :302-322, :457-529, :691-723, :818-887, :970-1003.
//   k_scatter_operators  one warp per element; T-transform folded in as a per-entry factor
//                        (solver.py:59-74), output = banded K0s, Cb, Mb and diag K6
//                        (mesh.py:518-619 + solver.py:115-133).
//   k_build_S            frequency-batched companion matrix S(w) (solver.py:115-134),
//                        column-major, 16-byte stores, write-only HBM traffic 16*n2^2 B/point.
#include "common.cuh"

#include "assemble_kernels.cuh"             // k_element_matrices, k_scatter_operators, k_build_S

// ------------------------------------------------------------------------------ launchers
extern "C" int axs_launch_element_matrices(int n_elem, const int* edesc_i, const double* edesc_d,
                                           const double* tables, const double* nodes, cplx* out,
                                           cudaStream_t st)
{
    const int warps = 4;
    size_t per_warp = (size_t)EL_MAX_M * EL_MAX_M * sizeof(double) + EL_MAX_M * sizeof(PointGeom)
                      + EL_MAX_M * sizeof(double);
    size_t smem = warps * per_warp;
    AXS_CUDA_OK(cudaFuncSetAttribute(k_element_matrices, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_element_matrices<<<(n_elem + warps - 1) / warps, warps * 32, smem, st>>>(n_elem, edesc_i, edesc_d, tables, nodes, out);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_scatter(int n_elem, const int* edesc_i, const cplx* elem, const int* lm,
                                  const int* lm_off, const int* tclass, int n_circ, int N, int hb,
                                  cplx* K0s, cplx* Cb, cplx* Mb, cplx* K6d, cudaStream_t st)
{
    const int warps = 4;
    k_scatter_operators<<<(n_elem + warps - 1) / warps, warps * 32, 0, st>>>(
        n_elem, edesc_i, elem, lm, lm_off, tclass, n_circ, N, hb, K0s, Cb, Mb, K6d);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_build_S(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                  const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                  cplx* S, cudaStream_t st)
{
    const int n2 = 2 * N;
    int bx = (n2 * n2 + 255) / 256;
    if (bx > 64) bx = 64;
    k_build_S<<<dim3(bx, nf), 256, 0, st>>>(N, hb, K0s, Cb, Mb, K1c, K6d, omega, S);
    return (int)cudaGetLastError();
}
