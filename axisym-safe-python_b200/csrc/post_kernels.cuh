// The kernels behind energy_ratio / energy_velocity / point_excited_waves and the subset mode of solve()
// (see track.cu for the launchers).  A header of its own so that tests/host/post_emul.cpp can compile this very
// source as host code and run it on an emulated thread block (tests/host/cta_emul.h) without a GPU.
#pragma once
#include "common.cuh"

// Kinetic-energy ratio PML / total per mode (solver.py:186-269):
//   q = conj(T) psi[N:],  E = q^H M q  (M given as COO triplets, fluid columns already x -1),
//   er = |E_pml| / |E_tot|.   grid = (ceil(n2/128), nf), thread = one mode.
__global__ void __launch_bounds__(128)
k_energy_ratio(int N, const cplx* __restrict__ psi_all, const int* __restrict__ tclass,
               int nnz_t, const int* __restrict__ rt, const int* __restrict__ ct, const cplx* __restrict__ vt,
               int nnz_p, const int* __restrict__ rp, const int* __restrict__ cp, const cplx* __restrict__ vp,
               double* __restrict__ er)
{
    const int n = 2 * N, f = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const cplx* psi = psi_all + (long long)f * n * n + (long long)N * n + j;   // psi[N + r][j] = psi[r * n]
    auto q_at = [&](int r) {
        const cplx v = psi[(long long)r * n];
        return tclass[r] ? mk(v.y, -v.x) : v;                                  // conj(T) = 1 or -i
    };
    cplx et = mk(0, 0), ep = mk(0, 0);
    for (int e = 0; e < nnz_t; ++e) et = cfma(cmul(cconj(q_at(rt[e])), vt[e]), q_at(ct[e]), et);
    for (int e = 0; e < nnz_p; ++e) ep = cfma(cmul(cconj(q_at(rp[e])), vp[e]), q_at(cp[e]), ep);
    er[(long long)f * n + j] = cabsd(ep) / cabsd(et);
}

// Quadratic forms of the tracked mode shapes with a list of sparse matrices (energy_velocity,
// solver.py:374-409):  out[f][s][m] = q^H X_m q,  q = conj(T) psi[N:, sel[s]],  X_m given as COO
// triplets e in [mat_off[m], mat_off[m+1]).  grid = (ceil(n_sel/128), nf), thread = one mode.
__global__ void __launch_bounds__(128)
k_quad_forms(int N, const cplx* __restrict__ psi_all, const int* __restrict__ tclass, int n_sel,
             const int* __restrict__ sel, int n_mat, const int* __restrict__ mat_off,
             const int* __restrict__ rows, const int* __restrict__ cols, const cplx* __restrict__ vals,
             cplx* __restrict__ out)
{
    const int n = 2 * N, f = blockIdx.y, s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sel) return;
    const int j = sel ? sel[s] : s;
    const cplx* psi = psi_all + (long long)f * n * n + (long long)N * n + j;
    auto q_at = [&](int r) {
        const cplx v = psi[(long long)r * n];
        return tclass[r] ? mk(v.y, -v.x) : v;                                  // conj(T) = 1 or -i
    };
    for (int m = 0; m < n_mat; ++m) {
        cplx acc = mk(0, 0);
        for (int e = mat_off[m]; e < mat_off[m + 1]; ++e)
            acc = cfma(cmul(cconj(q_at(rows[e])), vals[e]), q_at(cols[e]), acc);
        out[((long long)f * n_sel + s) * n_mat + m] = acc;
    }
}

// Modal projection (point_excited_waves, solver.py:411-435): out[f][j] = sum_r psi[f][r][j] * fvec[r]
// over all 2N rows (the reference multiplies psi_hat^T by the force vector [0; f_ext]).
__global__ void __launch_bounds__(128)
k_modal_projection(int n, const cplx* __restrict__ psi_all, const cplx* __restrict__ fvec, cplx* __restrict__ out)
{
    const int f = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const cplx* psi = psi_all + (long long)f * n * n + j;
    cplx acc = mk(0, 0);
    for (int r = 0; r < n; ++r) {
        const cplx fr = fvec[r];
        if (fr.x != 0.0 || fr.y != 0.0) acc = cfma(psi[(long long)r * n], fr, acc);
    }
    out[(long long)f * n + j] = acc;
}

// -----------------------------------------------------------------------------------------
// Subset mode of solve() (solver.py:137-140: scipy.sparse.linalg.eigs(S, k=m, sigma=w/c, which='LM'),
// i.e. the m eigenvalues nearest to sigma).  The whole spectrum is already on the device, so the
// subset is a selection:  k_rank_modes ranks the modes of one frequency by |k - sigma| (counting
// rank, ties by native index; CTA = one frequency, the spectrum staged in shared memory in tiles),
// k_zero_unselected clears the mode-shape columns of rank >= m so that the tracking arg-max
// (k_track_mma) only ever sees the selected columns.
__global__ void __launch_bounds__(256)
k_rank_modes(int n, const cplx* __restrict__ k_all, const double* __restrict__ sigma, int* __restrict__ rank_all)
{
    __shared__ double sd[1024];
    const int f = blockIdx.x, tid = threadIdx.x;
    const cplx* k = k_all + (long long)f * n;
    const double sg = sigma[f];
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int j = j0 + tid;
        double dj = 0.0;
        if (j < n) { const double dx = k[j].x - sg, dy = k[j].y; dj = dx * dx + dy * dy; }
        int cnt = 0;
        for (int l0 = 0; l0 < n; l0 += 1024) {
            __syncthreads();
            for (int l = tid; l < 1024 && l0 + l < n; l += 256) {
                const double dx = k[l0 + l].x - sg, dy = k[l0 + l].y;
                sd[l] = dx * dx + dy * dy;
            }
            __syncthreads();
            const int lim = (n - l0 < 1024) ? n - l0 : 1024;
            if (j < n)
                for (int l = 0; l < lim; ++l) {
                    const double dl = sd[l];
                    cnt += (dl < dj || (dl == dj && l0 + l < j)) ? 1 : 0;
                }
        }
        if (j < n) rank_all[(long long)f * n + j] = cnt;
    }
}

__global__ void __launch_bounds__(256)
k_zero_unselected(int n, int m, const int* __restrict__ rank_all, cplx* __restrict__ psi_all)
{
    const int f = blockIdx.y;
    const int* rank = rank_all + (long long)f * n;
    cplx* psi = psi_all + (long long)f * n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        const int c = idx % n;
        if (rank[c] >= m) psi[idx] = mk(0, 0);
    }
}
