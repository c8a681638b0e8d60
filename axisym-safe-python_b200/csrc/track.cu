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

// Y[f][r][j]:  r <  N : K6[r] * psi[N + r][j]
//              r >= N : K6[r-N] * psi[r-N][j] + sum_c C[r-N][c] psi[N + c][j]
__global__ void __launch_bounds__(256)
k_bpsi(int N, int hb, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
       const cplx* __restrict__ psi_all, cplx* __restrict__ Y_all)
{
    const int n = 2 * N, W = 2 * hb + 1;
    const int f = blockIdx.y;
    const cplx* psi = psi_all + (long long)f * n * n;
    cplx* Y = Y_all + (long long)f * n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        const int r = idx / n, j = idx - r * n;
        cplx acc;
        if (r < N) {
            acc = cmul(K6d[r], psi[(long long)(N + r) * n + j]);
        } else {
            const int rr = r - N;
            acc = cmul(K6d[rr], psi[(long long)rr * n + j]);
            const int c0 = max(0, rr - hb), c1 = min(N - 1, rr + hb);
            for (int c = c0; c <= c1; ++c)
                acc = cfma(Cb[(long long)rr * W + c - rr + hb], psi[(long long)(N + c) * n + j], acc);
        }
        Y[idx] = acc;
    }
}

// Tiled form of k_bpsi: CTA = one frequency x 64 modes.  The lower half of the mode block (63 x 64
// at 2N = 126) is staged in shared memory once and reused by every row of the banded C, which is
// streamed through shared memory in row chunks; thread = (mode, row group).
#define BP_COLS 64
__global__ void __launch_bounds__(256, 2)
k_bpsi_tiled(int N, int hb, int crows, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
             const cplx* __restrict__ psi_all, cplx* __restrict__ Y_all)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = 2 * N, W = 2 * hb + 1;
    const int f = blockIdx.y, tid = threadIdx.x;
    const int col = tid & (BP_COLS - 1), grp = tid >> 6;
    const int mode = blockIdx.x * BP_COLS + col;
    const bool live = mode < n;
    cplx* U = (cplx*)smem_raw;                        // [N][BP_COLS] lower half of the block
    cplx* Cs = U + (size_t)N * BP_COLS;               // [crows][W]
    const cplx* psi = psi_all + (long long)f * n * n;
    cplx* Y = Y_all + (long long)f * n * n;
    for (int idx = tid; idx < N * BP_COLS; idx += 256) {
        const int r = idx / BP_COLS, c = idx - r * BP_COLS;
        const int m = blockIdx.x * BP_COLS + c;
        U[idx] = (m < n) ? psi[(long long)(N + r) * n + m] : mk(0, 0);
    }
    for (int r0 = 0; r0 < N; r0 += crows) {
        const int r1 = min(N, r0 + crows);
        __syncthreads();                              // U staged / previous chunk consumed
        for (int idx = tid; idx < (r1 - r0) * W; idx += 256) Cs[idx] = Cb[(long long)r0 * W + idx];
        __syncthreads();
        int r = r0 + ((grp - r0) & 3);
        for (; r < r1; r += 4) {
            const cplx k6 = K6d[r];
            cplx acc0 = live ? cmul(k6, psi[(long long)r * n + mode]) : mk(0, 0), acc1 = mk(0, 0);
            const int c0 = max(0, r - hb), c1 = min(N - 1, r + hb);
            const cplx* crow = Cs + (r - r0) * W - r + hb;
            int c = c0;
            for (; c + 1 <= c1; c += 2) {
                acc0 = cfma(crow[c], U[c * BP_COLS + col], acc0);
                acc1 = cfma(crow[c + 1], U[(c + 1) * BP_COLS + col], acc1);
            }
            if (c <= c1) acc0 = cfma(crow[c], U[c * BP_COLS + col], acc0);
            if (live) {
                Y[(long long)(N + r) * n + mode] = cadd(acc0, acc1);
                Y[(long long)r * n + mode] = cmul(k6, U[r * BP_COLS + col]);
            }
        }
    }
}

// -----------------------------------------------------------------------------------------
// k_track_mma: the contraction P[a][b] = sum_r psiP[r][a] * Y[r][b] on the FP64 tensor path
// (mma.sync.aligned.m8n8k4.f64, "DMMA"; tcgen05 has no FP64 kind).  A complex product is four
// real MMAs per 8x8x4 tile: Pre += Are Bre + (-Aim) Bim,  Pim += Are Bim + Aim Bre.
// CTA = 32 rows (a) x all columns (b <= 192), 8 warps; warp w owns the n8 column tiles
// w, w+8, w+16 for all four m8 row tiles.  Shared tiles are padded so that the fragment loads
// (4 k-rows x 8 columns per quarter-warp) are bank-conflict free.
#define TM_KC 16
#define TM_SA 34                       // row stride of sA in cplx (32 + 2 -> 544 B = 32 mod 128)
#define TM_SB 194                      // row stride of sB in cplx (192 + 2)
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2)
k_track_mma(int n, const cplx* __restrict__ psi_all, const cplx* __restrict__ Y_all,
            int* __restrict__ arg_all, double* __restrict__ amax_all)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* sA = (cplx*)smem_raw;                       // [TM_KC][TM_SA]
    cplx* sB = sA + TM_KC * TM_SA;                    // [TM_KC][TM_SB]
    double* s_best = (double*)(sB + TM_KC * TM_SB);   // [8 warps][32 rows]
    int* s_bidx = (int*)(s_best + 8 * 32);
    const int p = blockIdx.y, a0 = blockIdx.x * 32;
    const cplx* psiP = psi_all + (long long)p * n * n;
    const cplx* Y = Y_all + (long long)(p + 1) * n * n;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;          // groupID, thread-in-group
    double best[4];
    int bidx[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) { best[mt] = -1.0; bidx[mt] = 0; }

    // columns b in chunks of 192 (one chunk for the shared-memory eigen path, 2N <= 166)
    for (int b0 = 0; b0 < n; b0 += 192) {
        const int nbc = (n - b0 < 192) ? n - b0 : 192;
        const int ntiles = (nbc + 7) >> 3;
        double cre[4][3][2], cim[4][3][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int t = 0; t < 3; ++t) { cre[mt][t][0] = cre[mt][t][1] = 0.0; cim[mt][t][0] = cim[mt][t][1] = 0.0; }

        for (int r0 = 0; r0 < n; r0 += TM_KC) {
            for (int idx = tid; idx < TM_KC * 32; idx += 256) {
                const int rr = idx >> 5, aa = idx & 31, r = r0 + rr, a = a0 + aa;
                sA[rr * TM_SA + aa] = (r < n && a < n) ? psiP[(long long)r * n + a] : mk(0, 0);
            }
            for (int idx = tid; idx < TM_KC * 192; idx += 256) {
                const int rr = idx / 192, bb = idx - rr * 192, r = r0 + rr;
                sB[rr * TM_SB + bb] = (r < n && bb < nbc) ? Y[(long long)r * n + b0 + bb] : mk(0, 0);
            }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < TM_KC; ks += 4) {
                cplx af[4];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) af[mt] = sA[(ks + tq) * TM_SA + mt * 8 + gq];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int nt = w + 8 * t;
                    if (nt < ntiles) {
                        const cplx bf = sB[(ks + tq) * TM_SB + nt * 8 + gq];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            dmma(cre[mt][t][0], cre[mt][t][1], af[mt].x, bf.x);
                            dmma(cre[mt][t][0], cre[mt][t][1], -af[mt].y, bf.y);
                            dmma(cim[mt][t][0], cim[mt][t][1], af[mt].x, bf.y);
                            dmma(cim[mt][t][0], cim[mt][t][1], af[mt].y, bf.x);
                        }
                    }
                }
            }
            __syncthreads();
        }
        // running row arg-max of |P|^2 (lowest column index wins ties, like np.argmax)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int nt = w + 8 * t;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int bl = nt * 8 + 2 * tq + e;
                    const double m2 = (nt < ntiles && bl < nbc) ? cre[mt][t][e] * cre[mt][t][e] + cim[mt][t][e] * cim[mt][t][e] : -1.0;
                    const int b = b0 + bl;
                    if (m2 > best[mt] || (m2 == best[mt] && b < bidx[mt])) { best[mt] = m2; bidx[mt] = b; }
                }
            }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        double bb = best[mt];
        int bi = bidx[mt];
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {              // across the 4 lanes of a group
            const double ob = __shfl_xor_sync(0xffffffffu, bb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > bb || (ob == bb && oi < bi)) { bb = ob; bi = oi; }
        }
        if (tq == 0) { s_best[w * 32 + mt * 8 + gq] = bb; s_bidx[w * 32 + mt * 8 + gq] = bi; }
    }
    __syncthreads();
    if (tid < 32) {
        double bb = s_best[tid];
        int bi = s_bidx[tid];
        for (int ww = 1; ww < 8; ++ww) {
            const double ob = s_best[ww * 32 + tid];
            const int oi = s_bidx[ww * 32 + tid];
            if (ob > bb || (ob == bb && oi < bi)) { bb = ob; bi = oi; }
        }
        const int a = a0 + tid;
        if (a < n) { arg_all[(long long)p * n + a] = bi; amax_all[(long long)p * n + a] = sqrt(bb); }
    }
}

__global__ void __launch_bounds__(256)
k_apply_order(int n, const int* __restrict__ order, const cplx* __restrict__ k_in,
              const cplx* __restrict__ psi_in, cplx* __restrict__ k_out, cplx* __restrict__ psi_out)
{
    const int f = blockIdx.y;
    const int* ord = order + (long long)f * n;
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < n; c += blockDim.x) k_out[(long long)f * n + c] = k_in[(long long)f * n + ord[c]];
    if (psi_in && psi_out) {
        const cplx* pi = psi_in + (long long)f * n * n;
        cplx* po = psi_out + (long long)f * n * n;
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
            int r = idx / n, c = idx - r * n;
            po[idx] = pi[(long long)r * n + ord[c]];
        }
    }
}

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

extern "C" int axs_launch_quad_forms(int N, int nf, const cplx* psi, const int* tclass, int n_sel, const int* sel,
                                     int n_mat, const int* mat_off, const int* rows, const int* cols,
                                     const cplx* vals, cplx* out, cudaStream_t st)
{
    k_quad_forms<<<dim3((n_sel + 127) / 128, nf), 128, 0, st>>>(N, psi, tclass, n_sel, sel, n_mat, mat_off, rows,
                                                               cols, vals, out);
    return (int)cudaGetLastError();
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
