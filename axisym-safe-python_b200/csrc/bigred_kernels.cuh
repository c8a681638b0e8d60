// The kernels of the blocked compact-WY Hessenberg reduction and back-transformation on FP64 tensor-core
// tiles (see eig_bigred.cu for the launch sequences and the overview).  A header of its own so that
// tests/host/bigred_emul.cpp can compile this very source as host code and run it, CTA by CTA, on the
// emulated thread block of tests/host/cta_emul.h without a GPU.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define BR_NB_MAX 32
#define BR_COL_NT 512
#define BR_COL_NW (BR_COL_NT / 32)
#define BR_QC 8                         // reflectors per pass of the row loops in k_br_col

// reflector q of the panel (column k0+q) at row i: unit at i = k0+q+1, zero above, stored below
__device__ __forceinline__ cplx br_v(const cplx* __restrict__ A, int n, int k0, int i, int q) {
    const int d = i - (k0 + q + 1);
    return d > 0 ? A[(long long)(k0 + q) * n + i] : (d == 0 ? mk(1.0, 0.0) : mk(0.0, 0.0));
}

struct BrScratch {
    cplx* Y;      // [nb][n]   Y = A V T, column q at Y + q n
    cplx* W;      // [nb][n]   W = T^H V^H A(:, trailing), row q at W + q n (indexed by absolute column)
    cplx* part;   // [CG][n]   partial sums of the products A v
    cplx* T;      // [nb][nb]  row-major upper triangular
    cplx* Tall;   // [panels][nb][nb] the T factor of every panel, kept for the back-transformation
};
#define BR_TALL_ROW 88                  // Tall starts at row 88 of the yw region (32 rows of n entries)
__device__ __host__ __forceinline__ BrScratch br_scratch(cplx* ywb, int n, int CG) {
    BrScratch s;
    s.Y = ywb;
    s.W = ywb + (long long)BR_NB_MAX * n;
    s.part = ywb + 2LL * BR_NB_MAX * n;
    s.T = ywb + (2LL * BR_NB_MAX + CG) * n;
    s.Tall = ywb + (long long)BR_TALL_ROW * n;
    return s;
}

// deterministic block sum of BR_QC complex numbers per thread -> out[BR_QC] in shared memory
__device__ __forceinline__ void br_block_csum(cplx (&acc)[BR_QC], cplx* s_part /*[NW][QC]*/, cplx* out, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q < BR_QC; ++q) acc[q] = warp_csum(acc[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < BR_QC; ++q) s_part[warp * BR_QC + q] = acc[q];
    }
    __syncthreads();
    if (tid < BR_QC) {
        cplx s = mk(0, 0);
        for (int w = 0; w < BR_COL_NW; ++w) s = cadd(s, s_part[w * BR_QC + tid]);
        out[tid] = s;
    }
    __syncthreads();
}

// -----------------------------------------------------------------------------------------
// k_br_col: the sequential part of a panel column (see the file header).  j = index inside the
// panel; finish_only: only complete Y / T of column j-1 (after the last column of a panel).
__global__ void __launch_bounds__(BR_COL_NT, 1)
k_br_col(int n, int k0, int j, int CG, int finish_only, AxsWorkBig wk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    cplx* b_s = (cplx*)smem_raw;                           // [n] the column being updated
    __shared__ cplx s_z[BR_NB_MAX], s_w[BR_NB_MAX], s_wt[BR_NB_MAX], s_vc[BR_NB_MAX];
    __shared__ cplx s_red[BR_COL_NW * BR_QC], s_out[BR_QC];
    __shared__ double s_dred[BR_COL_NW];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cplx* A = wk.A + (long long)f * n * n;
    cplx* tau = wk.tau + (long long)f * n;
    const BrScratch sc = br_scratch(wk.yw + (long long)f * AXS_BIG_YW * n, n, CG);
    const int nb = BR_NB_MAX;

    // ---- a. finish column jp = j-1: z = V^H v_jp, Y(:, jp) = tau (A v - Y z), T(:, jp) ----------------
    if (j > 0) {
        const int jp = j - 1, cp = k0 + jp;
        const cplx tp = tau[cp];
        const cplx* vp = A + (long long)cp * n;             // v_jp below row cp+1 (unit at cp+1)
        for (int q0 = 0; q0 < jp; q0 += BR_QC) {
            cplx acc[BR_QC];
#pragma unroll
            for (int q = 0; q < BR_QC; ++q) acc[q] = mk(0, 0);
            for (int i = cp + 1 + tid; i < n; i += BR_COL_NT) {
                const cplx vi = (i == cp + 1) ? mk(1.0, 0.0) : vp[i];
#pragma unroll
                for (int q = 0; q < BR_QC; ++q)
                    if (q0 + q < jp) acc[q] = cfmac(vi, A[(long long)(k0 + q0 + q) * n + i], acc[q]);   // conj(V(i,q)) v_i
            }
            br_block_csum(acc, s_red, s_out, tid);
            if (tid < BR_QC && q0 + tid < jp) s_z[q0 + tid] = s_out[tid];
            __syncthreads();
        }
        for (int i = k0 + 1 + tid; i < n; i += BR_COL_NT) {
            cplx yv = mk(0, 0);
            for (int g = 0; g < CG; ++g) yv = cadd(yv, sc.part[(long long)g * n + i]);
            for (int q = 0; q < jp; ++q) yv = cfms(sc.Y[(long long)q * n + i], s_z[q], yv);
            sc.Y[(long long)jp * n + i] = cmul(tp, yv);
        }
        if (tid < jp) {                                      // T(q, jp) = -tau sum_{p >= q} T(q, p) z_p
            cplx s = mk(0, 0);
            for (int p = tid; p < jp; ++p) s = cfma(sc.T[tid * nb + p], s_z[p], s);
            sc.T[tid * nb + jp] = cneg(cmul(tp, s));
        }
        if (tid == 0) sc.T[jp * nb + jp] = tp;
        __syncthreads();
    }
    if (finish_only) return;

    // ---- b. column c up to date with the panel's j reflectors ------------------------------------------
    const int c = k0 + j;
    cplx* colc = A + (long long)c * n;
    if (tid < j) s_vc[tid] = cconj(br_v(A, n, k0, c, tid));   // conj(V(c, q))
    __syncthreads();
    for (int i = k0 + 1 + tid; i < n; i += BR_COL_NT) {
        cplx b = colc[i];
        for (int q = 0; q < j; ++q) b = cfms(sc.Y[(long long)q * n + i], s_vc[q], b);
        b_s[i] = b;
    }
    __syncthreads();
    if (j > 0) {
        for (int q0 = 0; q0 < j; q0 += BR_QC) {               // w = V^H b
            cplx acc[BR_QC];
#pragma unroll
            for (int q = 0; q < BR_QC; ++q) acc[q] = mk(0, 0);
            for (int i = k0 + 1 + tid; i < n; i += BR_COL_NT) {
                const cplx bi = b_s[i];
#pragma unroll
                for (int q = 0; q < BR_QC; ++q)
                    if (q0 + q < j) acc[q] = cfmac(bi, br_v(A, n, k0, i, q0 + q), acc[q]);
            }
            br_block_csum(acc, s_red, s_out, tid);
            if (tid < BR_QC && q0 + tid < j) s_w[q0 + tid] = s_out[tid];
            __syncthreads();
        }
        if (tid < j) {                                        // w' = T^H w
            cplx s = mk(0, 0);
            for (int p = 0; p <= tid; ++p) s = cfmac(s_w[p], sc.T[p * nb + tid], s);
            s_wt[tid] = s;
        }
        __syncthreads();
        for (int i = k0 + 1 + tid; i < n; i += BR_COL_NT) {   // b -= V w'
            cplx b = b_s[i];
            for (int q = 0; q < j; ++q) b = cfms(br_v(A, n, k0, i, q), s_wt[q], b);
            b_s[i] = b;
        }
        __syncthreads();
    }
    // ---- c. reflector from b (zlarfg): annihilates rows c+2 .. n-1 --------------------------------------
    double part = 0.0;
    for (int i = c + 2 + tid; i < n; i += BR_COL_NT) part += cabs2(b_s[i]);
    part = warp_sum(part);
    if (lane == 0) s_dred[warp] = part;
    __syncthreads();
    double xnorm2 = 0.0;
#pragma unroll
    for (int w = 0; w < BR_COL_NW; ++w) xnorm2 += s_dred[w];
    const cplx alpha = b_s[c + 1];
    cplx tj = mk(0, 0), scl = mk(0, 0);
    double beta = alpha.x;
    const bool trivial = (xnorm2 == 0.0 && alpha.y == 0.0);
    if (!trivial) {
        beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
        tj = mk((beta - alpha.x) / beta, -alpha.y / beta);
        scl = cdiv(mk(1.0, 0.0), mk(alpha.x - beta, alpha.y));
    }
    for (int i = k0 + 1 + tid; i < n; i += BR_COL_NT) {
        cplx o;
        if (i <= c) o = b_s[i];
        else if (i == c + 1) o = trivial ? alpha : mk(beta, 0.0);
        else o = trivial ? mk(0, 0) : cmul(b_s[i], scl);
        colc[i] = o;
    }
    if (tid == 0) tau[c] = tj;
}

// -----------------------------------------------------------------------------------------
// k_br_gemv: part[g][i] = sum over the columns jj of group g of A(i, jj) v(jj), rows i = k0+1 .. n-1,
// columns c+1 .. n-1 (v(c+1) = 1, v below it stored in column c).  grid = (row blocks * CG, nf).
#define BR_GV_NT 256
#define BR_GV_VS 512                    // v entries staged per pass
__global__ void __launch_bounds__(BR_GV_NT)
k_br_gemv(int n, int k0, int c, int CG, AxsWorkBig wk)
{
    __shared__ cplx s_v[BR_GV_VS];
    const int f = blockIdx.y, tid = threadIdx.x;
    const int g = blockIdx.x % CG, rb = blockIdx.x / CG;
    const cplx* A = wk.A + (long long)f * n * n;
    const BrScratch sc = br_scratch(wk.yw + (long long)f * AXS_BIG_YW * n, n, CG);
    const int ncols = n - c - 1;
    const int L = (ncols + CG - 1) / CG;
    const int j0 = c + 1 + g * L;
    int j1 = j0 + L; if (j1 > n) j1 = n;
    const int i = k0 + 1 + rb * BR_GV_NT + tid;
    const bool live = i < n;
    const cplx* vcol = A + (long long)c * n;
    cplx acc0 = mk(0, 0), acc1 = mk(0, 0), acc2 = mk(0, 0), acc3 = mk(0, 0);
    for (int jb = j0; jb < j1; jb += BR_GV_VS) {
        const int cnt = (j1 - jb < BR_GV_VS) ? j1 - jb : BR_GV_VS;
        __syncthreads();
        for (int t = tid; t < cnt; t += BR_GV_NT) s_v[t] = (jb + t == c + 1) ? mk(1.0, 0.0) : vcol[jb + t];
        __syncthreads();
        if (live) {
            const cplx* ap = A + (long long)jb * n + i;
            int t = 0;
            for (; t + 4 <= cnt; t += 4) {
                const cplx a0 = ap[(long long)t * n], a1 = ap[(long long)(t + 1) * n];
                const cplx a2 = ap[(long long)(t + 2) * n], a3 = ap[(long long)(t + 3) * n];
                acc0 = cfma(a0, s_v[t], acc0);
                acc1 = cfma(a1, s_v[t + 1], acc1);
                acc2 = cfma(a2, s_v[t + 2], acc2);
                acc3 = cfma(a3, s_v[t + 3], acc3);
            }
            for (; t < cnt; ++t) acc0 = cfma(ap[(long long)t * n], s_v[t], acc0);
        }
    }
    if (live) sc.part[(long long)g * n + i] = cadd(cadd(acc0, acc1), cadd(acc2, acc3));
}

// -----------------------------------------------------------------------------------------
// k_br_gemm: complex GEMM tiles on the FP64 tensor path (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA; a
// complex product is four real MMAs).  CTA tile 64 (m) x 32 (n), 8 warps x (one m8 row tile x four n8
// column tiles), K in chunks of 16 staged in padded shared memory with register prefetch of the next
// chunk.  The four uses differ in operand addressing and in the epilogue only.
// Two more uses serve the blocked BACK-TRANSFORMATION of the Hessenberg-basis eigenvectors X (psi, row
// major [row][mode]):  X <- (I - V T V^H) X per panel, panels in reverse order:
//     BXW   W = V^H X(k0+1:n, :)       (then W <- T W: k_bx_tw)
//     BXA   X(k0+1:n, :) -= V W
enum { BR_YTOP = 0, BR_RIGHT = 1, BR_LEFTW = 2, BR_LEFTA = 3, BR_BXW = 4, BR_BXA = 5 };
#define BR_TM 64
#define BR_TN 32
#define BR_KC 16
#define BR_SA (BR_TM + 2)
#define BR_SB (BR_TN + 2)
#define BR_NLA (BR_KC * BR_TM / 256)    // A-operand elements per thread and chunk (4)
#define BR_NLB (BR_KC * BR_TN / 256)    // B-operand elements per thread and chunk (2)

#ifndef AXS_HOST_DMMA                  /* the host emulation models the instruction itself (tests/host/shim/dmma.h) */
__device__ __forceinline__ void br_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

template <int MODE>
__global__ void __launch_bounds__(256)
k_br_gemm(int n, int k0, int nbc, int CG, AxsWorkBig wk, cplx* __restrict__ X_all)
{
    __shared__ __align__(16) cplx sA[BR_KC * BR_SA];
    __shared__ __align__(16) cplx sB[BR_KC * BR_SB];
    const int f = blockIdx.z, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    cplx* A = wk.A + (long long)f * n * n;
    const BrScratch sc = br_scratch(wk.yw + (long long)f * AXS_BIG_YW * n, n, CG);
    cplx* X = (MODE == BR_BXW || MODE == BR_BXA) ? X_all + (long long)f * n * n : nullptr;
    const int t0 = k0 + nbc;                                // first trailing column
    // logical index ranges [lo, hi) of m, n, k for this use
    int m_lo, m_hi, n_lo, n_hi, k_lo, k_hi;
    if (MODE == BR_YTOP)       { m_lo = 0;      m_hi = k0 + 1; n_lo = 0;  n_hi = nbc; k_lo = k0 + 1; k_hi = n; }
    else if (MODE == BR_RIGHT) { m_lo = 0;      m_hi = n;      n_lo = t0; n_hi = n;   k_lo = 0;      k_hi = nbc; }
    else if (MODE == BR_LEFTW) { m_lo = t0;     m_hi = n;      n_lo = 0;  n_hi = nbc; k_lo = k0 + 1; k_hi = n; }
    else if (MODE == BR_LEFTA) { m_lo = k0 + 1; m_hi = n;      n_lo = t0; n_hi = n;   k_lo = 0;      k_hi = nbc; }
    else if (MODE == BR_BXW)   { m_lo = 0;      m_hi = n;      n_lo = 0;  n_hi = nbc; k_lo = k0 + 1; k_hi = n; }
    else                       { m_lo = k0 + 1; m_hi = n;      n_lo = 0;  n_hi = n;   k_lo = 0;      k_hi = nbc; }
    const int m0 = m_lo + blockIdx.x * BR_TM, n0 = n_lo + blockIdx.y * BR_TN;
    if (m0 >= m_hi || n0 >= n_hi) return;

    // operand elements: A-operand (m, k), B-operand (k, nn); out-of-range -> 0
    auto opA = [&](int m, int k) -> cplx {
        if (m >= m_hi || k >= k_hi) return mk(0, 0);
        if (MODE == BR_YTOP) return A[(long long)k * n + m];                       // A(i, j), m = i, k = j
        if (MODE == BR_RIGHT) return sc.Y[(long long)k * n + m];                   // Y(i, q)
        if (MODE == BR_LEFTW) return A[(long long)m * n + k];                      // A(i, j), m = j, k = i
        if (MODE == BR_BXW) return X[(long long)k * n + m];                        // X(i, mode), m = mode, k = i
        return br_v(A, n, k0, m, k);                                               // V(i, q)  (LEFTA, BXA)
    };
    auto opB = [&](int k, int nn) -> cplx {
        if (k >= k_hi || nn >= n_hi) return mk(0, 0);
        if (MODE == BR_YTOP) return br_v(A, n, k0, k, nn);                         // V(j, q)
        if (MODE == BR_RIGHT) return cconj(br_v(A, n, k0, nn, k));                 // conj(V(j, q)), nn = j, k = q
        if (MODE == BR_LEFTW || MODE == BR_BXW) return cconj(br_v(A, n, k0, k, nn));   // conj(V(i, q)), k = i, nn = q
        return sc.W[(long long)k * n + nn];                                        // W(q, j)  (LEFTA, BXA: j = mode)
    };
    // thread -> staged element: the memory-contiguous index of each operand is the fastest one
    const bool a_kfast = (MODE == BR_LEFTW);
    const bool b_kfast = (MODE == BR_YTOP || MODE == BR_LEFTW || MODE == BR_BXW);
    cplx ra[BR_NLA], rb[BR_NLB];
    auto fetch = [&](int kc) {
#pragma unroll
        for (int e = 0; e < BR_NLA; ++e) {
            const int idx = tid + e * 256;
            const int kk = a_kfast ? (idx % BR_KC) : (idx / BR_TM), mm = a_kfast ? (idx / BR_KC) : (idx % BR_TM);
            ra[e] = opA(m0 + mm, kc + kk);
        }
#pragma unroll
        for (int e = 0; e < BR_NLB; ++e) {
            const int idx = tid + e * 256;
            const int kk = b_kfast ? (idx % BR_KC) : (idx / BR_TN), nn = b_kfast ? (idx / BR_KC) : (idx % BR_TN);
            rb[e] = opB(kc + kk, n0 + nn);
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int e = 0; e < BR_NLA; ++e) {
            const int idx = tid + e * 256;
            const int kk = a_kfast ? (idx % BR_KC) : (idx / BR_TM), mm = a_kfast ? (idx / BR_KC) : (idx % BR_TM);
            sA[kk * BR_SA + mm] = ra[e];
        }
#pragma unroll
        for (int e = 0; e < BR_NLB; ++e) {
            const int idx = tid + e * 256;
            const int kk = b_kfast ? (idx % BR_KC) : (idx / BR_TN), nn = b_kfast ? (idx / BR_KC) : (idx % BR_TN);
            sB[kk * BR_SB + nn] = rb[e];
        }
    };

    double cre[4][2], cim[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) { cre[t][0] = cre[t][1] = 0.0; cim[t][0] = cim[t][1] = 0.0; }
    fetch(k_lo);
    for (int kc = k_lo; kc < k_hi; kc += BR_KC) {
        stage();
        __syncthreads();
        if (kc + BR_KC < k_hi) fetch(kc + BR_KC);
#pragma unroll
        for (int ks = 0; ks < BR_KC; ks += 4) {
            const cplx af = sA[(ks + tq) * BR_SA + w * 8 + gq];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const cplx bf = sB[(ks + tq) * BR_SB + t * 8 + gq];
                br_dmma(cre[t][0], cre[t][1], af.x, bf.x);
                br_dmma(cre[t][0], cre[t][1], -af.y, bf.y);
                br_dmma(cim[t][0], cim[t][1], af.x, bf.y);
                br_dmma(cim[t][0], cim[t][1], af.y, bf.x);
            }
        }
        __syncthreads();
    }
    // epilogue: this lane holds C(m0 + 8 w + gq, n0 + 8 t + 2 tq + e)
    const int m = m0 + w * 8 + gq;
    if (m >= m_hi) return;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int nn = n0 + t * 8 + 2 * tq + e;
            if (nn >= n_hi) continue;
            const cplx v = mk(cre[t][e], cim[t][e]);
            if (MODE == BR_YTOP) sc.Y[(long long)nn * n + m] = v;                 // Y0(i, q) (x T in k_br_top)
            else if (MODE == BR_LEFTW || MODE == BR_BXW) sc.W[(long long)nn * n + m] = v;   // W0(q, j) / W0(q, mode)
            else if (MODE == BR_BXA) {                                            // X(i, mode) -= ...
                cplx* p = X + (long long)m * n + nn;
                *p = csub(*p, v);
            } else {                                                              // A(i, j) -= ...
                cplx* p = A + (long long)nn * n + m;
                *p = csub(*p, v);
            }
        }
}

// k_br_top: rows i = 0 .. k0 (above the panel): Y(i, :) <- Y0(i, :) T, then the right update of the
// panel's own columns A(i, k0+jj) -= sum_{q < jj} Y(i, q) conj(V(k0+jj, q)).  One thread per row.
__global__ void __launch_bounds__(256)
k_br_top(int n, int k0, int nbc, int CG, int panel, AxsWorkBig wk)
{
    __shared__ cplx s_T[BR_NB_MAX * BR_NB_MAX];
    __shared__ cplx s_Vc[BR_NB_MAX * BR_NB_MAX];            // conj(V(k0+jj, q))
    const int f = blockIdx.y, tid = threadIdx.x;
    cplx* A = wk.A + (long long)f * n * n;
    const BrScratch sc = br_scratch(wk.yw + (long long)f * AXS_BIG_YW * n, n, CG);
    for (int idx = tid; idx < BR_NB_MAX * BR_NB_MAX; idx += 256) {
        const int p = idx / BR_NB_MAX, q = idx - p * BR_NB_MAX;
        s_T[idx] = (p < nbc && q < nbc && p <= q) ? sc.T[idx] : mk(0, 0);
        s_Vc[idx] = (p < nbc && q < nbc) ? cconj(br_v(A, n, k0, k0 + p, q)) : mk(0, 0);   // p = jj
        if (blockIdx.x == 0 && panel >= 0) sc.Tall[(long long)panel * BR_NB_MAX * BR_NB_MAX + idx] = s_T[idx];   // for the back-transformation
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + tid;
    if (i > k0) return;
    cplx y[BR_NB_MAX];
#pragma unroll
    for (int q = 0; q < BR_NB_MAX; ++q) y[q] = (q < nbc) ? sc.Y[(long long)q * n + i] : mk(0, 0);
#pragma unroll
    for (int q = BR_NB_MAX - 1; q >= 0; --q) {              // in place, right to left: y_q = sum_{p <= q} y_p T(p, q)
        cplx s = mk(0, 0);
#pragma unroll
        for (int p = 0; p <= q; ++p) s = cfma(y[p], s_T[p * BR_NB_MAX + q], s);
        y[q] = s;
    }
#pragma unroll
    for (int q = 0; q < BR_NB_MAX; ++q)
        if (q < nbc) sc.Y[(long long)q * n + i] = y[q];
    for (int jj = 1; jj < nbc; ++jj) {
        cplx* p = A + (long long)(k0 + jj) * n + i;
        cplx a = *p;
#pragma unroll
        for (int q = 0; q < BR_NB_MAX; ++q)
            if (q < jj) a = cfms(y[q], s_Vc[jj * BR_NB_MAX + q], a);
        *p = a;
    }
}

// k_br_tw: W(:, j) <- T^H W(:, j) for the trailing columns j = t0 .. n-1.  One thread per column.
__global__ void __launch_bounds__(256)
k_br_tw(int n, int k0, int nbc, int CG, AxsWorkBig wk)
{
    __shared__ cplx s_T[BR_NB_MAX * BR_NB_MAX];
    const int f = blockIdx.y, tid = threadIdx.x;
    const BrScratch sc = br_scratch(wk.yw + (long long)f * AXS_BIG_YW * n, n, CG);
    for (int idx = tid; idx < BR_NB_MAX * BR_NB_MAX; idx += 256) {
        const int p = idx / BR_NB_MAX, q = idx - p * BR_NB_MAX;
        s_T[idx] = (p < nbc && q < nbc && p <= q) ? sc.T[idx] : mk(0, 0);
    }
    __syncthreads();
    const int jcol = k0 + nbc + blockIdx.x * 256 + tid;
    if (jcol >= n) return;
    cplx wv[BR_NB_MAX];
#pragma unroll
    for (int q = 0; q < BR_NB_MAX; ++q) wv[q] = (q < nbc) ? sc.W[(long long)q * n + jcol] : mk(0, 0);
#pragma unroll
    for (int q = BR_NB_MAX - 1; q >= 0; --q) {              // w'_q = sum_{p <= q} conj(T(p, q)) w_p
        cplx s = mk(0, 0);
#pragma unroll
        for (int p = 0; p <= q; ++p) s = cfmac(wv[p], s_T[p * BR_NB_MAX + q], s);
        wv[q] = s;
    }
#pragma unroll
    for (int q = 0; q < BR_NB_MAX; ++q)
        if (q < nbc) sc.W[(long long)q * n + jcol] = wv[q];
}

// k_bx_tw: W(:, mode) <- T W(:, mode) with the T factor of panel `panel` (blocked back-transformation).
__global__ void __launch_bounds__(256)
k_bx_tw(int n, int nbc, int CG, int panel, AxsWorkBig wk)
{
    __shared__ cplx s_T[BR_NB_MAX * BR_NB_MAX];
    const int f = blockIdx.y, tid = threadIdx.x;
    const BrScratch sc = br_scratch(wk.yw + (long long)f * AXS_BIG_YW * n, n, CG);
    for (int idx = tid; idx < BR_NB_MAX * BR_NB_MAX; idx += 256) s_T[idx] = sc.Tall[(long long)panel * BR_NB_MAX * BR_NB_MAX + idx];
    __syncthreads();
    const int mode = blockIdx.x * 256 + tid;
    if (mode >= n) return;
    cplx wv[BR_NB_MAX];
#pragma unroll
    for (int q = 0; q < BR_NB_MAX; ++q) wv[q] = (q < nbc) ? sc.W[(long long)q * n + mode] : mk(0, 0);
#pragma unroll
    for (int q = 0; q < BR_NB_MAX; ++q) {                   // in place, top to bottom: w'_q = sum_{p >= q} T(q, p) w_p
        cplx s = mk(0, 0);
#pragma unroll
        for (int p = q; p < BR_NB_MAX; ++p) s = cfma(s_T[q * BR_NB_MAX + p], wv[p], s);
        wv[q] = s;
    }
#pragma unroll
    for (int q = 0; q < BR_NB_MAX; ++q)
        if (q < nbc) sc.W[(long long)q * n + mode] = wv[q];
}

// k_bx_scale: undo the balancing, X(i, :) *= D(i)
__global__ void __launch_bounds__(256)
k_bx_scale(int n, cplx* __restrict__ X_all, AxsWorkBig wk)
{
    const int f = blockIdx.y;
    cplx* X = X_all + (long long)f * n * n;
    const double* dsc = wk.scale + (long long)f * n;
    const long long nn = (long long)n * n;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < nn; idx += (long long)gridDim.x * 256) {
        const int i = (int)(idx / n);
        X[idx] = cscale(X[idx], dsc[i]);
    }
}

// k_br_pack: packed copies of H (QR input Hc, inverse-iteration input Hm), contiguous subdiagonal for
// k_modes_col, max |H| (tau[n-1].x).  grid = (column blocks, nf); the maximum goes through an
// atomicMax on the bit pattern of the non-negative double.
__global__ void __launch_bounds__(256)
k_br_pack(int n, AxsWorkBig wk, unsigned long long* __restrict__ hmax_bits)
{
    __shared__ double s_m[8];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cplx* A = wk.A + (long long)f * n * n;
    cplx* Hc = wk.Hc + (long long)f * hcol_size(n);
    cplx* Hm = wk.Hm + (long long)f * hcol_size(n);
    double hm = 0.0;
    for (int c = blockIdx.x; c < n; c += gridDim.x) {
        const int rows = (c + 2 < n) ? c + 2 : n;
        const cplx* col = A + (long long)c * n;
        const int off = hcol_off(c);
        for (int r = tid; r < rows; r += 256) {
            const cplx a = col[r];
            Hc[off + r] = a;
            Hm[off + r] = a;
            hm = fmax(hm, cabs1(a));
        }
    }
    for (int o = 16; o > 0; o >>= 1) hm = fmax(hm, __shfl_xor_sync(0xffffffffu, hm, o));
    if (lane == 0) s_m[warp] = hm;
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < 8; ++q) hm = fmax(hm, s_m[q]);
        atomicMax(hmax_bits + f, (unsigned long long)__double_as_longlong(hm));
    }
}
// second pass (after k_br_pack of the same batch): subdiagonal + max |H| into their final places.  The
// scratch region yw is dead by now; its first n entries become the contiguous subdiagonal.
__global__ void __launch_bounds__(256)
k_br_pack_fin(int n, AxsWorkBig wk, const unsigned long long* __restrict__ hmax_bits)
{
    const int f = blockIdx.y;
    const cplx* A = wk.A + (long long)f * n * n;
    cplx* sub = wk.yw + (long long)f * AXS_BIG_YW * n;
    for (int c = blockIdx.x * 256 + threadIdx.x; c < n - 1; c += gridDim.x * 256) sub[c] = A[(long long)c * n + c + 1];
    if (blockIdx.x == 0 && threadIdx.x == 0)
        wk.tau[(long long)f * n + n - 1] = mk(__longlong_as_double((long long)hmax_bits[f]), 0.0);
}

// ------------------------------------------------------------------------------ launcher
