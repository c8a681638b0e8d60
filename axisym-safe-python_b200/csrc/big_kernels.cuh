// The kernels of the global-memory ("large") instance of the eigen path (see eig_big.cu for the launchers and
// the overview): k_reduce_big, k_hqr_big, k_modes_col, k_backx_big, k_backx_slab, k_norm_big, k_unpack_big.  A
// header of its own so that tests/host/big_emul.cpp can compile this very source as host code and run the
// one-CTA-per-matrix instances on the emulated thread block of tests/host/cta_emul.h without a GPU (the
// thread-block-cluster instance of k_hqr_big needs distributed shared memory and stays GPU-only).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// =========================================================================================
// k_reduce_big
// =========================================================================================
#define RB_NT 512
#define RB_NW (RB_NT / 32)
#define RB_SLOTS 4                      // rows per lane in one row block (128 rows)

__device__ __forceinline__ double block_sum_d(double v, double* red, int warp, int lane) {
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < RB_NW; ++q) s += red[q];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(RB_NT, 1)
k_reduce_big(int N, int hb, const cplx* __restrict__ K0s, const cplx* __restrict__ Cb,
             const cplx* __restrict__ Mb, const cplx* __restrict__ K1c, const cplx* __restrict__ K6d,
             const double* __restrict__ omega, const cplx* __restrict__ S_in, AxsWorkBig wk, int balance_only)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    __shared__ double red[2 * RB_NW];
    __shared__ float redf[2 * RB_NW];
    const int n = 2 * N;
    const long long nn = (long long)n * n;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx* A = wk.A + (long long)f * nn;
    float* Wc = (float*)(wk.Hc + (long long)f * hcol_size(n));   // |a|^2, [c*n + r]
    float* Wr = Wc + nn;                                         // |a|^2, [r*n + c]

    // ---- phase A: build S(w) (or copy the given dense matrix), |a_ij|^2 tables ---------
    const double w = omega ? omega[f] : 0.0;
    const cplx* Sf = S_in ? S_in + (long long)f * nn : nullptr;
    for (long long idx = tid; idx < nn; idx += RB_NT) {
        const int c = (int)(idx / n), r = (int)(idx - (long long)c * n);
        const cplx val = Sf ? Sf[idx] : companion_entry(N, hb, K0s, Cb, Mb, K1c, K6d, w, r, c);
        A[idx] = val;
        const float m2 = (float)fmin(val.x * val.x + val.y * val.y, 1e37);
        Wc[idx] = m2;
        Wr[(long long)r * n + c] = m2;
    }
    // ---- phase B: balancing (scaling only), LAPACK zgebal loop ---------------------------
    double* d = (double*)smem_raw;          // [n]
    float* d2 = (float*)(d + n);            // [n] d^2
    float* di2 = d2 + n;                    // [n] 1/d^2
    for (int i = tid; i < n; i += RB_NT) { d[i] = 1.0; d2[i] = 1.0f; di2[i] = 1.0f; }
    __syncthreads();
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool noconv = false;
        for (int i = 0; i < n; ++i) {
            const float* wc = Wc + (long long)i * n;
            const float* wr = Wr + (long long)i * n;
            float cs = 0.f, rs = 0.f;
            for (int kk = tid; kk < n; kk += RB_NT) { cs += wc[kk] * di2[kk]; rs += wr[kk] * d2[kk]; }
            cs = warp_sumf(cs); rs = warp_sumf(rs);
            if (lane == 0) { redf[2 * warp] = cs; redf[2 * warp + 1] = rs; }
            __syncthreads();
            float cst = 0.f, rst = 0.f;
#pragma unroll
            for (int q = 0; q < RB_NW; ++q) { cst += redf[2 * q]; rst += redf[2 * q + 1]; }
            const double di = d[i];
            double c = sqrt((double)cst) * di, r = sqrt((double)rst) / di;
            double fsc = 1.0;
            bool upd = false;
            if (c != 0.0 && r != 0.0 && isfinite(c) && isfinite(r)) {
                double g = r * 0.5;
                const double s = c + r;
                int guard = 0;
                while (c < g && guard < 200) { fsc *= 2.0; c *= 2.0; r *= 0.5; g *= 0.5; ++guard; }
                g = c * 0.5;
                while (g >= r && guard < 400) { fsc *= 0.5; c *= 0.5; g *= 0.5; r *= 2.0; ++guard; }
                upd = (c + r < 0.95 * s);
            }
            __syncthreads();                               // everybody has read redf, d[i]
            if (upd) {                                     // uniform: all threads computed the same
                noconv = true;
                if (tid == 0) {
                    const double dn = di * fsc;
                    d[i] = dn;
                    d2[i] = (float)(dn * dn);
                    di2[i] = (float)(1.0 / (dn * dn));
                }
                __syncthreads();
            }
        }
        if (!noconv) break;
    }
    for (long long idx = tid; idx < nn; idx += RB_NT) {
        const int c = (int)(idx / n), r = (int)(idx - (long long)c * n);
        const double sc = d[c] / d[r];                     // exact: powers of two
        const cplx a = A[idx];
        A[idx] = mk(a.x * sc, a.y * sc);
    }
    for (int i = tid; i < n; i += RB_NT) wk.scale[(long long)f * n + i] = d[i];
    __syncthreads();
    if (balance_only) return;                              // the blocked reduction (eig_bigred.cu) takes over

    // ---- phase C: Householder reduction to Hessenberg form -------------------------------
    // One read + one write of the trailing matrix per reflector: the rank-2 update of reflector k
    //     A[i,j] -= tau w_i conj(v_j)  (all i)   and   -= conj(tau) v_i z_j  (i > k),
    //     z_j = y_j - tau (v^H w) conj(v_j),    w = A v,   y_j = v^H A[:, j]
    // is fused with the products w', y' of reflector k+1: column k+1 is updated first, reflector
    // k+1 is generated from it, and the pass over the columns j >= k+2 updates every entry and
    // immediately accumulates w' += a_ij v'_j and y'_j += conj(v'_i) a_ij.  The pass is a 2-D
    // (row block x column group) decomposition over the warps; the partial sums are combined in
    // a fixed order, so results are bit-reproducible.
    cplx* v = (cplx*)smem_raw;                             // [n] current reflector (reuses the balancing tables)
    cplx* ywb = wk.yw + (long long)f * AXS_BIG_YW * n;
    const int RBH = 32 * RB_SLOTS;                         // rows per row block
    const int nrb = (n + RBH - 1) / RBH;
    int CG = (48 + nrb - 1) / nrb;                         // column groups: >= 48 tasks for 16 warps
    if (CG > AXS_BIG_CG) CG = AXS_BIG_CG;
    const int ntask = nrb * CG;
    cplx* ypart = ywb + 4 * (long long)n;                  // [nrb][n]  per row block
    cplx* wpart = ypart + (long long)AXS_BIG_RB * n;       // [CG][n]   per column group

    // prologue: reflector 0 and its products y, w (plain pass over the whole matrix)
    cplx tau = mk(0, 0);
    {
        cplx* col0 = A;
        double part = 0.0;
        for (int i = 2 + tid; i < n; i += RB_NT) part += cabs2(col0[i]);
        const double xnorm2 = block_sum_d(part, red, warp, lane);
        const cplx alpha = col0[1];
        __syncthreads();
        if (!(xnorm2 == 0.0 && alpha.y == 0.0)) {
            const double beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
            tau = mk((beta - alpha.x) / beta, -alpha.y / beta);
            const cplx sc = cdiv(mk(1.0, 0.0), mk(alpha.x - beta, alpha.y));
            for (int i = 2 + tid; i < n; i += RB_NT) col0[i] = cmul(col0[i], sc);
            if (tid == 0) col0[1] = mk(beta, 0.0);
        } else {
            for (int i = 2 + tid; i < n; i += RB_NT) col0[i] = mk(0, 0);   // v = e_1, tau = 0
        }
        if (tid == 0) wk.tau[(long long)f * n] = tau;
        __syncthreads();
    }
    for (int k = 0; k < n - 2; ++k) {
        cplx* ycur = ywb + ((k & 1) ? 2 * (long long)n : 0);
        cplx* wcur = ycur + n;
        cplx* ynext = ywb + ((k & 1) ? 0 : 2 * (long long)n);
        cplx* wnext = ynext + n;
        cplx* colk = A + (long long)k * n;
        // current reflector into shared memory: v[k+1] = 1, v[i] = A[i,k] below
        for (int i = tid; i < n; i += RB_NT) v[i] = (i <= k) ? mk(0, 0) : (i == k + 1) ? mk(1.0, 0.0) : colk[i];
        __syncthreads();
        if (k == 0) {
            // products of reflector 0: w = A[:, 1:] v, y_j = v^H A[1:, j]
            for (int task = warp; task < ntask; task += RB_NW) {
                const int rb = task / CG, g = task - rb * CG;
                const int i0 = rb * RBH + lane;
                cplx vi[RB_SLOTS], wacc[RB_SLOTS];
#pragma unroll
                for (int q = 0; q < RB_SLOTS; ++q) {
                    const int i = i0 + 32 * q;
                    vi[q] = (i < n) ? v[i] : mk(0, 0);
                    wacc[q] = mk(0, 0);
                }
                for (int j = 1 + g; j < n; j += CG) {
                    const cplx* col = A + (long long)j * n;
                    const cplx vj = v[j];
                    cplx yp = mk(0, 0);
#pragma unroll
                    for (int q = 0; q < RB_SLOTS; ++q) {
                        const int i = i0 + 32 * q;
                        const cplx e = (i < n) ? col[i] : mk(0, 0);
                        wacc[q] = cfma(e, vj, wacc[q]);
                        yp = cfmac(e, vi[q], yp);
                    }
                    yp = warp_csum(yp);
                    if (lane == 0) ypart[(long long)rb * n + j] = yp;
                }
#pragma unroll
                for (int q = 0; q < RB_SLOTS; ++q) {
                    const int i = i0 + 32 * q;
                    if (i < n) wpart[(long long)g * n + i] = wacc[q];
                }
            }
            __syncthreads();
            for (int j = 1 + tid; j < n; j += RB_NT) {
                cplx sy = mk(0, 0);
                for (int rb = 0; rb < nrb; ++rb) sy = cadd(sy, __ldcg(&ypart[(long long)rb * n + j]));
                ycur[j] = sy;
            }
            for (int i = tid; i < n; i += RB_NT) {
                cplx sw = mk(0, 0);
                for (int g = 0; g < CG; ++g) sw = cadd(sw, __ldcg(&wpart[(long long)g * n + i]));
                wcur[i] = sw;
            }
            __syncthreads();
        }
        // alpha2 = v^H w
        cplx ap = mk(0, 0);
        for (int i = k + 1 + tid; i < n; i += RB_NT) ap = cfmac(__ldcg(&wcur[i]), v[i], ap);
        ap = warp_csum(ap);
        if (lane == 0) { red[2 * warp] = ap.x; red[2 * warp + 1] = ap.y; }
        __syncthreads();
        cplx al = mk(0, 0);
#pragma unroll
        for (int q = 0; q < RB_NW; ++q) { al.x += red[2 * q]; al.y += red[2 * q + 1]; }
        __syncthreads();
        const cplx tal = cmul(tau, al);
        const cplx ctau = cconj(tau);
        // ---- column k+1 first: update it, then generate reflector k+1 from it -----------------
        const bool has_next = (k + 1 < n - 2);
        cplx* coln = A + (long long)(k + 1) * n;
        cplx tau2 = mk(0, 0);
        {
            const cplx cvj = cconj(v[k + 1]);
            const cplx t = cmul(tau, cvj);
            const cplx z = cmul(ctau, csub(__ldcg(&ycur[k + 1]), cmul(tal, cvj)));
            double part = 0.0;
            for (int i = tid; i < n; i += RB_NT) {
                const cplx a = cfms(v[i], z, cfms(__ldcg(&wcur[i]), t, coln[i]));
                coln[i] = a;
                if (i >= k + 3) part += cabs2(a);
            }
            if (has_next) {
                const double xnorm2 = block_sum_d(part, red, warp, lane);   // (barriers: column is visible)
                const cplx alpha = coln[k + 2];
                __syncthreads();
                if (!(xnorm2 == 0.0 && alpha.y == 0.0)) {
                    const double beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
                    tau2 = mk((beta - alpha.x) / beta, -alpha.y / beta);
                    const cplx sc = cdiv(mk(1.0, 0.0), mk(alpha.x - beta, alpha.y));
                    for (int i = k + 3 + tid; i < n; i += RB_NT) coln[i] = cmul(coln[i], sc);
                    if (tid == 0) coln[k + 2] = mk(beta, 0.0);
                } else {
                    for (int i = k + 3 + tid; i < n; i += RB_NT) coln[i] = mk(0, 0);
                }
                if (tid == 0) wk.tau[(long long)f * n + k + 1] = tau2;
            }
            __syncthreads();
        }
        // ---- fused pass over the columns j >= k+2 ------------------------------------------------
        for (int task = warp; task < ntask; task += RB_NW) {
            const int rb = task / CG, g = task - rb * CG;
            const int i0 = rb * RBH + lane;
            const bool any_y = has_next && (rb * RBH + RBH - 1 > k + 1);
            cplx vi[RB_SLOTS], wi[RB_SLOTS], vi2[RB_SLOTS], wacc[RB_SLOTS];
#pragma unroll
            for (int q = 0; q < RB_SLOTS; ++q) {
                const int i = i0 + 32 * q;
                vi[q] = (i < n) ? v[i] : mk(0, 0);
                wi[q] = (i < n) ? __ldcg(&wcur[i]) : mk(0, 0);
                vi2[q] = (!has_next || i >= n || i < k + 2) ? mk(0, 0) : (i == k + 2) ? mk(1.0, 0.0) : coln[i];
                wacc[q] = mk(0, 0);
            }
            for (int j = k + 2 + g; j < n; j += CG) {
                cplx* col = A + (long long)j * n;
                const cplx cvj = cconj(v[j]);
                const cplx t = cmul(tau, cvj);
                const cplx z = cmul(ctau, csub(__ldcg(&ycur[j]), cmul(tal, cvj)));
                const cplx v2j = (j == k + 2) ? mk(1.0, 0.0) : coln[j];
                cplx e[RB_SLOTS];
#pragma unroll
                for (int q = 0; q < RB_SLOTS; ++q) {
                    const int i = i0 + 32 * q;
                    e[q] = (i < n) ? col[i] : mk(0, 0);
                }
                cplx yp = mk(0, 0);
#pragma unroll
                for (int q = 0; q < RB_SLOTS; ++q) {
                    const int i = i0 + 32 * q;
                    const cplx a = cfms(vi[q], z, cfms(wi[q], t, e[q]));
                    if (i < n) col[i] = a;
                    wacc[q] = cfma(a, v2j, wacc[q]);
                    yp = cfmac(a, vi2[q], yp);
                }
                if (any_y) {
                    yp = warp_csum(yp);
                    if (lane == 0) ypart[(long long)rb * n + j] = yp;
                }
            }
            if (has_next) {
#pragma unroll
                for (int q = 0; q < RB_SLOTS; ++q) {
                    const int i = i0 + 32 * q;
                    if (i < n) wpart[(long long)g * n + i] = wacc[q];
                }
            }
        }
        __syncthreads();
        if (has_next) {
            const int rb0 = (k + 2) / RBH;                 // first row block with rows > k+1
            for (int j = k + 2 + tid; j < n; j += RB_NT) {
                cplx sy = mk(0, 0);
                for (int rb = rb0; rb < nrb; ++rb) sy = cadd(sy, __ldcg(&ypart[(long long)rb * n + j]));
                ynext[j] = sy;
            }
            for (int i = tid; i < n; i += RB_NT) {
                cplx sw = mk(0, 0);
                for (int g = 0; g < CG; ++g) sw = cadd(sw, __ldcg(&wpart[(long long)g * n + i]));
                wnext[i] = sw;
            }
        }
        tau = tau2;
        __syncthreads();
    }
    cplx* yv = ywb;                                        // (scratch is free now)

    // ---- phase D: two packed copies of H (QR input, inverse-iteration input), max |H| -------
    cplx* Hc = wk.Hc + (long long)f * hcol_size(n);
    cplx* Hm = wk.Hm + (long long)f * hcol_size(n);
    double hm = 0.0;
    for (long long idx = tid; idx < nn; idx += RB_NT) {
        const int c = (int)(idx / n), r = (int)(idx - (long long)c * n);
        if (r <= c + 1) {
            const cplx a = A[idx];
            Hc[hcol_off(c) + r] = a;
            Hm[hcol_off(c) + r] = a;
            if (r == c + 1) yv[c] = a;                     // contiguous subdiagonal for k_modes_col
            hm = fmax(hm, cabs1(a));
        }
    }
    for (int o = 16; o > 0; o >>= 1) hm = fmax(hm, __shfl_xor_sync(0xffffffffu, hm, o));
    if (lane == 0) red[warp] = hm;
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < RB_NW; ++q) hm = fmax(hm, red[q]);
        wk.tau[(long long)f * n + n - 1] = mk(hm, 0.0);
    }
}

// =========================================================================================
// k_hqr_big
// =========================================================================================
#define HB_NT 256
#define HB_NB 8                         // bulges in a chain
#define HB_D 3                          // rows between bulges
#define HB_S 24                         // macro steps per window
#define HB_WS (HB_S + HB_D * HB_NB)     // 48: largest window
#define HB_WLD (HB_WS + 1)              // window leading dimension (column-major)
#define HB_TILE 128                     // columns (rows) per off-diagonal tile: 2 CTAs per SM, so the
                                        // latency-bound phase D of one matrix overlaps phase O of another
#define HB_TLD (HB_TILE + 1)
#define HB_WORKERS (HB_NT - 32)
#define HB_ITEMS 2                      // ceil(HB_NB * HB_WS / HB_WORKERS)

// A thread-block CLUSTER of CL CTAs (1, 2, 4 or 8; launch attribute) works on one matrix: rank 0 chases
// the bulge chain through the diagonal window (phase D, serial), then all CL CTAs share the tiles of the
// off-diagonal strips (phase O, the bulk of the flops for 2N in the thousands); the recorded rotations
// travel through distributed shared memory, the matrix through L2 with a cluster barrier on either side.
// With a batch that fills the GPU by itself the launcher uses CL = 1 (one CTA per matrix, as before).
// CLUSTER = false is the one-CTA-per-matrix instance (all cluster code folds away).  In the cluster
// instance every CTA reads matrix entries that other SMs have written, so the reads of the deflation scan
// and of the shifts go through L2 (__ldcg) and every cluster barrier is followed by a fence.
template <bool CLUSTER>
__global__ void __launch_bounds__(HB_NT, 2)
k_hqr_big(int n, int m_stop, cplx* __restrict__ H_all, cplx* __restrict__ eig_all,
          int* __restrict__ info, int* __restrict__ ucur)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    cplx* buf = (cplx*)smem_raw;                          // window [HB_WS][HB_WLD] / tile [HB_WS][HB_TLD]
    __shared__ cplx s_rec[HB_S][HB_NB][2];
    __shared__ cplx s_shift[HB_NB];
    __shared__ cplx s_ra[2][HB_NB], s_rb[2][HB_NB];
    __shared__ cplx s_rr[2][HB_NB];
    __shared__ int s_l;
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = CLUSTER ? (int)cluster.num_blocks() : 1, crank = CLUSTER ? (int)cluster.block_rank() : 0;
    const int f = CLUSTER ? blockIdx.x / CL : blockIdx.x, tid = threadIdx.x;
    const bool lead = CLUSTER ? (crank == 0) : true;
    cplx* Hg = H_all + (long long)f * hcol_size(n);
    cplx* eig = eig_all + (long long)f * n;
#define GH(i, j) Hg[hcol_off(j) + (i)]
#define GHR(i, j) (CLUSTER ? __ldcg(&Hg[hcol_off(j) + (i)]) : Hg[hcol_off(j) + (i)])   /* read that may see another SM's write */

    int u = n - 1, its = 0, failed = 0;
    while (u >= m_stop) {
        // ---- deflation scan --------------------------------------------------------------
        if (tid == 0) s_l = 0;
        __syncthreads();
        for (int k = 1 + tid; k <= u; k += HB_NT) {
            const double sub = cabs1(GHR(k, k - 1));
            bool neg = false;
            if (sub <= SMALL_D) neg = true;
            else {
                const cplx hkk = GHR(k, k), hk1k1 = GHR(k - 1, k - 1);
                double tst = cabs1(hk1k1) + cabs1(hkk);
                if (tst == 0.0) {
                    if (k - 2 >= 0) tst += cabs1(GHR(k - 1, k - 2));
                    if (k + 1 <= u) tst += cabs1(GHR(k + 1, k));
                }
                if (sub <= EPS_D * tst) {
                    const double up = cabs1(GHR(k - 1, k));
                    const double ab = fmax(sub, up), ba = fmin(sub, up);
                    const double dd = cabs1(csub(hk1k1, hkk));
                    const double aa = fmax(cabs1(hkk), dd), bb = fmin(cabs1(hkk), dd);
                    const double sm = aa + ab;
                    if (ba * (ab / sm) <= fmax(SMALL_D, EPS_D * (bb * (aa / sm)))) neg = true;
                }
            }
            if (neg) atomicMax(&s_l, k);
        }
        __syncthreads();
        const int l = s_l;
        // (every CTA of the cluster takes the same decisions from the same global data; only the lead writes.
        // A subdiagonal entry found negligible stays negligible for the others whether or not they already see the zero.)
        if (l > 0 && tid == 0 && lead) GH(l, l - 1) = mk(0, 0);
        if (l == u) {
            if (tid == 0 && lead) eig[u] = GHR(u, u);
            --u; its = 0;
            __syncthreads();
            continue;
        }
        ++its;
        if (its > 120) {
            if (tid == 0 && lead) eig[u] = GHR(u, u);
            ++failed; --u; its = 0;
            __syncthreads();
            continue;
        }
        const int m = u - l + 1;
        int nb = (m - 2) / HB_D + 1;
        if (nb > HB_NB) nb = HB_NB;
        // ---- shifts: eigenvalues of the trailing 2x2 diagonal blocks -------------------------
        if (tid < (nb + 1) / 2) {
            const int i = u - 2 * tid;
            const cplx a11 = GHR(i - 1, i - 1), a12 = GHR(i - 1, i), a21 = GHR(i, i - 1), a22 = GHR(i, i);
            const cplx hd = cscale(csub(a11, a22), 0.5);
            const cplx disc = csqrt_(cadd(cmul(hd, hd), cmul(a12, a21)));
            const cplx mid = cscale(cadd(a11, a22), 0.5);
            cplx e1 = cadd(mid, disc), e2 = csub(mid, disc);
            if (cabs1(csub(e2, a22)) < cabs1(csub(e1, a22))) { const cplx tmp = e1; e1 = e2; e2 = tmp; }
            if (its % 10 == 0) {                         // exceptional shifts against stagnation
                const double ex = 0.75 * cabs1(GHR(u, u - 1));
                e1 = cadd(e1, mk(ex, 0)); e2 = cadd(e2, mk(0, ex));
            }
            s_shift[2 * tid] = e1;
            if (2 * tid + 1 < HB_NB) s_shift[2 * tid + 1] = e2;
        }
        __syncthreads();
        if (tid == 0) {                                  // rotation that introduces bulge 0
            cplx a, b, r;
            make_rot_rc(csub(GHR(l, l), s_shift[0]), GHR(l + 1, l), a, b, r);
            s_ra[0][0] = a; s_rb[0][0] = b; s_rr[0][0] = r;
        }
        __syncthreads();
        const int nsteps = (m - 1) + (nb - 1) * HB_D;
        if (CLUSTER && CL > 1) {
            // every CTA of the cluster has taken its decisions (deflation scan, shifts) from the matrix as
            // the previous batch left it: only now may the lead start to rewrite the diagonal window
            cluster.sync();
        }

        for (int t0 = 0; t0 < nsteps; t0 += HB_S) {
            const int t1 = (t0 + HB_S < nsteps) ? t0 + HB_S : nsteps;
            int kmin = l + t0 - HB_D * (nb - 1); if (kmin < l) kmin = l;
            int kmax = l + t1 - 1; if (kmax > u - 1) kmax = u - 1;
            const int wlo = (kmin - 1 > l) ? kmin - 1 : l;
            const int whi = (kmax + 2 < u) ? kmax + 2 : u;
            const int ws = whi - wlo + 1;
            // ---- phase D: chase the chain through the diagonal window (lead CTA) ----------------
            if (lead) {
            for (int idx = tid; idx < ws * ws; idx += HB_NT) {
                const int j = idx / ws, i = idx - j * ws;
                buf[j * HB_WLD + i] = (i <= j + 1) ? GHR(wlo + i, wlo + j) : mk(0, 0);
            }
#define WH(i, j) buf[((j) - wlo) * HB_WLD + ((i) - wlo)]
            // fixed work items of this window: (bulge, column) for the row rotations,
            // (bulge, row) for the column rotations; the last warp owns the critical chains
            int it_b[HB_ITEMS], it_x[HB_ITEMS];
#pragma unroll
            for (int q = 0; q < HB_ITEMS; ++q) {
                const int item = (tid < HB_WORKERS) ? tid + q * HB_WORKERS : (1 << 30);
                it_b[q] = -1; it_x[q] = 0;
                if (item < nb * ws) { it_b[q] = item / ws; it_x[q] = wlo + (item - it_b[q] * ws); }
            }
            __syncthreads();
            for (int t = t0; t < t1; ++t) {
                const int cur = t & 1, nxt = cur ^ 1;
                // phase R: rows (k, k+1) of columns k-1 .. whi
#pragma unroll
                for (int q = 0; q < HB_ITEMS; ++q) {
                    const int b = it_b[q], col = it_x[q];
                    const int k = l + t - b * HB_D;
                    if (b >= 0 && k >= l && k <= u - 1 && col >= k - 1) {
                        if (col == k - 1) WH(k, col) = s_rr[cur][b];
                        else {
                            const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                            const cplx p = WH(k, col), qv = WH(k + 1, col);
                            cplx top, bot;
                            rot_rows<true>(ga, gb, p, qv, top, bot);
                            WH(k, col) = top;
                            WH(k + 1, col) = bot;
                        }
                    }
                }
                __syncthreads();
                // phase C: columns (k, k+1) of rows wlo .. k; leaders: row k+1, new bulge, next rotation
                if (tid >= HB_WORKERS) {
                    const int b = tid - HB_WORKERS;
                    const int k = l + t - b * HB_D;
                    if (b < nb && k >= l && k <= u - 1) {
                        const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                        s_rec[t - t0][b][0] = ga; s_rec[t - t0][b][1] = gb;
                        const cplx c0 = WH(k + 1, k), c1 = WH(k + 1, k + 1);
                        cplx n0, n1;
                        rot_cols<true>(ga, gb, c0, c1, n0, n1);
                        WH(k + 1, k) = n0;
                        WH(k + 1, k + 1) = n1;
                        if (k + 2 <= u) {
                            const cplx h = WH(k + 2, k + 1);
                            const cplx y = cmulc(h, gb);                    // new bulge H(k+2, k)
                            WH(k + 2, k + 1) = cscale(h, ga.x);
                            cplx a2, b2, r2;
                            make_rot_rc(n0, y, a2, b2, r2);
                            s_ra[nxt][b] = a2; s_rb[nxt][b] = b2; s_rr[nxt][b] = r2;
                        }
                    }
                    if (b == nb && (t + 1) % HB_D == 0 && (t + 1) / HB_D < nb) {   // next bulge starts
                        const int bn = (t + 1) / HB_D;
                        cplx a2, b2, r2;
                        make_rot_rc(csub(WH(l, l), s_shift[bn]), WH(l + 1, l), a2, b2, r2);
                        s_ra[nxt][bn] = a2; s_rb[nxt][bn] = b2; s_rr[nxt][bn] = r2;
                    }
                }
#pragma unroll
                for (int q = 0; q < HB_ITEMS; ++q) {
                    const int b = it_b[q], r = it_x[q];
                    const int k = l + t - b * HB_D;
                    if (b >= 0 && k >= l && k <= u - 1 && r <= k) {
                        const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                        const cplx c0 = WH(r, k), c1 = WH(r, k + 1);
                        cplx n0, n1;
                        rot_cols<true>(ga, gb, c0, c1, n0, n1);
                        WH(r, k) = n0;
                        WH(r, k + 1) = n1;
                    }
                }
                __syncthreads();
            }
            for (int idx = tid; idx < ws * ws; idx += HB_NT) {
                const int j = idx / ws, i = idx - j * ws;
                if (i <= j + 1) GH(wlo + i, wlo + j) = buf[j * HB_WLD + i];
            }
            __syncthreads();
            }                                              // lead
            const int nt = t1 - t0;
            if (CLUSTER && CL > 1) {
                // the window is back in global memory and the rotations are recorded: hand both to the cluster
                __threadfence();
                cluster.sync();
                __threadfence();
                if (!lead) {
                    const cplx* src = cluster.map_shared_rank(&s_rec[0][0][0], 0);
                    cplx* dst = &s_rec[0][0][0];
                    for (int idx = tid; idx < nt * HB_NB * 2; idx += HB_NT) dst[idx] = src[idx];
                    __syncthreads();
                }
            }
            // ---- phase O (right): rows wlo..whi of the columns right of the window ------------
            for (int c0 = whi + 1 + crank * HB_TILE; c0 <= u; c0 += CL * HB_TILE) {
                const int nc = (u - c0 + 1 < HB_TILE) ? u - c0 + 1 : HB_TILE;
                for (int idx = tid; idx < nc * ws; idx += HB_NT) {
                    const int cid = idx / ws, rl = idx - cid * ws;
                    buf[rl * HB_TLD + cid] = GHR(wlo + rl, c0 + cid);
                }
                __syncthreads();
                if (tid < nc) {
                    for (int b = 0; b < nb; ++b) {
                        int ta = HB_D * b - t0; if (ta < 0) ta = 0;
                        int tb = u - 1 - l - t0 + HB_D * b; if (tb > nt - 1) tb = nt - 1;
                        if (ta > tb) continue;
                        cplx* p = buf + (l + t0 + ta - HB_D * b - wlo) * HB_TLD + tid;
                        cplx carry = *p;
                        for (int tt = ta; tt <= tb; ++tt) {
                            const cplx ga = s_rec[tt][b][0], gb = s_rec[tt][b][1];
                            const cplx low = p[HB_TLD];
                            cplx top, bot;
                            rot_rows<true>(ga, gb, carry, low, top, bot);
                            *p = top;
                            carry = bot;
                            p += HB_TLD;
                        }
                        *p = carry;
                    }
                }
                __syncthreads();
                for (int idx = tid; idx < nc * ws; idx += HB_NT) {
                    const int cid = idx / ws, rl = idx - cid * ws;
                    GH(wlo + rl, c0 + cid) = buf[rl * HB_TLD + cid];
                }
                __syncthreads();
            }
            // ---- phase O (up): columns wlo..whi of the rows above the window -------------------
            for (int r0 = l + crank * HB_TILE; r0 < wlo; r0 += CL * HB_TILE) {
                const int nr = (wlo - r0 < HB_TILE) ? wlo - r0 : HB_TILE;
                for (int idx = tid; idx < nr * ws; idx += HB_NT) {
                    const int cl = idx / nr, rid = idx - cl * nr;
                    buf[cl * HB_TLD + rid] = GHR(r0 + rid, wlo + cl);
                }
                __syncthreads();
                if (tid < nr) {
                    for (int b = 0; b < nb; ++b) {
                        int ta = HB_D * b - t0; if (ta < 0) ta = 0;
                        int tb = u - 1 - l - t0 + HB_D * b; if (tb > nt - 1) tb = nt - 1;
                        if (ta > tb) continue;
                        cplx* p = buf + (l + t0 + ta - HB_D * b - wlo) * HB_TLD + tid;
                        cplx carry = *p;
                        for (int tt = ta; tt <= tb; ++tt) {
                            const cplx ga = s_rec[tt][b][0], gb = s_rec[tt][b][1];
                            const cplx right = p[HB_TLD];
                            cplx n0, n1;
                            rot_cols<true>(ga, gb, carry, right, n0, n1);
                            *p = n0;
                            carry = n1;
                            p += HB_TLD;
                        }
                        *p = carry;
                    }
                }
                __syncthreads();
                for (int idx = tid; idx < nr * ws; idx += HB_NT) {
                    const int cl = idx / nr, rid = idx - cl * nr;
                    GH(r0 + rid, wlo + cl) = buf[cl * HB_TLD + rid];
                }
                __syncthreads();
            }
            if (CLUSTER && CL > 1) {                       // strips of all CTAs are in global memory; the lead
                __threadfence();                           // may overwrite its rotation record
                cluster.sync();
                __threadfence();
            }
        }
#undef WH
    }
#undef GH
#undef GHR
    if (tid == 0 && lead) { info[f] = failed; ucur[f] = u; }
}

static size_t hqr_big_smem() { return (size_t)HB_WS * HB_TLD * sizeof(cplx); }

// =========================================================================================
// k_modes_col
// =========================================================================================
template <int NT, int R, int M>
__global__ void __launch_bounds__(NT, (NT == 256 && R < 8) ? 2 : 1)
k_modes_col(int n, int nf, const cplx* __restrict__ eig_all, cplx* __restrict__ psi_all, AxsWorkBig wk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    cplx* s_sub = (cplx*)smem_raw;                 // [n] s_sub[j] = H(j+1, j)
    __shared__ cplx s_mul[2][M], s_z[2][M];
    __shared__ int s_sw[2][M];
    __shared__ cplx s_pc[M], s_pa[M];              // carry / acc of the row that has just become final
    __shared__ double s_red[NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = (n + M - 1) / M;
    int f_loaded = -1;
    unsigned char* base = wk.msc + (long long)blockIdx.x * wk.msc_stride;
    cplx* mul_s = (cplx*)base;                      // [M][n]
    cplx* z_s = mul_s + (size_t)M * n;              // [M][n]
    cplx* x_s = z_s + (size_t)M * n;                // [M][n]
    unsigned char* sw_s = (unsigned char*)(x_s + (size_t)M * n);   // [M][n]

    for (int item = blockIdx.x; item < nf * groups; item += gridDim.x) {
        const int f = item / groups, mode0 = (item - f * groups) * M;
        const cplx* Hm = wk.Hm + (long long)f * hcol_size(n);
        const double tiny = fmax(EPS_D * wk.tau[(long long)f * n + n - 1].x, SMALL_D);
        if (f != f_loaded) {                           // uniform
            const cplx* sub = wk.yw + (long long)f * AXS_BIG_YW * n;
            for (int i = tid; i < n - 1; i += NT) s_sub[i] = sub[i];
            f_loaded = f;
            __syncthreads();
        }
        cplx lam[M];
#pragma unroll
        for (int m = 0; m < M; ++m) lam[m] = eig_all[(long long)f * n + ((mode0 + m < n) ? mode0 + m : n - 1)];

        // pivot of column jn = (row jn+1 is final): swap flag, multiplier, solution entry
        auto pivot = [&](int jrow, int m, cplx carry, cplx acc, int slot) {
            // jrow = the finished row; decides column jrow-1.  The subdiagonal H(jrow, jrow-1) comes
            // from shared memory so that no global load sits on the critical chain.
            cplx rkk = carry, a2 = mk(0, 0);
            bool sw = false;
            if (jrow > 0) {
                const cplx a = s_sub[jrow - 1];
                sw = cabs1(a) > cabs1(carry);
                a2 = sw ? carry : a;
                rkk = sw ? a : carry;
            }
            if (cabs1(rkk) < tiny) rkk = mk(tiny, 0.0);
            const double rinv = 1.0 / cabs2(rkk);
            const cplx inv = mk(rkk.x * rinv, -rkk.y * rinv);
            const cplx z = cmul(acc, inv);
            z_s[(size_t)m * n + jrow] = z;
            if (jrow > 0) {
                const cplx mu = cmul(a2, inv);
                mul_s[(size_t)m * n + jrow - 1] = mu;
                sw_s[(size_t)m * n + jrow - 1] = sw ? 1 : 0;
                s_mul[slot][m] = mu; s_z[slot][m] = z; s_sw[slot][m] = sw ? 1 : 0;
            }
        };

        cplx carry[R][M], acc[R][M];
        cplx hn[R];                                    // column j-1, loaded one step ahead (L2 latency)
        if (n >= 2) {
            const cplx* col = Hm + hcol_off(n - 2);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int k = tid + NT * q;
                hn[q] = (k <= n - 2) ? col[k] : mk(0, 0);
            }
        }
        {
            const cplx* col = Hm + hcol_off(n - 1);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int k = tid + NT * q;
                const cplx h = (k < n) ? col[k] : mk(0, 0);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    carry[q][m] = (k == n - 1) ? csub(h, lam[m]) : h;
                    acc[q][m] = mk(1.0, 0.0);
                    if (k == n - 1) { s_pc[m] = carry[q][m]; s_pa[m] = acc[q][m]; }
                }
            }
            if (tid == (n - 1) % NT) {
#pragma unroll
                for (int m = 0; m < M; ++m) pivot(n - 1, m, s_pc[m], s_pa[m], 0);
            }
        }
        __syncthreads();
        // one column step: all rows k <= j absorb column j (already in hc), the next column is
        // fetched into hx meanwhile, the owner of row j decides the pivot of column j-1
        auto step = [&](int j, int slot, cplx (&hc)[R], cplx (&hx)[R]) {
            if (j > 0) {
                const cplx* col = Hm + hcol_off(j - 1) + tid;
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (tid + NT * q <= j - 1) hx[q] = col[NT * q];
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const cplx mu = s_mul[slot][m], z = s_z[slot][m];
                if (s_sw[slot][m]) {                   // uniform branch: column swap
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int k = tid + NT * q;
                        if (k <= j) {
                            cplx a = hc[q];
                            if (k == j) a = csub(a, lam[m]);
                            acc[q][m] = cfms(a, z, acc[q][m]);
                            carry[q][m] = cfms(mu, a, carry[q][m]);
                            if (k == j) { s_pc[m] = carry[q][m]; s_pa[m] = acc[q][m]; }
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int k = tid + NT * q;
                        if (k <= j) {
                            cplx a = hc[q];
                            if (k == j) a = csub(a, lam[m]);
                            const cplx c2 = carry[q][m];
                            acc[q][m] = cfms(c2, z, acc[q][m]);
                            carry[q][m] = cfms(mu, c2, a);
                            if (k == j) { s_pc[m] = carry[q][m]; s_pa[m] = acc[q][m]; }
                        }
                    }
                }
            }
            if (tid == j % NT) {                       // this thread holds row j
#pragma unroll
                for (int m = 0; m < M; ++m) pivot(j, m, s_pc[m], s_pa[m], slot ^ 1);
            }
            __syncthreads();
        };
        cplx hb[R];
#pragma unroll
        for (int q = 0; q < R; ++q) hb[q] = mk(0, 0);
        int j = n - 2;
        for (; j >= 1; j -= 2) {                       // two steps per iteration: register ping-pong
            step(j, 0, hn, hb);
            step(j - 1, 1, hb, hn);
        }
        if (j == 0) step(0, 0, hn, hb);
        // x = W z: undo the column operations (forward), one thread per mode
        if (tid < M) {
            const int m = tid;
            const cplx* mu = mul_s + (size_t)m * n;
            const cplx* zz = z_s + (size_t)m * n;
            const unsigned char* sw = sw_s + (size_t)m * n;
            cplx* xo = x_s + (size_t)m * n;
            cplx xj = zz[0];
            for (int j0 = 0; j0 < n - 1; j0 += 8) {
                cplx mr[8], zr[8];
                unsigned char sr[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int jj = (j0 + e < n - 1) ? j0 + e : n - 2;
                    mr[e] = mu[jj]; zr[e] = zz[jj + 1]; sr[e] = sw[jj];
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (j0 + e < n - 1) {
                        const cplx xj1 = cfms(mr[e], xj, zr[e]);
                        if (sr[e]) xo[j0 + e] = xj1;
                        else { xo[j0 + e] = xj; xj = xj1; }
                    }
                }
            }
            xo[n - 1] = xj;
        }
        __syncthreads();
        // unit max-norm, hand over to k_backx_big through psi
        cplx* X = psi_all + (long long)f * n * n;
        for (int m = 0; m < M; ++m) {
            const cplx* xo = x_s + (size_t)m * n;
            double mx = 0.0;
            for (int i = tid; i < n; i += NT) mx = fmax(mx, cabs1(xo[i]));
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) s_red[warp] = mx;
            __syncthreads();
            mx = s_red[0];
            for (int q = 1; q < NT / 32; ++q) mx = fmax(mx, s_red[q]);
            __syncthreads();
            const double inv = mx > 0.0 ? 1.0 / mx : 1.0;
            if (mode0 + m < n)
                for (int i = tid; i < n; i += NT) X[(long long)i * n + mode0 + m] = cscale(xo[i], inv);
        }
        __syncthreads();
    }
}

// =========================================================================================
// k_backx_big: X <- D Q X, Q = H_0 H_1 ... H_{n-3}; T threads per mode keep 16 entries each.
// =========================================================================================
#define BX_RS 16
__global__ void __launch_bounds__(512, 1)
k_backx_big(int n, int T, cplx* __restrict__ psi_all, AxsWorkBig wk)
{
    __shared__ cplx s_part[2][16];
    const int NTb = blockDim.x;
    const int C = NTb / T;                              // modes per CTA
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mloc = tid / T, tt = tid - mloc * T;
    const int mode = blockIdx.x * C + mloc;
    const bool live = mode < n;
    const int wpm = T >> 5;                             // warps per mode
    const cplx* A = wk.A + (long long)f * n * n;
    const cplx* tau = wk.tau + (long long)f * n;
    const double* dsc = wk.scale + (long long)f * n;
    cplx* psi = psi_all + (long long)f * n * n;

    cplx x[BX_RS];
#pragma unroll
    for (int q = 0; q < BX_RS; ++q) {
        const int i = tt + T * q;
        x[q] = (live && i < n) ? psi[(long long)i * n + mode] : mk(0, 0);
    }
    int par = 0;
    cplx tnext = (n >= 3) ? tau[n - 3] : mk(0, 0);
    for (int kk = n - 3; kk >= 0; --kk) {
        const cplx t = tnext;
        if (kk > 0) tnext = tau[kk - 1];
        if (t.x == 0.0 && t.y == 0.0) continue;         // uniform
        const cplx* vcol = A + (long long)kk * n;
        const int qmin = (kk + 1) / T;
        cplx dot = mk(0, 0);
#pragma unroll
        for (int q = 0; q < BX_RS; ++q) {
            const int i = tt + T * q;
            if (q >= qmin && i > kk && i < n) {
                const cplx vi = (i == kk + 1) ? mk(1.0, 0.0) : vcol[i];
                dot = cfmac(x[q], vi, dot);             // conj(v_i) x_i
            }
        }
        dot = warp_csum(dot);
        if (wpm > 1) {
            if (lane == 0) s_part[par][warp] = dot;
            __syncthreads();
            dot = mk(0, 0);
            for (int ww = 0; ww < wpm; ++ww) dot = cadd(dot, s_part[par][mloc * wpm + ww]);
            par ^= 1;
        }
        const cplx td = cmul(t, dot);
#pragma unroll
        for (int q = 0; q < BX_RS; ++q) {
            const int i = tt + T * q;
            if (q >= qmin && i > kk && i < n) {
                const cplx vi = (i == kk + 1) ? mk(1.0, 0.0) : vcol[i];
                x[q] = cfms(td, vi, x[q]);
            }
        }
    }
    if (live) {
#pragma unroll
        for (int q = 0; q < BX_RS; ++q) {
            const int i = tt + T * q;
            if (i < n) psi[(long long)i * n + mode] = cscale(x[q], dsc[i]);
        }
    }
}

// k_backx_slab: the same back-transformation for sizes whose slab of C mode vectors fits the
// shared memory next to a ring of four reflector buffers: CTA = one frequency x C modes, 256 threads =
// C columns x (256 / C) row groups, reflectors prefetched with cp.async two steps ahead (L2
// latency), one block barrier per reflector.
#include <cuda_pipeline.h>
__global__ void __launch_bounds__(256, 1)
k_backx_slab(int n, int C, cplx* __restrict__ psi_all, AxsWorkBig wk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    cplx* Xs = (cplx*)smem_raw;                        // [n][C]
    cplx* vbuf = Xs + (size_t)n * C;                   // [4][n] ring: reflectors are fetched two steps ahead
    __shared__ cplx s_red[2][8][8];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col = tid & (C - 1), grp = tid / C, G = 256 / C;
    const int m0 = blockIdx.x * C;
    const cplx* A = wk.A + (long long)f * n * n;
    const cplx* tau = wk.tau + (long long)f * n;
    const double* dsc = wk.scale + (long long)f * n;
    cplx* psi = psi_all + (long long)f * n * n;

    for (int idx = tid; idx < n * C; idx += 256) {
        const int i = idx / C, c = idx - i * C;
        Xs[idx] = (m0 + c < n) ? psi[(long long)i * n + m0 + c] : mk(0, 0);
    }
    auto stage = [&](int kk) {                         // reflector kk: rows kk+2 .. n-1 of column kk
        cplx* dst = vbuf + (size_t)(kk & 3) * n;
        const cplx* src = A + (long long)kk * n;
        for (int i = kk + 2 + tid; i < n; i += 256) __pipeline_memcpy_async(dst + i, src + i, sizeof(cplx));
        if (tid == 0) __pipeline_memcpy_async(dst + kk + 1, tau + kk, sizeof(cplx));   // slot of v[kk+1] = 1
        __pipeline_commit();
    };
    if (n >= 3) stage(n - 3);
    if (n >= 4) stage(n - 4); else __pipeline_commit();
    __pipeline_wait_prior(1);
    __syncthreads();
    for (int kk = n - 3; kk >= 0; --kk) {
        const cplx* vv = vbuf + (size_t)(kk & 3) * n;
        if (kk > 1) stage(kk - 2); else __pipeline_commit();   // (empty group keeps the accounting uniform)
        const cplx t = vv[kk + 1];
        const bool skip = (t.x == 0.0 && t.y == 0.0);
        const int i0 = kk + 1 + ((grp - (kk + 1)) % G + G) % G;   // first row >= kk+1 owned by this group
        cplx part = mk(0, 0);
        if (!skip) {
            int i = i0;
            if (i == kk + 1) { part = Xs[i * C + col]; i += G; }
            cplx p1 = mk(0, 0);
            for (; i + G < n; i += 2 * G) {
                part = cfmac(Xs[i * C + col], vv[i], part);
                p1 = cfmac(Xs[(i + G) * C + col], vv[i + G], p1);
            }
            if (i < n) part = cfmac(Xs[i * C + col], vv[i], part);
            part = cadd(part, p1);
            for (int o = 16; o >= C; o >>= 1) {
                part.x += __shfl_xor_sync(0xffffffffu, part.x, o);
                part.y += __shfl_xor_sync(0xffffffffu, part.y, o);
            }
            if (lane < C) s_red[kk & 1][warp][lane] = part;
        }
        __pipeline_wait_prior(1);                          // reflector kk-1 has landed, kk-2 may be in flight
        __syncthreads();
        if (!skip) {
            cplx dot = s_red[kk & 1][0][col];
#pragma unroll
            for (int w = 1; w < 8; ++w) dot = cadd(dot, s_red[kk & 1][w][col]);
            const cplx td = cmul(t, dot);
            int i = i0;
            if (i == kk + 1) { Xs[i * C + col] = csub(Xs[i * C + col], td); i += G; }
            for (; i < n; i += G) Xs[i * C + col] = cfms(td, vv[i], Xs[i * C + col]);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < n * C; idx += 256) {
        const int i = idx / C, c = idx - i * C;
        if (m0 + c < n) psi[(long long)i * n + m0 + c] = cscale(Xs[idx], dsc[i]);
    }
}

// k_norm_big: CTA = one frequency x 32 modes, 256 threads = 32 modes x 8 row groups.
__global__ void __launch_bounds__(256)
k_norm_big(int N, int hb, const cplx* __restrict__ Cb, const cplx* __restrict__ K6d,
           const cplx* __restrict__ eig_all, cplx* __restrict__ psi_all)
{
    __shared__ cplx s_red[8][32];
    const int n = 2 * N, W = 2 * hb + 1;
    const int f = blockIdx.y, tid = threadIdx.x, col = tid & 31, grp = tid >> 5;
    const int mode = blockIdx.x * 32 + col;
    const bool live = mode < n;
    cplx* psi = psi_all + (long long)f * n * n;
    const cplx lam = eig_all[(long long)f * n + (live ? mode : n - 1)];
    const int mc = live ? mode : n - 1;
    cplx s6 = mk(0, 0), sc = mk(0, 0);
    for (int r = grp; r < N; r += 8) {
        const cplx ur = psi[(long long)(N + r) * n + mc];
        s6 = cfma(cmul(ur, ur), K6d[r], s6);
        cplx cu = mk(0, 0);
        const int c0 = max(0, r - hb), c1 = min(N - 1, r + hb);
        const cplx* crow = Cb + (long long)r * W - r + hb;
        for (int c = c0; c <= c1; ++c) cu = cfma(crow[c], psi[(long long)(N + c) * n + mc], cu);
        sc = cfma(ur, cu, sc);
    }
    s_red[grp][col] = cadd(cscale(cmul(lam, s6), 2.0), sc);
    __syncthreads();
    cplx nrm = s_red[0][col];
#pragma unroll
    for (int g = 1; g < 8; ++g) nrm = cadd(nrm, s_red[g][col]);
    const cplx inv = cdiv(mk(1.0, 0.0), csqrt_(nrm));
    __syncthreads();                                    // all reads of the un-normalised psi are done
    if (live) {
        for (int r = grp; r < N; r += 8) {
            const cplx u = cmul(psi[(long long)(N + r) * n + mode], inv);
            psi[(long long)(N + r) * n + mode] = u;
            psi[(long long)r * n + mode] = cmul(lam, u);
        }
    }
}

__global__ void k_unpack_big(int n, AxsWorkBig wk, cplx* __restrict__ Hd)
{
    const int f = blockIdx.y;
    const cplx* Hm = wk.Hm + (long long)f * hcol_size(n);
    cplx* out = Hd + (long long)f * n * n;
    const long long nn = (long long)n * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nn;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / n), r = (int)(idx - (long long)c * n);
        out[idx] = (r <= c + 1) ? Hm[hcol_off(c) + r] : mk(0, 0);
    }
}

// ------------------------------------------------------------------------------ launchers
