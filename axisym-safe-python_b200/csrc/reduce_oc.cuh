// k_balance + k_reduce_oc: S(w), zgebal scaling and the on-chip Householder Hessenberg reduction of the
// shared-memory eigen path (see eig_reduce.cu for the launchers and the overview).  A header of its own so
// that tests/host/sweep_emul.cpp can compile this very source as host code and run it on an emulated thread
// block (tests/host/cta_emul.h) without a GPU.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define BAL_THREADS 256
#define BAL_RQ 6                        // ceil(AXS_MAX_SMALL_N2 / 32) row slots per lane

// zgebal('S') loop semantics (Gauss-Seidel over the rows, power-of-2 factors, 2-norms including the
// diagonal), organised so that the sequential part is short:
//   per sweep   all threads compute the column / row sums  cs_i = sum_k |a_ki|^2 / d_k^2,
//               rs_i = sum_k |a_ik|^2 d_k^2  with the d of the sweep start (fp64 accumulation);
//   resolve     one warp walks i = 0..n-1: the owner lane broadcasts (cs_i, rs_i, d_i), every lane
//               takes the same decision in the SQUARED domain (no sqrt / division on the chain;
//               c < r/2  <=>  c^2 < r^2/4), and if d_i changes every lane corrects the sums of its
//               rows m > i by the one term that changed.  The final test (c'+r') >= 0.95 (c+r) is
//               only evaluated when the factor is not 1 and uses rho = r/c = r^2 rsqrt(r^2 c^2).
__global__ void __launch_bounds__(BAL_THREADS, 3)
k_balance(int N, int hb, const cplx* __restrict__ K0s, const cplx* __restrict__ Cb,
          const cplx* __restrict__ Mb, const cplx* __restrict__ K1c, const cplx* __restrict__ K6d,
          const double* __restrict__ omega, const cplx* __restrict__ S_in, AxsWork wk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    const int n = 2 * N;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx* A = wk.A + (long long)f * n * n;
    double* d = (double*)smem_raw;                   // [n]
    double* dinv = d + n;                            // [n] 1/d (exact: powers of two)
    double* csP = dinv + n;                          // [2][n] partial column sums
    double* rsP = csP + 2 * n;                       // [2][n] partial row sums
    float* Wn = (float*)(rsP + 2 * n);               // [n*n] |a_ij|^2 (fp32 is enough to pick powers of 2)
    __shared__ int s_noconv;

    const double w = omega ? omega[f] : 0.0;
    const cplx* Sf = S_in ? S_in + (long long)f * n * n : nullptr;
    for (int idx = tid; idx < n * n; idx += BAL_THREADS) {
        int c = idx / n, r = idx - c * n;
        cplx val = Sf ? Sf[idx] : companion_entry(N, hb, K0s, Cb, Mb, K1c, K6d, w, r, c);
        A[idx] = val;
        Wn[idx] = (float)val.x * (float)val.x + (float)val.y * (float)val.y;
    }
    for (int i = tid; i < n; i += BAL_THREADS) { d[i] = 1.0; dinv[i] = 1.0; }
    __syncthreads();
    const int half = (n + 1) >> 1;
    for (int sweep = 0; sweep < 40; ++sweep) {
        // ---- sums with the scaling of the sweep start (two threads per row) ------------------
        for (int item = tid; item < 2 * n; item += BAL_THREADS) {
            const int h = item >= n, i = item - h * n;
            const int k0 = h * half, k1 = h ? n : half;
            double cs = 0.0, rs = 0.0;
            for (int kk = k0; kk < k1; ++kk) {
                const double dk = d[kk], ik = dinv[kk];
                cs = fma((double)Wn[kk + i * n], ik * ik, cs);
                rs = fma((double)Wn[i + kk * n], dk * dk, rs);
            }
            csP[item] = cs; rsP[item] = rs;
        }
        __syncthreads();
        // ---- sequential resolve ---------------------------------------------------------------
        if (warp == 0) {
            double cs[BAL_RQ], rs[BAL_RQ], dl[BAL_RQ], di[BAL_RQ];
#pragma unroll
            for (int q = 0; q < BAL_RQ; ++q) {
                const int m = lane + 32 * q;
                const bool in = m < n;
                cs[q] = in ? csP[m] + csP[n + m] : 0.0;
                rs[q] = in ? rsP[m] + rsP[n + m] : 0.0;
                dl[q] = in ? d[m] : 1.0;
                di[q] = in ? dinv[m] : 1.0;
            }
            bool noconv = false;
#pragma unroll
            for (int q = 0; q < BAL_RQ; ++q) {
                if (32 * q < n) {
                    const int lmax = (n - 32 * q) < 32 ? (n - 32 * q) : 32;
                    for (int l = 0; l < lmax; ++l) {
                        const int i = 32 * q + l;
                        const double ci = __shfl_sync(0xffffffffu, cs[q], l);
                        const double ri = __shfl_sync(0xffffffffu, rs[q], l);
                        const double dd = __shfl_sync(0xffffffffu, dl[q], l);
                        const double id = __shfl_sync(0xffffffffu, di[q], l);
                        const double c2 = ci * (dd * dd), r2 = ri * (id * id);
                        if (c2 == 0.0 || r2 == 0.0) continue;
                        double fsc = 1.0, finv = 1.0, cn = c2, rn = r2;
                        int guard = 0;
                        while (cn < 0.25 * rn && guard < 200) { fsc *= 2.0; finv *= 0.5; cn *= 4.0; rn *= 0.25; ++guard; }
                        while (0.25 * cn >= rn && guard < 400) { fsc *= 0.5; finv *= 2.0; cn *= 0.25; rn *= 4.0; ++guard; }
                        if (fsc == 1.0) continue;
                        const double rho = r2 * rsqrt(r2 * c2);          // r / c
                        if (fsc + rho * finv >= 0.95 * (1.0 + rho)) continue;
                        noconv = true;
                        const double nd = dd * fsc, nid = id * finv;
                        const double dd2 = nd * nd - dd * dd, di2 = nid * nid - id * id;
                        if (lane == l) { dl[q] = nd; di[q] = nid; }
#pragma unroll
                        for (int qq = 0; qq < BAL_RQ; ++qq) {
                            const int m = lane + 32 * qq;
                            if (m > i && m < n) {
                                cs[qq] = fma((double)Wn[i + m * n], di2, cs[qq]);
                                rs[qq] = fma((double)Wn[m + i * n], dd2, rs[qq]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < BAL_RQ; ++q) {
                const int m = lane + 32 * q;
                if (m < n) { d[m] = dl[q]; dinv[m] = di[q]; }
            }
            if (lane == 0) s_noconv = noconv ? 1 : 0;
        }
        __syncthreads();
        if (!s_noconv) break;
    }
    for (int idx = tid; idx < n * n; idx += BAL_THREADS) {
        int c = idx / n, r = idx - c * n;
        double sc = d[c] * dinv[r];                 // exact: powers of two
        cplx a = A[idx];
        A[idx] = mk(a.x * sc, a.y * sc);
    }
    for (int i = tid; i < n; i += BAL_THREADS) wk.scale[(long long)f * n + i] = d[i];
}

static size_t balance_smem(int n) {
    return (size_t)6 * n * sizeof(double) + (size_t)n * n * sizeof(float) + 16;
}

// -----------------------------------------------------------------------------------------
// Fused pass body for one pair of trailing columns (col0 = column j, col1 = column j+1):
//   A[i,j] -= w_i t_j + v_i z_j                     (rank-2 update of the current reflector)
//   wacc_i += A'[i,j] v'_j ;  y'_j = v'^H A'[:,j]   (products of the next reflector)
// GUARD = false: columns are padded to 32 R rows (pad rows hold zeros and stay zero).
// stage != nullptr: column 0 of the pair is also copied there (next step's pivot column).
template <int R, bool GUARD>
__device__ __forceinline__ void oc_pair(int n, int lane, bool two, cplx* col0, cplx* col1, cplx* stage,
                                        const cplx (&wi)[R], const cplx (&vi)[R], const cplx (&vn)[R],
                                        cplx t0, cplx t1, cplx z0, cplx z1, cplx n0, cplx n1,
                                        cplx (&wacc)[R], cplx& y0, cplx& y1)
{
    cplx e0[R], e1[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int i = lane + 32 * q;
        if (GUARD) {
            e0[q] = (i < n) ? col0[i] : mk(0, 0);
            e1[q] = (i < n && two) ? col1[i] : mk(0, 0);
        } else { e0[q] = col0[i]; e1[q] = col1[i]; }
    }
    cplx ya[2] = {mk(0, 0), mk(0, 0)}, yb[2] = {mk(0, 0), mk(0, 0)};   // two chains per column sum
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int i = lane + 32 * q;
        const cplx a0 = cfms(vi[q], z0, cfms(wi[q], t0, e0[q]));
        const cplx a1 = cfms(vi[q], z1, cfms(wi[q], t1, e1[q]));
        if (GUARD) {
            if (i < n) { col0[i] = a0; if (two) col1[i] = a1; }
            if (stage) stage[i] = a0;                    // rows >= n carry zeros
        } else { col0[i] = a0; col1[i] = a1; }
        wacc[q] = cfma(a0, n0, cfma(a1, n1, wacc[q]));   // rows >= n carry zeros, no guard needed
        ya[q & 1] = cfmac(a0, vn[q], ya[q & 1]);
        yb[q & 1] = cfmac(a1, vn[q], yb[q & 1]);
    }
    y0 = cadd(ya[0], ya[1]); y1 = cadd(yb[0], yb[1]);
}

// software-pipelined form of oc_pair for padded on-chip columns: the loads of the next pair are
// issued before the arithmetic of the current one
template <int R>
__device__ __forceinline__ void oc_load(int lane, const cplx* col0, const cplx* col1, cplx (&e0)[R], cplx (&e1)[R]) {
#pragma unroll
    for (int q = 0; q < R; ++q) { e0[q] = col0[lane + 32 * q]; e1[q] = col1[lane + 32 * q]; }
}
template <int R>
__device__ __forceinline__ void oc_apply(int lane, cplx* col0, cplx* col1, const cplx (&e0)[R], const cplx (&e1)[R],
                                         const cplx (&wi)[R], const cplx (&vi)[R], const cplx (&vn)[R],
                                         cplx t0, cplx t1, cplx z0, cplx z1, cplx n0, cplx n1,
                                         cplx (&wacc)[R], cplx& y0, cplx& y1)
{
    cplx ya[2] = {mk(0, 0), mk(0, 0)}, yb[2] = {mk(0, 0), mk(0, 0)};
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int i = lane + 32 * q;
        const cplx a0 = cfms(vi[q], z0, cfms(wi[q], t0, e0[q]));
        const cplx a1 = cfms(vi[q], z1, cfms(wi[q], t1, e1[q]));
        col0[i] = a0; col1[i] = a1;
        wacc[q] = cfma(a0, n0, cfma(a1, n1, wacc[q]));
        ya[q & 1] = cfmac(a0, vn[q], ya[q & 1]);
        yb[q & 1] = cfmac(a1, vn[q], yb[q & 1]);
    }
    y0 = cadd(ya[0], ya[1]); y1 = cadd(yb[0], yb[1]);
}

// two independent complex sums reduced together (6 shuffle rounds instead of 10):
// on return lane 0 holds sum(a), lane 16 holds sum(b), both in `a`
__device__ __forceinline__ void warp_csum2(cplx& a, cplx b, int lane) {
    const bool hi = lane >= 16;
    cplx keep = hi ? b : a, give = hi ? a : b;
    keep.x += __shfl_xor_sync(0xffffffffu, give.x, 16);
    keep.y += __shfl_xor_sync(0xffffffffu, give.y, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        keep.x += __shfl_xor_sync(0xffffffffu, keep.x, o);
        keep.y += __shfl_xor_sync(0xffffffffu, keep.y, o);
    }
    a = keep;
}

// PAD: on-chip columns use the leading dimension 32 R (no row guards on chip); otherwise n.
// Per reflector: pass -> barrier -> column step part 1 -> barrier -> part 2 -> barrier.
template <int NWARP, int R, bool PAD, bool PIPE>
__global__ void __launch_bounds__(NWARP * 32, 1)
k_reduce_oc(int n, int jg, AxsWork wk)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    constexpr int NT = NWARP * 32;
    constexpr int LV = 32 * R;                       // vectors are always padded (zeros beyond n)
    const int ld = PAD ? LV : n;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx* A = wk.A + (long long)f * n * n;

    cplx* v = (cplx*)smem_raw;                       // [LV] current reflector
    cplx* v2 = v + LV;                               // [LV] next reflector
    cplx* wv = v2 + LV;                              // [LV] w = A v
    cplx* y = wv + LV;                               // [LV] y = A^H v
    cplx* tj = y + LV;                               // [LV] t_j = tau conj(v_j)
    cplx* zj = tj + LV;                              // [LV] z_j = conj(tau) (y_j - tau alpha conj(v_j))
    cplx* stage = zj + LV;                           // [LV] next pivot column while it lives in global memory
    cplx* dummy = stage + LV;                        // [LV] zero column: partner of an unpaired column
    cplx* wpart = dummy + LV;                        // [NWARP][LV]
    double* red = (double*)(wpart + NWARP * LV);     // [NWARP] norm partials
    double* red2 = red + NWARP;                      // [2*NWARP] alpha partials
    cplx* s_alpha = (cplx*)(red2 + 2 * NWARP);       // [2]: alpha, tau of the next reflector
    cplx* As = s_alpha + 2;                          // [(n - jg) * ld] trailing columns jg..n-1

    // ---- load the on-chip columns, zero the padding -----------------------------------------
    for (int idx = tid; idx < (n - jg) * ld; idx += NT) {
        const int cc = idx / ld, r = idx - cc * ld;
        As[idx] = (r < n) ? A[(long long)(jg + cc) * n + r] : mk(0, 0);
    }
    for (int i = tid; i < LV; i += NT) {
        v[i] = v2[i] = wv[i] = y[i] = tj[i] = zj[i] = dummy[i] = mk(0, 0);
        stage[i] = (i < n) ? A[i] : mk(0, 0);        // column 0
    }
    for (int i = tid; i < NWARP * LV; i += NT) wpart[i] = mk(0, 0);
    for (int i = tid; i < 2 * NWARP; i += NT) red2[i] = 0.0;
    __syncthreads();

    cplx tau = mk(0, 0);
    const int nrw = (n + 31) >> 5;                   // warps that own rows in the column steps
#ifdef OC_TIMING
    long long tc[4] = {0, 0, 0, 0}, t0c = clock64();
#define OC_MARK(i) { const long long t1c = clock64(); tc[i] += t1c - t0c; t0c = t1c; }
#else
#define OC_MARK(i)
#endif
    for (int k = -1; k <= n - 3; ++k) {
        const bool has_next = (k + 1 < n - 2);
        const int c = k + 1;
        // ---- column step, part 1: finish w and alpha of reflector k, update column c ----------
        cplx al = mk(0, 0);
#pragma unroll
        for (int q = 0; q < NWARP; ++q) { al.x += red2[2 * q]; al.y += red2[2 * q + 1]; }
        const cplx tal = cmul(tau, al);
        const cplx ctau = cconj(tau);
        cplx a_i = mk(0, 0);
        if (warp < nrw) {                            // tid < LV
            cplx wsum = mk(0, 0);
#pragma unroll
            for (int q = 0; q < NWARP; ++q) wsum = cadd(wsum, wpart[q * LV + tid]);
            wv[tid] = wsum;
            const cplx vt = v[tid];
            const cplx cvj = cconj(vt);
            tj[tid] = cmul(tau, cvj);                // per-column factors of this step's rank-2 update
            zj[tid] = cmul(ctau, csub(y[tid], cmul(tal, cvj)));
            const cplx old = (c < jg) ? stage[tid] : As[(c - jg) * ld + (PAD ? tid : (tid < n ? tid : 0))];
            const cplx zc = cmul(ctau, csub(y[c], tal));         // v[c] = 1
            a_i = cfms(vt, zc, cfms(wsum, tau, old));            // rows >= n: vt = wsum = 0
            if (!PAD && tid >= n) a_i = mk(0, 0);
            if (has_next) {
                double part = (tid < n && tid >= k + 3) ? cabs2(a_i) : 0.0;
                part = warp_sum(part);
                if (lane == 0) red[warp] = part;
                if (tid == k + 2) *s_alpha = a_i;
            }
        }
        __syncthreads();                                         // #1
        OC_MARK(0)
        // ---- part 2: reflector k+1 from the updated column -----------------------------------
        cplx tau2 = mk(0, 0);
        if (has_next) {
            if (warp < nrw) {
                double xnorm2 = 0.0;
                for (int q = 0; q < nrw; ++q) xnorm2 += red[q];
                const cplx alpha = *s_alpha;
                double beta = alpha.x;
                cplx sc = mk(0, 0);
                const bool trivial = (xnorm2 == 0.0 && alpha.y == 0.0);
                if (!trivial) {
                    beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
                    const double rb = 1.0 / beta;
                    tau2 = mk((beta - alpha.x) * rb, -alpha.y * rb);
                    const cplx dn = mk(alpha.x - beta, alpha.y);
                    const double rd = 1.0 / cabs2(dn);
                    sc = mk(dn.x * rd, -dn.y * rd);
                }
                if (tid < n) {
                    cplx vi = mk(0, 0), fin = a_i;
                    if (tid == k + 2) { vi = mk(1.0, 0.0); if (!trivial) fin = mk(beta, 0.0); }
                    else if (tid > k + 2) { vi = cmul(a_i, sc); fin = vi; }   // reflector kept below the subdiagonal
                    v2[tid] = vi;
                    A[(long long)c * n + tid] = fin;
                }
                if (tid == 0) { wk.tau[(long long)f * n + c] = tau2; s_alpha[1] = tau2; }
            }
        } else if (tid < n) {
            v2[tid] = mk(0, 0);                                  // last step: update only
            A[(long long)c * n + tid] = a_i;
        }
        __syncthreads();                                         // #2
        OC_MARK(1)
        if (has_next) tau2 = s_alpha[1];
        // ---- fused pass over the trailing columns j >= k+2 -----------------------------------
        const int j0 = k + 2;
        const int gend = jg < j0 ? j0 : jg;                      // global columns [j0, gend), on-chip [gend, n)
        const int ngp = (gend - j0 + 1) >> 1;                    // pairs in the global range
        cplx wi[R], vi[R], vn[R], wacc[R];
        {
            // a warp without a column pair in this step (late steps) skips the register hoists
            const int nsp0 = (n - gend + 1) >> 1;
            int wr0 = warp - (ngp % NWARP);
            if (wr0 < 0) wr0 += NWARP;
            const bool busy = (warp < ngp) || (wr0 < nsp0);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int i = lane + 32 * q;
                wi[q] = busy ? wv[i] : mk(0, 0); vi[q] = busy ? v[i] : mk(0, 0); vn[q] = busy ? v2[i] : mk(0, 0);
                wacc[q] = mk(0, 0);
            }
        }
        cplx alp = mk(0, 0);                                     // lanes 0 / 16: sum_j y'_j v'_j
        for (int p = warp; p < ngp; p += NWARP) {
            const int j = j0 + 2 * p;
            const bool two = (j + 1 < gend);
            const cplx t0 = tj[j], z0 = zj[j], n0 = v2[j];
            const cplx t1 = two ? tj[j + 1] : mk(0, 0), z1 = two ? zj[j + 1] : mk(0, 0), n1 = two ? v2[j + 1] : mk(0, 0);
            cplx y0, y1;
            cplx* col0 = A + (long long)j * n;
            oc_pair<R, true>(n, lane, two, col0, col0 + n, p == 0 ? stage : nullptr, wi, vi, vn,
                             t0, t1, z0, z1, n0, n1, wacc, y0, y1);
            warp_csum2(y0, y1, lane);
            if (lane == 0) { y[j] = y0; alp = cfma(y0, n0, alp); }
            if (lane == 16 && two) { y[j + 1] = y0; alp = cfma(y0, n1, alp); }
        }
        {
            const int nsp = (n - gend + 1) >> 1;                 // pairs in the on-chip range
            int wrot = warp - (ngp % NWARP);                     // warps without a global pair go first
            if (wrot < 0) wrot += NWARP;
            if (PAD && PIPE) {
                cplx e0[R], e1[R], f0[R], f1[R];
                int p = wrot;
                if (p < nsp) {
                    const int j = gend + 2 * p;
                    cplx* c0p = As + (j - jg) * ld;
                    oc_load<R>(lane, c0p, (j + 1 < n) ? c0p + ld : dummy, e0, e1);
                }
                while (p < nsp) {
                    const int j = gend + 2 * p;
                    const bool two = (j + 1 < n);
                    const int j1 = two ? j + 1 : j;
                    const int pn = p + NWARP;
                    if (pn < nsp) {                              // next pair's loads go out first
                        const int jn = gend + 2 * pn;
                        cplx* c0n = As + (jn - jg) * ld;
                        oc_load<R>(lane, c0n, (jn + 1 < n) ? c0n + ld : dummy, f0, f1);
                    }
                    const cplx t0 = tj[j], z0 = zj[j], n0 = v2[j];
                    cplx t1 = tj[j1], z1 = zj[j1], n1 = v2[j1];
                    cplx* col0 = As + (j - jg) * ld;
                    cplx* col1 = col0 + ld;
                    if (!two) { t1 = z1 = n1 = mk(0, 0); col1 = dummy; }
                    cplx y0, y1;
                    oc_apply<R>(lane, col0, col1, e0, e1, wi, vi, vn, t0, t1, z0, z1, n0, n1, wacc, y0, y1);
                    warp_csum2(y0, y1, lane);
                    if (lane == 0) { y[j] = y0; alp = cfma(y0, n0, alp); }
                    if (lane == 16 && two) { y[j + 1] = y0; alp = cfma(y0, n1, alp); }
#pragma unroll
                    for (int q = 0; q < R; ++q) { e0[q] = f0[q]; e1[q] = f1[q]; }
                    p = pn;
                }
            } else
            for (int p = wrot; p < nsp; p += NWARP) {
                const int j = gend + 2 * p;
                const bool two = (j + 1 < n);
                const int j1 = two ? j + 1 : j;
                const cplx t0 = tj[j], z0 = zj[j], n0 = v2[j];
                cplx t1 = tj[j1], z1 = zj[j1], n1 = v2[j1];
                cplx* col0 = As + (j - jg) * ld;
                cplx* col1 = col0 + ld;
                if (!two) { t1 = z1 = n1 = mk(0, 0); col1 = dummy; }
                cplx y0, y1;
                oc_pair<R, !PAD>(n, lane, true, col0, col1, nullptr, wi, vi, vn, t0, t1, z0, z1, n0, n1, wacc, y0, y1);
                warp_csum2(y0, y1, lane);
                if (lane == 0) { y[j] = y0; alp = cfma(y0, n0, alp); }
                if (lane == 16 && two) { y[j + 1] = y0; alp = cfma(y0, n1, alp); }
            }
        }
        OC_MARK(2)
#pragma unroll
        for (int q = 0; q < R; ++q) wpart[warp * LV + lane + 32 * q] = wacc[q];
        {
            const double ox = __shfl_sync(0xffffffffu, alp.x, 16), oy = __shfl_sync(0xffffffffu, alp.y, 16);
            if (lane == 0) { red2[2 * warp] = alp.x + ox; red2[2 * warp + 1] = alp.y + oy; }
        }
        __syncthreads();                                         // #3
        OC_MARK(3)
        { cplx* tp = v; v = v2; v2 = tp; }
        tau = tau2;
    }
#ifdef OC_TIMING
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == NWARP - 1))
        printf("oc timing warp %d: part1 %lld part2 %lld pass %lld wait3 %lld cycles\n", warp, tc[0], tc[1], tc[2], tc[3]);
#endif
    // the last column never passes through the column step
    if (n - 1 >= jg) for (int i = tid; i < n; i += NT) A[(long long)(n - 1) * n + i] = As[(n - 1 - jg) * ld + i];
    __syncthreads();

    // ---- packed copies of H ---------------------------------------------------------------------
    cplx* Hc = wk.Hc + (long long)f * hcol_size(n);
    cplx* Hr = wk.Hr + (long long)f * hrow_size(n);
    for (int idx = tid; idx < n * n; idx += NT) {
        int c = idx / n, r = idx - c * n;
        if (r <= c + 1) {
            cplx a = A[idx];
            Hc[hcol_off(c) + r] = a;
            Hr[hrow_off(r, n) + c - hrow_first(r)] = a;
        }
    }
}

template <int NWARP, int R>
static size_t reduce_oc_fixed() {
    return (size_t)(8 + NWARP) * 32 * R * sizeof(cplx) + (size_t)(3 * NWARP) * sizeof(double) + 2 * sizeof(cplx);
}

