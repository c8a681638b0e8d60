// Kernels 2 + 3 of the SAFE sweep, shared-memory ("small", 2N <= 166) instance:
// one CTA per frequency point.  Together they replace scipy.linalg.eig (LAPACK zgeev,
// solver.py:136) and the B-normalisation (solver.py:142).
//
//   k_reduce   S(w) built straight from the banded operators (never round-trips HBM as a
//              separate batch), zgebal-style power-of-2 scaling (Gauss-Seidel sweeps on a
//              fp32 |a_ij|^2 table in shared memory), unblocked Householder reduction to
//              Hessenberg form with the two-sided update fused into 2 passes over the
//              L2-resident work matrix (one read pass for w = A v, y = A^H v, one
//              read-modify-write pass for the rank-2 update).
//   k_hqr      explicit single-shift (Wilkinson) QR on the packed Hessenberg matrix held in
//              shared memory: systolic column wavefront for the left rotations (one thread
//              per column keeps its running element in a register), row-parallel right
//              rotations; standard + Ahues-Tisseur deflation; exceptional shifts.
//   k_modes    eigenvectors by one step of inverse iteration on the Hessenberg matrix, done
//              as a bottom-up pivoted LU (column operations) fused with the back
//              substitution, so the per-mode state is O(n): 32 modes per warp, state in
//              shared memory; then back-transformation through the Householder reflectors,
//              un-balancing, B-normalisation psi = v / sqrt(2k u^T K6 u + u^T C u) and the
//              store in the reference psi_hat layout.
#include "common.cuh"
#include "aed.cuh"
// AXS_AED_BUILD: which early-deflation window routine the QR kernels carry: 0 = shared memory (default, the
// faster one on B200), 1 = registers + shuffles, 2 = both behind the tuning switch hqr_aed_reg.  1 and 2 are
// A/B builds (tests/run_ab_aed2.sh): a kernel that carries both routines runs EITHER of them 1.7 - 3.6 x
// slower than a kernel that carries one (measured), so the product build has exactly one.
#ifndef AXS_AED_BUILD
#define AXS_AED_BUILD 0
#endif
#include <stdlib.h>
#include <stdio.h>

// =========================================================================================
// k_reduce
// =========================================================================================
#define RED_THREADS 256
#define RED_WARPS (RED_THREADS / 32)
#define RED_ROWS 6                      // ceil(AXS_MAX_SMALL_N2 / 32)

__global__ void __launch_bounds__(RED_THREADS, 2)
k_reduce(int N, int hb, const cplx* __restrict__ K0s, const cplx* __restrict__ Cb,
         const cplx* __restrict__ Mb, const cplx* __restrict__ K1c, const cplx* __restrict__ K6d,
         const double* __restrict__ omega, const cplx* __restrict__ S_in, AxsWork wk, int prebal)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = 2 * N;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx* A = wk.A + (long long)f * n * n;

    // shared layout
    cplx* v = (cplx*)smem_raw;                       // [n]
    cplx* wv = v + n;                                // [n]
    cplx* y = wv + n;                                // [n]
    cplx* wpart = y + n;                             // [RED_WARPS][n]
    double* d = (double*)(wpart + RED_WARPS * n);    // [n]
    double* red = d + n;                             // [RED_WARPS * 2 + 8]
    float* d2 = (float*)(red + RED_WARPS * 2 + 8);   // [n] d^2
    float* di2 = d2 + n;                             // [n] 1/d^2
    float* Wn = di2 + n;                             // [n*n] |a_ij|^2 (fp32 is enough to pick powers of 2)

    // prebal: k_balance (eig_reduce.cu) has already left the balanced matrix in A and D in wk.scale
    if (!prebal) {
    // ---- phase A: build S(w) (or copy the given dense matrix) --------------------------
    const double w = omega ? omega[f] : 0.0;
    const cplx* Sf = S_in ? S_in + (long long)f * n * n : nullptr;
    for (int idx = tid; idx < n * n; idx += RED_THREADS) {
        int c = idx / n, r = idx - c * n;
        cplx val = Sf ? Sf[idx] : companion_entry(N, hb, K0s, Cb, Mb, K1c, K6d, w, r, c);
        A[idx] = val;
        Wn[idx] = (float)val.x * (float)val.x + (float)val.y * (float)val.y;
    }
    for (int i = tid; i < n; i += RED_THREADS) { d[i] = 1.0; d2[i] = 1.0f; di2[i] = 1.0f; }
    __syncthreads();

    // ---- phase B: balancing (scaling only), LAPACK zgebal loop, one warp ---------------
    if (warp == 0) {
        for (int sweep = 0; sweep < 40; ++sweep) {
            bool noconv = false;
            for (int i = 0; i < n; ++i) {
                float cs = 0.f, rs = 0.f;
                for (int kk = lane; kk < n; kk += 32) {
                    cs += Wn[kk + i * n] * di2[kk];
                    rs += Wn[i + kk * n] * d2[kk];
                }
                cs = warp_sumf(cs); rs = warp_sumf(rs);
                double c = sqrt((double)cs) * d[i];
                double r = sqrt((double)rs) / d[i];
                if (c == 0.0 || r == 0.0) continue;
                double g = r * 0.5, fsc = 1.0, s = c + r;
                int guard = 0;
                while (c < g && guard < 200) { fsc *= 2.0; c *= 2.0; r *= 0.5; g *= 0.5; ++guard; }
                g = c * 0.5;
                while (g >= r && guard < 400) { fsc *= 0.5; c *= 0.5; g *= 0.5; r *= 2.0; ++guard; }
                if (c + r >= 0.95 * s) continue;
                noconv = true;
                if (lane == 0) {
                    double dn = d[i] * fsc;
                    d[i] = dn;
                    d2[i] = (float)(dn * dn);
                    di2[i] = (float)(1.0 / (dn * dn));
                }
                __syncwarp();
            }
            if (!noconv) break;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += RED_THREADS) {
        int c = idx / n, r = idx - c * n;
        double sc = d[c] / d[r];                    // exact: powers of two
        cplx a = A[idx];
        A[idx] = mk(a.x * sc, a.y * sc);
    }
    for (int i = tid; i < n; i += RED_THREADS) wk.scale[(long long)f * n + i] = d[i];
    }
    __syncthreads();

    // ---- phase C: Householder reduction to Hessenberg form ------------------------------
    // One pass over the L2-resident trailing matrix per reflector: the rank-2 update of reflector
    // k is fused with the products w' = A v', y' = A^H v' of reflector k+1 (column k+1 is updated
    // first and the next reflector generated from it; same scheme as k_reduce_big).
    cplx* v2 = (cplx*)(Wn + (size_t)n * n);          // next reflector and its products
    cplx* w2 = v2 + n;
    cplx* y2 = w2 + n;

    // reflector from column k (rows k+1..n-1) into vd; returns tau (uniform).  zlarfg.
    auto gen_reflector = [&](int k, cplx* vd) -> cplx {
        cplx* colk = A + (long long)k * n;
        double part = 0.0;
        for (int i = k + 2 + tid; i < n; i += RED_THREADS) part += cabs2(colk[i]);
        part = warp_sum(part);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        double xnorm2 = 0.0;
        for (int q = 0; q < RED_WARPS; ++q) xnorm2 += red[q];
        const cplx alpha = colk[k + 1];
        cplx tau = mk(0, 0);
        double beta = alpha.x;
        cplx sc = mk(0, 0);
        const bool trivial = (xnorm2 == 0.0 && alpha.y == 0.0);
        if (!trivial) {
            beta = -copysign(sqrt(cabs2(alpha) + xnorm2), alpha.x);
            tau = mk((beta - alpha.x) / beta, -alpha.y / beta);
            sc = cdiv(mk(1.0, 0.0), mk(alpha.x - beta, alpha.y));
        }
        for (int i = tid; i < n; i += RED_THREADS) {
            cplx vi = mk(0, 0);
            if (i == k + 1) vi = mk(1.0, 0.0);
            else if (i > k + 1) {
                vi = cmul(colk[i], sc);
                colk[i] = vi;                                 // keep the reflector for the back-transformation
            }
            vd[i] = vi;
        }
        __syncthreads();                                      // every thread has read alpha
        if (tid == 0) {
            if (!trivial) colk[k + 1] = mk(beta, 0.0);
            wk.tau[(long long)f * n + k] = tau;
        }
        return tau;
    };

    cplx tau = gen_reflector(0, v);
    {
        // products of reflector 0: w = A[:, 1:] v and y_j = v^H A[:, j]
        cplx wacc[RED_ROWS];
#pragma unroll
        for (int q = 0; q < RED_ROWS; ++q) wacc[q] = mk(0, 0);
        for (int j = 1 + 2 * warp; j < n; j += 2 * RED_WARPS) {
            const bool two = (j + 1 < n);
            const cplx* col0 = A + (long long)j * n;
            const cplx* col1 = col0 + n;
            cplx e0[RED_ROWS], e1[RED_ROWS];
#pragma unroll
            for (int q = 0; q < RED_ROWS; ++q) {
                const int i = lane + 32 * q;
                e0[q] = (i < n) ? col0[i] : mk(0, 0);
                e1[q] = (i < n && two) ? col1[i] : mk(0, 0);
            }
            const cplx v0 = v[j], v1 = two ? v[j + 1] : mk(0, 0);
            cplx y0 = mk(0, 0), y1 = mk(0, 0);
#pragma unroll
            for (int q = 0; q < RED_ROWS; ++q) {
                const int i = lane + 32 * q;
                if (i < n) {
                    wacc[q] = cfma(e0[q], v0, cfma(e1[q], v1, wacc[q]));
                    const cplx vi = v[i];
                    y0 = cfmac(e0[q], vi, y0);
                    y1 = cfmac(e1[q], vi, y1);
                }
            }
            y0 = warp_csum(y0); y1 = warp_csum(y1);
            if (lane == 0) { y[j] = y0; if (two) y[j + 1] = y1; }
        }
#pragma unroll
        for (int q = 0; q < RED_ROWS; ++q) {
            const int i = lane + 32 * q;
            if (i < n) wpart[warp * n + i] = wacc[q];
        }
        __syncthreads();
        for (int i = tid; i < n; i += RED_THREADS) {
            cplx sacc = mk(0, 0);
            for (int q = 0; q < RED_WARPS; ++q) sacc = cadd(sacc, wpart[q * n + i]);
            wv[i] = sacc;
        }
        __syncthreads();
    }
    for (int k = 0; k < n - 2; ++k) {
        // alpha2 = v^H w
        cplx ap = mk(0, 0);
        for (int i = k + 1 + tid; i < n; i += RED_THREADS) ap = cfmac(wv[i], v[i], ap);
        ap = warp_csum(ap);
        if (lane == 0) { red[2 * warp] = ap.x; red[2 * warp + 1] = ap.y; }
        __syncthreads();
        cplx al = mk(0, 0);
        for (int q = 0; q < RED_WARPS; ++q) { al.x += red[2 * q]; al.y += red[2 * q + 1]; }
        const cplx tal = cmul(tau, al);
        const cplx ctau = cconj(tau);
        // column k+1 first: A[i,j] -= tau w_i conj(v_j) + conj(tau) v_i z_j,  z_j = y_j - tau alpha2 conj(v_j)
        {
            const int j = k + 1;
            cplx* col = A + (long long)j * n;
            const cplx cvj = cconj(v[j]);
            const cplx t = cmul(tau, cvj);
            const cplx z = cmul(ctau, csub(y[j], cmul(tal, cvj)));
            for (int i = tid; i < n; i += RED_THREADS) col[i] = cfms(v[i], z, cfms(wv[i], t, col[i]));
        }
        __syncthreads();
        const bool has_next = (k + 1 < n - 2);
        cplx tau2 = mk(0, 0);
        if (has_next) tau2 = gen_reflector(k + 1, v2);        // (barriers inside; uniform branch)
        // fused pass over the columns j >= k+2
        cplx wacc[RED_ROWS];
#pragma unroll
        for (int q = 0; q < RED_ROWS; ++q) wacc[q] = mk(0, 0);
        for (int j = k + 2 + 2 * warp; j < n; j += 2 * RED_WARPS) {
            const bool two = (j + 1 < n);
            cplx* col0 = A + (long long)j * n;
            cplx* col1 = col0 + n;
            cplx e0[RED_ROWS], e1[RED_ROWS];
#pragma unroll
            for (int q = 0; q < RED_ROWS; ++q) {
                const int i = lane + 32 * q;
                e0[q] = (i < n) ? col0[i] : mk(0, 0);
                e1[q] = (i < n && two) ? col1[i] : mk(0, 0);
            }
            const cplx cv0 = cconj(v[j]), cv1 = two ? cconj(v[j + 1]) : mk(0, 0);
            const cplx t0 = cmul(tau, cv0), t1 = cmul(tau, cv1);
            const cplx z0 = cmul(ctau, csub(y[j], cmul(tal, cv0)));
            const cplx z1 = two ? cmul(ctau, csub(y[j + 1], cmul(tal, cv1))) : mk(0, 0);
            const cplx n0 = has_next ? v2[j] : mk(0, 0), n1 = (has_next && two) ? v2[j + 1] : mk(0, 0);
            cplx y0 = mk(0, 0), y1 = mk(0, 0);
#pragma unroll
            for (int q = 0; q < RED_ROWS; ++q) {
                const int i = lane + 32 * q;
                if (i < n) {
                    const cplx wi = wv[i], vi = v[i];
                    const cplx a0 = cfms(vi, z0, cfms(wi, t0, e0[q]));
                    const cplx a1 = cfms(vi, z1, cfms(wi, t1, e1[q]));
                    col0[i] = a0;
                    if (two) col1[i] = a1;
                    if (has_next) {
                        const cplx vn = v2[i];
                        wacc[q] = cfma(a0, n0, cfma(a1, n1, wacc[q]));
                        y0 = cfmac(a0, vn, y0);
                        y1 = cfmac(a1, vn, y1);
                    }
                }
            }
            if (has_next) {
                y0 = warp_csum(y0); y1 = warp_csum(y1);
                if (lane == 0) { y2[j] = y0; if (two) y2[j + 1] = y1; }
            }
        }
        if (has_next) {
#pragma unroll
            for (int q = 0; q < RED_ROWS; ++q) {
                const int i = lane + 32 * q;
                if (i < n) wpart[warp * n + i] = wacc[q];
            }
            __syncthreads();
            for (int i = tid; i < n; i += RED_THREADS) {
                cplx sacc = mk(0, 0);
                for (int q = 0; q < RED_WARPS; ++q) sacc = cadd(sacc, wpart[q * n + i]);
                w2[i] = sacc;
            }
        }
        __syncthreads();
        { cplx* tp = v; v = v2; v2 = tp; tp = wv; wv = w2; w2 = tp; tp = y; y = y2; y2 = tp; }
        tau = tau2;
    }

    // ---- phase D: packed copies of H -----------------------------------------------------
    cplx* Hc = wk.Hc + (long long)f * hcol_size(n);
    cplx* Hr = wk.Hr + (long long)f * hrow_size(n);
    for (int idx = tid; idx < n * n; idx += RED_THREADS) {
        int c = idx / n, r = idx - c * n;
        if (r <= c + 1) {
            cplx a = A[idx];
            Hc[hcol_off(c) + r] = a;
            Hr[hrow_off(r, n) + c - hrow_first(r)] = a;
        }
    }
}

static size_t reduce_smem(int n) {
    return (size_t)(3 + RED_WARPS) * n * sizeof(cplx) + (size_t)n * sizeof(double)
           + (RED_WARPS * 2 + 8) * sizeof(double) + (size_t)2 * n * sizeof(float)
           + (size_t)n * n * sizeof(float) + (size_t)3 * n * sizeof(cplx) + 16;
}

#include "hqr_mb.cuh"                       // k_hqr_mb (multi-bulge QR, one CTA per matrix) + hqr_mb_smem

#include "modes.cuh"                        // k_modes (inverse iteration on the Hessenberg matrix) + modes_smem

#include "backnorm.cuh"                     // k_backnorm (per-reflector back-transformation + normalisation) + backnorm_smem


// =========================================================================================
// debug accessor: dense H and scale factors out of the workspace
// =========================================================================================
__global__ void k_unpack_H(int n, AxsWork wk, cplx* __restrict__ Hd)
{
    const int f = blockIdx.y;
    const cplx* Hc = wk.Hc + (long long)f * hcol_size(n);
    cplx* out = Hd + (long long)f * n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        int c = idx / n, r = idx - c * n;
        out[idx] = (r <= c + 1) ? Hc[hcol_off(c) + r] : mk(0, 0);
    }
}

// ------------------------------------------------------------------------------ launchers
extern "C" int axs_launch_reduce_oc(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                    const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                    const cplx* S_in, AxsWork wk, cudaStream_t st);

extern "C" int axs_launch_balance(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                  const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                  const cplx* S_in, AxsWork wk, cudaStream_t st);

extern "C" int axs_launch_reduce(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                 const cplx* K1c, const cplx* K6d, const double* omega, int nf,
                                 const cplx* S_in, AxsWork wk, cudaStream_t st)
{
    const int n = 2 * N;
    // 2N >= 96: trailing block on chip (eig_reduce.cu); smaller matrices stay L2-resident, 2 CTAs per SM
    const int rc = axs_launch_reduce_oc(N, hb, K0s, Cb, Mb, K1c, K6d, omega, nf, S_in, wk, st);
    if (rc != -1) return rc;
    // small matrices: the short-chain balancing kernel first (tuning reduce_prebal = 0: the fused one-warp loop)
    int prebal = 0;
    if (axs_tuning()->reduce_prebal) {
        const int rb = axs_launch_balance(N, hb, K0s, Cb, Mb, K1c, K6d, omega, nf, S_in, wk, st);
        if (rb) return rb;
        prebal = 1;
    }
    size_t smem = reduce_smem(n);
    AXS_SMEM_ONCE(smem, k_reduce);
    k_reduce<<<nf, RED_THREADS, smem, st>>>(N, hb, K0s, Cb, Mb, K1c, K6d, omega, S_in, wk, prebal);
    return (int)cudaGetLastError();
}

// launch one k_hqr_mb stage
// AED window of the stage with `per_sm` co-resident CTAs (1 -> stage 1, 2, 3, >= 4 -> the last instances):
// tuning hqr_aed, masked per stage by hqr_aed_stages (bit per_sm - 1)
static int aed_window_of_stage(int per_sm)
{
    int aed = axs_tuning()->hqr_aed;
    if (aed > AED_MAX) aed = AED_MAX;
    const int bit = (per_sm >= 4) ? 3 : per_sm - 1;
    if (aed < 2 || !((axs_tuning()->hqr_aed_stages >> bit) & 1)) return 0;
    return aed;
}

template <int NT, int MINB, int NITEMS, bool OPT>
static int launch_mb(int nf, size_t smem, cudaStream_t st, int n, int bmax, int first, int m_stop,
                     const cplx* Hc, cplx* Hp, cplx* eig, int* info, int* ucur, int* stats)
{
    int aed = aed_window_of_stage(MINB);
    if (aed >= 2) {
        if (axs_tuning()->hqr_aed_reg) aed |= 0x100;
        AXS_SMEM_ONCE(smem, k_hqr_mb<NT, MINB, NITEMS, OPT, true>);
        k_hqr_mb<NT, MINB, NITEMS, OPT, true><<<nf, NT, smem, st>>>(n, bmax, first, m_stop, aed, Hc, Hp, eig, info, ucur, stats);
    } else {
        AXS_SMEM_ONCE(smem, k_hqr_mb<NT, MINB, NITEMS, OPT, false>);
        k_hqr_mb<NT, MINB, NITEMS, OPT, false><<<nf, NT, smem, st>>>(n, bmax, first, m_stop, 0, Hc, Hp, eig, info, ucur, stats);
    }
    return (int)cudaGetLastError();
}

// Largest active window whose packed storage fits `per_sm` times into the shared memory of an SM
// (per CTA 2 KB: static shared memory of k_hqr_mb + the 1 KB the system reserves; 5 KB with the AED scratch).
extern "C" int axs_hqr_stage_capacity(int per_sm)
{
    const size_t slack = (aed_window_of_stage(per_sm) >= 2) ? 5120 : 2048;
    int m = 0;
    while (hqr_mb_smem(m + 1) + slack <= (size_t)(227 * 1024 / per_sm)) ++m;
    return m;
}

// The cascade below stage 1: starts with the instance whose occupancy fits the current window
// (512 threads x 2 CTAs/SM, 384 x 3, 256 x 4, 256 x 5, and 128 x 8 for windows <= 35) and hands the
// shrinking window down.  Hc_first != nullptr: the matrices are the un-parked packed Hessenberg
// matrices of size n (small problems that skip stage 1); otherwise the leading blocks parked in Hp
// (window <= m_cur, top index in ucur).  Also the tail of the large path.
static int hqr_cascade(int n, int nf, int m_cur, const cplx* Hc_first, cplx* Hp, cplx* eig, int* info, int* ucur,
                       int* stats, cudaStream_t st)
{
    const int stages = axs_tuning()->hqr_stages;       // 4 stages measured fastest on B200 (36.3 vs 37.7 ms)
    int bulges = axs_tuning()->hqr_bulges;
    if (bulges < 1 || bulges > MB_MAX) bulges = MB_MAX;
    const int m3 = (stages >= 3) ? axs_hqr_stage_capacity(3) : 0;
    const int m4 = (stages >= 4) ? axs_hqr_stage_capacity(4) : 0;
    const int m5 = (stages >= 5 || Hc_first) ? axs_hqr_stage_capacity(5) : 0;
    int first = Hc_first ? 1 : 0;
    const cplx* src = Hc_first ? Hc_first : Hp;
    if (m_cur <= 35 && bulges * (m_cur + 1) <= 3 * (128 - 32) && stages >= 4)
        return launch_mb<128, 8, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, 0, src, Hp, eig, info, ucur, stats);
    if (m3 == 0 || m_cur > m3) {
        const int stop = (m3 >= 8 && m3 < m_cur) ? m3 : 0;
        const int rc = launch_mb<512, 2, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (rc || !stop) return rc;
        first = 0; src = Hp; m_cur = stop;
    }
    if (m4 == 0 || m_cur > m4) {
        const int stop = (m4 >= 8 && m4 < m_cur) ? m4 : 0;
        const int rc = launch_mb<384, 3, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (rc || !stop) return rc;
        first = 0; src = Hp; m_cur = stop;
    }
    if (!(m5 >= 8 && m_cur <= m5)) {
        const int stop = (m5 >= 8 && m5 < m_cur) ? m5 : 0;
        const int rc = launch_mb<256, 4, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (rc || !stop) return rc;
        first = 0; src = Hp; m_cur = stop;
    }
    return launch_mb<256, 5, 3, false>(nf, hqr_mb_smem(m_cur), st, n, bulges, first, 0, src, Hp, eig, info, ucur, stats);
}

extern "C" int axs_launch_hqr_stages(int n, int nf, int m_first, cplx* Hp, cplx* eig, int* info, int* ucur,
                                     int* stats, cudaStream_t st)
{
    return hqr_cascade(n, nf, m_first, nullptr, Hp, eig, info, ucur, stats, st);
}

extern "C" int axs_launch_hqr(int n, int nf, cplx* eig, int* info, AxsWork wk, cudaStream_t st)
{
    int bulges = axs_tuning()->hqr_bulges;             // 8 concurrent bulges: fastest measured on B200
    if (bulges < 1 || bulges > MB_MAX) bulges = MB_MAX;
    // stage 1: whole matrix, 512 threads, one CTA per SM; stage 2: window fits twice into shared
    // memory (512 threads, 64 registers, 2 CTAs per SM); stage 3: fits three times (384 threads,
    // 56 registers, 3 CTAs per SM); ...  tuning hqr_stages = 1|2|3 limits the cascade (A/B testing).
    const int stages = axs_tuning()->hqr_stages;
    int m2 = (stages >= 2) ? axs_hqr_stage_capacity(2) : 0;
    if (m2 >= n || m2 < 8) m2 = 0;
    size_t smem = hqr_mb_smem(n);
    // problems that fit into the shared memory of an SM at least twice skip stage 1 and enter the
    // cascade at the instance that fits
    if (m2 == 0 && axs_tuning()->hqr_small_direct)
        return hqr_cascade(n, nf, n, wk.Hc, wk.Hp, eig, info, wk.ucur, wk.stats, st);
    int rc1;
    if (bulges * n <= 3 * 384)
        // leader sub-partition kept free in phase C (the 12 phase-C warps x 3 item slots must hold
        // bulges x n column-rotation items: 2N <= 144)
        rc1 = launch_mb<512, 1, 3, true>(nf, smem, st, n, bulges, 1, m2, wk.Hc, wk.Hp, eig, info, wk.ucur, wk.stats);
    else
        rc1 = launch_mb<512, 1, 3, false>(nf, smem, st, n, bulges, 1, m2, wk.Hc, wk.Hp, eig, info, wk.ucur, wk.stats);
    if (rc1 || m2 == 0) return rc1;
    return axs_launch_hqr_stages(n, nf, m2, wk.Hp, eig, info, wk.ucur, wk.stats, st);
}

extern "C" int axs_launch_backnorm_wy(int N, int hb, const cplx* Cb, const cplx* K6d, int nf,
                                      const cplx* eig, cplx* psi, AxsWork wk, cudaStream_t st);

extern "C" int axs_launch_modes(int N, int hb, const cplx* Cb, const cplx* K6d, int nf,
                                const cplx* eig, cplx* psi, AxsWork wk, cudaStream_t st)
{
    const int n = 2 * N;
    int nw = 4;
    while (nw > 1 && modes_smem(n, nw) > 227 * 1024) --nw;
    size_t smem = modes_smem(n, nw);
    AXS_SMEM_ONCE(smem, k_modes);
    k_modes<<<dim3((n + 8 * nw - 1) / (8 * nw), nf), 32 * nw, smem, st>>>(N, hb, Cb, K6d, eig, psi, wk);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    // 2N <= 128: blocked compact-WY back-transformation on DMMA tiles (eig_backx.cu); above that the
    // per-reflector kernel with a 64-mode slab in shared memory
    {
        const int rc = axs_launch_backnorm_wy(N, hb, Cb, K6d, nf, eig, psi, wk, st);
        if (rc != -1) return rc;
    }
    size_t smem2 = backnorm_smem(n);
    AXS_SMEM_ONCE(smem2, k_backnorm);
    k_backnorm<<<dim3((n + BN_COLS - 1) / BN_COLS, nf), 256, smem2, st>>>(N, hb, Cb, K6d, eig, psi, wk);
    return (int)cudaGetLastError();
}

extern "C" int axs_launch_unpack(int n, int nf, AxsWork wk, cplx* Hd, cudaStream_t st)
{
    k_unpack_H<<<dim3(32, nf), 256, 0, st>>>(n, wk, Hd);
    return (int)cudaGetLastError();
}
