// k_reduce: S(w), zgebal scaling (or pre-balanced input) and the L2-resident Householder Hessenberg reduction
// of the shared-memory eigen path for 2N < 96 (see eig_small.cu).  A header of its own so that
// tests/host/sweep_emul.cpp can compile this very source as host code and run it on an emulated thread block
// (tests/host/cta_emul.h) without a GPU.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
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
    AXS_DYNAMIC_SMEM(smem_raw);
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
