// Shared helpers for the axisafe_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

typedef double2 cplx;   // (x = re, y = im); 16-byte aligned -> LDS.128 / LDG.128

#define AXS_HD __host__ __device__ __forceinline__

AXS_HD cplx mk(double re, double im) { return make_double2(re, im); }
AXS_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
AXS_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
AXS_HD cplx cneg(cplx a) { return mk(-a.x, -a.y); }
AXS_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
AXS_HD cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
AXS_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
AXS_HD cplx cmulc(cplx a, cplx b) { return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// acc + a*b
AXS_HD cplx cfma(cplx a, cplx b, cplx acc) {
    return mk(fma(a.x, b.x, fma(-a.y, b.y, acc.x)), fma(a.x, b.y, fma(a.y, b.x, acc.y)));
}
// acc - a*b
AXS_HD cplx cfms(cplx a, cplx b, cplx acc) {
    return mk(fma(-a.x, b.x, fma(a.y, b.y, acc.x)), fma(-a.x, b.y, fma(-a.y, b.x, acc.y)));
}
// acc + a*conj(b)
AXS_HD cplx cfmac(cplx a, cplx b, cplx acc) {
    return mk(fma(a.x, b.x, fma(a.y, b.y, acc.x)), fma(a.y, b.x, fma(-a.x, b.y, acc.y)));
}
AXS_HD double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }
AXS_HD double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
AXS_HD double cabsd(cplx a) { return hypot(a.x, a.y); }
AXS_HD cplx cdiv(cplx a, cplx b) {
    // Smith's algorithm (robust against overflow of |b|^2)
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return mk((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return mk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
AXS_HD cplx csqrt_(cplx z) {
    // principal square root
    double m = hypot(z.x, z.y);
    if (m == 0.0) return mk(0.0, 0.0);
    double t = sqrt(0.5 * (m + fabs(z.x)));
    if (z.x >= 0.0) return mk(t, 0.5 * z.y / t);
    return mk(0.5 * fabs(z.y) / t, z.y >= 0.0 ? t : -t);
}

// ---- packed storage of an n x n upper-Hessenberg matrix --------------------------------
// column-major: column j holds rows 0..min(j+1, n-1) contiguously.
AXS_HD int hcol_off(int j) { return (j * (j + 3)) >> 1; }
AXS_HD int hcol_size(int n) { return ((n * (n + 3)) >> 1) - 1; }
// row-major: row 0 holds columns 0..n-1, row i>=1 holds columns i-1..n-1 contiguously.
AXS_HD int hrow_off(int i, int n) { return i == 0 ? 0 : n + (i - 1) * (n + 1) - (((i - 1) * i) >> 1); }
AXS_HD int hrow_first(int i) { return i == 0 ? 0 : i - 1; }
AXS_HD int hrow_size(int n) { return hrow_off(n - 1, n) + 2; }

#define AXS_CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return (int)e_; } while (0)

// limits of the shared-memory ("small") eigen path
#define AXS_MAX_SMALL_N2 166

// workspace layout for nf problems of size n2 (all offsets 256-B aligned)
struct AxsWork {
    cplx*   A;      // [nf][n2*n2]  dense column-major work matrix (S -> reflectors below subdiag)
    cplx*   Hc;     // [nf][hcol_size] packed column-major Hessenberg (input of the QR kernel)
    cplx*   Hr;     // [nf][hrow_size] packed row-major Hessenberg (input of inverse iteration)
    cplx*   tau;    // [nf][n2]
    double* scale;  // [nf][n2]  balancing factors D
    int*    ucur;   // [nf]      active-window top between QR stages
    cplx*   Hp;     // [nf][hcol_size] leading block parked between QR stages (Hc stays intact)
};

static inline long long axs_align(long long x) { return (x + 255) & ~255LL; }

static inline long long axs_work_layout(int n2, int nf, void* base, AxsWork* w) {
    long long off = 0;
    char* b = (char*)base;
    long long szA = axs_align((long long)nf * n2 * n2 * sizeof(cplx));
    long long szHc = axs_align((long long)nf * hcol_size(n2) * sizeof(cplx));
    long long szHr = axs_align((long long)nf * hrow_size(n2) * sizeof(cplx));
    long long szT = axs_align((long long)nf * n2 * sizeof(cplx));
    long long szS = axs_align((long long)nf * n2 * sizeof(double));
    if (w) {
        w->A = (cplx*)(b + off);
    }
    off += szA;
    if (w) w->Hc = (cplx*)(b + off);
    off += szHc;
    if (w) w->Hr = (cplx*)(b + off);
    off += szHr;
    if (w) w->tau = (cplx*)(b + off);
    off += szT;
    if (w) w->scale = (double*)(b + off);
    off += szS;
    if (w) w->ucur = (int*)(b + off);
    off += axs_align((long long)nf * sizeof(int));
    if (w) w->Hp = (cplx*)(b + off);
    off += szHc;
    return off;
}
