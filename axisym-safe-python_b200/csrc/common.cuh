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

// ---- device helpers shared by eig_small.cu / eig_big.cu ----------------------------------
#ifdef __CUDACC__
#define EPS_D 2.220446049250313e-16
#define SMALL_D (2.2250738585072014e-308 / EPS_D)

__device__ __forceinline__ cplx band_get(const cplx* __restrict__ B, int W, int hb, int r, int c) {
    int d = c - r;
    return (d < -hb || d > hb) ? mk(0, 0) : B[(long long)r * W + d + hb];
}

__device__ __forceinline__ cplx companion_entry(int N, int hb, const cplx* K0s, const cplx* Cb,
                                                const cplx* Mb, const cplx* K1c, const cplx* K6d,
                                                double w, int r, int c)
{
    const int W = 2 * hb + 1;
    if (r >= N) return mk((r - N == c) ? 1.0 : 0.0, 0.0);
    cplx num;
    if (c < N) {
        if (c - r < -hb || c - r > hb) return mk(0, 0);
        num = band_get(Cb, W, hb, r, c);
    } else {
        int cc = c - N;
        if (cc - r < -hb || cc - r > hb) return mk(0, 0);
        num = band_get(K0s, W, hb, r, cc);
        num = cfma(mk(-w * w, 0), band_get(Mb, W, hb, r, cc), num);
        if (K1c) num = cfma(mk(0, w), band_get(K1c, W, hb, r, cc), num);
    }
    return cneg(cdiv(num, K6d[r]));
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sumf(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx warp_csum(cplx v) {
    v.x = warp_sum(v.x); v.y = warp_sum(v.y);
    return v;
}

__device__ __forceinline__ cplx& HC(cplx* H, int i, int j) { return H[hcol_off(j) + i]; }

__device__ __forceinline__ void make_rot(cplx x, cplx y, cplx& a, cplx& b, double& r) {
    const double r2 = cabs2(x) + cabs2(y);
    if (r2 > 0.0) {
        const double ir = rsqrt(r2);
        a = mk(x.x * ir, -x.y * ir);
        b = mk(y.x * ir, -y.y * ir);
        r = r2 * ir;
    } else { a = mk(1, 0); b = mk(0, 0); r = 0.0; }
}

// Rotation with a REAL cosine (zlartg form): G = [c s; -conj(s) c], G [x; y] = [r; 0] with
// c = |x| / n, s = (x / |x|) conj(y) / n, r = (x / |x|) n, n = sqrt(|x|^2 + |y|^2).  Applying it
// costs 12 instead of 18 FP64 instructions per element pair; generating it needs two rsqrt.
__device__ __forceinline__ void make_rot_rc(cplx x, cplx y, cplx& c, cplx& s, cplx& r) {
    const double nx = cabs2(x), ny = cabs2(y), n2 = nx + ny;
    if (nx > 0.0) {
        const cplx xy = cmulc(x, y);                 // off the rsqrt chain
        const double w = rsqrt(nx) * rsqrt(n2);      // 1 / (|x| n)
        c = mk(nx * w, 0.0);
        s = cscale(xy, w);
        r = cscale(x, n2 * w);
    } else if (ny > 0.0) {
        const double iry = rsqrt(ny);
        c = mk(0, 0); s = mk(y.x * iry, -y.y * iry); r = mk(ny * iry, 0.0);
    } else { c = mk(1, 0); s = mk(0, 0); r = mk(0, 0); }
}
// acc - a * conj(b)
__device__ __forceinline__ cplx cfmsc(cplx a, cplx b, cplx acc) {
    return mk(fma(-a.x, b.x, fma(-a.y, b.y, acc.x)), fma(-a.y, b.x, fma(a.x, b.y, acc.y)));
}
// row rotation [p; q] <- G [p; q] and column rotation [c0 c1] <- [c0 c1] G^H, both variants
template <bool RC>
__device__ __forceinline__ void rot_rows(cplx ga, cplx gb, cplx p, cplx q, cplx& top, cplx& bot) {
    if (RC) {
        top = cfma(gb, q, cscale(p, ga.x));
        bot = cfmsc(p, gb, cscale(q, ga.x));
    } else {
        top = cfma(ga, p, cmul(gb, q));
        bot = csub(cmulc(q, ga), cmulc(p, gb));
    }
}
template <bool RC>
__device__ __forceinline__ void rot_cols(cplx ga, cplx gb, cplx c0, cplx c1, cplx& n0, cplx& n1) {
    if (RC) {
        n0 = cfmac(c1, gb, cscale(c0, ga.x));            // c c0 + conj(s) c1
        n1 = cfms(gb, c0, cscale(c1, ga.x));             // -s c0 + c c1
    } else {
        n0 = cfmac(c0, ga, cmulc(c1, gb));               // c0 conj(a) + c1 conj(b)
        n1 = csub(cmul(c1, ga), cmul(c0, gb));           // -c0 b + c1 a
    }
}
template <bool RC>
__device__ __forceinline__ void gen_rot(cplx x, cplx y, cplx& a, cplx& b, cplx& r) {
    if (RC) make_rot_rc(x, y, a, b, r);
    else { double rr; make_rot(x, y, a, b, rr); r = mk(rr, 0.0); }
}

#endif

#define AXS_CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return (int)e_; } while (0)

// Opt-in to more than 48 KB of dynamic shared memory ONCE per (kernel, device) instead of on every
// launch: axs_smem_once (api.cu) remembers the largest size it has set for a kernel on a device and
// only calls cudaFuncSetAttribute when a launch needs more (the attribute is never lowered).
extern "C" int axs_smem_once(const void* kernel, size_t bytes);
#define AXS_SMEM_ONCE(bytes, ...) do {                                                              \
        const int axs_rc_ = axs_smem_once((const void*)(__VA_ARGS__), (size_t)(bytes));              \
        if (axs_rc_) return axs_rc_;                                                                 \
    } while (0)

// Kernel-selection switches.  Read ONCE from the environment when the library is first used
// (api.cu) and adjustable afterwards through axs_set_tuning() - the launchers never call getenv.
struct AxsTuning {
    int reduce_oc;          // AXS_REDUCE_OC: warps per CTA of the on-chip Hessenberg reduction (12 | 8); 0 = L2-resident k_reduce
    int reduce_prebal;      // AXS_REDUCE_PREBAL: 1 = k_balance in front of k_reduce (2N < 96)
    int hqr_stages;         // AXS_HQR_STAGES: length of the shrinking-window cascade (4)
    int hqr_bulges;         // AXS_HQR_BULGES: bulges in flight per batch (8)
    int hqr_small_direct;   // AXS_HQR_SMALL_DIRECT: problems that fit an SM twice skip the one-CTA-per-SM stage (1)
    int hqr_aed;            // AXS_HQR_AED: aggressive-early-deflation window (0 = off, the default; 2..8)
    int hqr_aed_reg;        // AXS_HQR_AED_REG: 1 = register-resident window routine (shuffles), 0 = shared-memory routine
    int hqr_aed_stages;     // AXS_HQR_AED_STAGES: bit s = AED in the stage with s + 1 CTAs per SM (bit 3: the >= 4 instances); 15 = all
    int backnorm_wy;        // AXS_BACKNORM_WY: blocked compact-WY back-transformation on DMMA tiles (1)
    int backnorm_wy_modes;  // AXS_BACKNORM_WY_MODES: modes per CTA (64 | 32)
    int bpsi_tiled;         // AXS_BPSI_TILED: tiled B psi kernel (1)
    int big_blocked;        // AXS_BIG_BLOCKED: large path, blocked multi-CTA Hessenberg reduction (1); 0 = one CTA per frequency
    int big_panel;          // AXS_BIG_PANEL: panel width of the blocked reduction (32)
    int big_qr_cluster;     // AXS_BIG_QR_CLUSTER: CTAs per matrix in the windowed QR (0 = from the batch size; 1, 2, 4, 8)
};
extern "C" const AxsTuning* axs_tuning(void);

// limits of the shared-memory ("small") eigen path
#define AXS_MAX_SMALL_N2 166

// workspace layout for nf problems of size n2 (all offsets 256-B aligned)
#define AXS_QR_STATS 16
struct AxsWork {
    cplx*   A;      // [nf][n2*n2]  dense column-major work matrix (S -> reflectors below subdiag)
    cplx*   Hc;     // [nf][hcol_size] packed column-major Hessenberg (input of the QR kernel)
    cplx*   Hr;     // [nf][hrow_size] packed row-major Hessenberg (input of inverse iteration)
    cplx*   tau;    // [nf][n2]
    double* scale;  // [nf][n2]  balancing factors D
    int*    ucur;   // [nf]      active-window top between QR stages
    cplx*   Hp;     // [nf][hcol_size] leading block parked between QR stages (Hc stays intact)
    int*    stats;  // [nf][AXS_QR_STATS] QR counters per matrix: macro steps, batches, AED calls, AED deflations,
                    //           then per stage (4 x 3) kilo-cycles of thread 0 in the AED window, its application, the sweeps
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
    if (w) w->stats = (int*)(b + off);
    off += axs_align((long long)nf * AXS_QR_STATS * sizeof(int));
    return off;
}

// ---- global-memory ("large", 2N > AXS_MAX_SMALL_N2) eigen path -------------------------------
// Largest 2N the large path accepts (per-thread register slots of k_modes_col / k_backx_big).
#define AXS_MAX_BIG_N2 8192
#define AXS_BIG_MODE_GRID 592           // persistent CTAs of k_modes_col (4 per SM)
#define AXS_BIG_RB 64                   // row blocks of 128 rows (2N <= 8192)
#define AXS_BIG_CG 32                   // column groups
#define AXS_BIG_YW (4 + AXS_BIG_RB + AXS_BIG_CG + 28)   // y, w (double buffered) + partial sums; the blocked
                                        // reduction (eig_bigred.cu) keeps Y, W (32 rows each), 16 partial rows, T and - from row 88 - the T
                                        // factors of all panels (32 rows) here

struct AxsWorkBig {
    cplx*   A;      // [nf][n2*n2]  dense column-major work matrix (S -> reflectors below subdiag)
    cplx*   Hc;     // [nf][hcol_size] packed column-major Hessenberg, QR works in place here
                    //                 (holds the fp32 |a_ij|^2 tables while balancing)
    cplx*   Hm;     // [nf][hcol_size] second packed copy: input of the inverse iteration
    cplx*   tau;    // [nf][n2]        tau[n2-1].x = max |H_ij|
    double* scale;  // [nf][n2]
    int*    ucur;   // [nf]
    cplx*   yw;     // [nf][AXS_BIG_YW*n2] y = A^H v, w = A v of the Householder update + partial sums
    unsigned char* msc;   // per-CTA scratch of k_modes_col (multipliers, pivots, z, x)
    long long msc_stride; // bytes per CTA
};

// modes per CTA of k_modes_col for a given problem size
static inline int axs_big_modes_per_cta(int n2) { return n2 <= 512 ? 4 : n2 <= 1024 ? 2 : 1; }

static inline long long axs_work_layout_big(int n2, int nf, void* base, AxsWorkBig* w) {
    long long off = 0;
    char* b = (char*)base;
    const long long szA = axs_align((long long)nf * n2 * n2 * sizeof(cplx));
    const long long szH = axs_align((long long)nf * hcol_size(n2) * sizeof(cplx));
    const long long szT = axs_align((long long)nf * n2 * sizeof(cplx));
    const long long szS = axs_align((long long)nf * n2 * sizeof(double));
    const long long stride = axs_align((long long)axs_big_modes_per_cta(n2) * n2 * (3 * sizeof(cplx) + 1));
    if (w) w->A = (cplx*)(b + off);
    off += szA;
    if (w) w->Hc = (cplx*)(b + off);
    off += szH;
    if (w) w->Hm = (cplx*)(b + off);
    off += szH;
    if (w) w->tau = (cplx*)(b + off);
    off += szT;
    if (w) w->scale = (double*)(b + off);
    off += szS;
    if (w) w->ucur = (int*)(b + off);
    off += axs_align((long long)nf * sizeof(int));
    if (w) w->yw = (cplx*)(b + off);
    off += AXS_BIG_YW * szT;
    if (w) { w->msc = (unsigned char*)(b + off); w->msc_stride = stride; }
    off += stride * AXS_BIG_MODE_GRID;
    return off;
}
