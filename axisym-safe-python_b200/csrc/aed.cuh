// Aggressive early deflation (AED) on the trailing window of the active Hessenberg block:
// the step LAPACK's zhseqr (behind scipy.linalg.eig, solver.py:136) performs in zlaqr2 / zlaqr3 in
// front of every multishift sweep.  One WARP works on the nw x nw window (nw <= AED_MAX) held in
// shared memory:
//   1. Schur form T = Z^H W Z of the window by the implicit single-shift QR algorithm (zlahqr),
//   2. spike test from the bottom (zlaqr2): the window eigenvalue T(ns, ns) is converged when
//      |spike| |Z(1, ns)| <= max(smlnum, ulp |T(ns, ns)|); undeflatable eigenvalues are moved to the
//      top of the window by adjacent swaps (ztrexc),
//   3. the leading ns x ns part and the spike vector are returned to Hessenberg form (here by
//      plane rotations with bulge chasing instead of zlarfg + zgehrd: same similarity class),
//   4. the undeflated window eigenvalues are the shifts of the next sweep.
// The CTA then applies Z to the rows of the active block above the window and writes T back.
//
// The code is written against three macros so that the very same source runs as a sequential host
// function (tests/host/aed_host.cu compiles it with -DAED_HOST and checks it against LAPACK):
//   AED_LANES(v, lo, hi)  loop whose iterations are independent (device: strided over the lanes)
//   AED_SYNC()            all iterations before it are visible to every lane (device: __syncwarp)
#pragma once
#include "common.cuh"

#define AED_MAX 8

struct AedScratch {
    cplx T[AED_MAX][AED_MAX + 1];       // window / Schur form / Hessenberg + triangular result
    cplx Z[AED_MAX][AED_MAX + 1];       // accumulated unitary
    cplx shift[AED_MAX];                // undeflated window eigenvalues (shifts of the next sweep)
    cplx sv[AED_MAX];                   // spike vector while it is reduced
    cplx spike_new;                     // H(kw, kw-1) after the deflation
    int ns;                             // number of undeflated eigenvalues (0 .. nw)
    int ok;                             // 0: the window QR did not converge, the window is left alone
    int iters;                          // QR sweeps spent on the window (statistics)
#ifdef AED_PROFILE
    long long prof[8];                  // cycles: load, schur, reorder, rehess; steps: schur rotations, swaps, rehess rotations, calls
#endif
};

#ifdef AED_HOST
#define AED_FN inline
#define AED_INL inline
#define AED_LANES(v, lo, hi) for (int v = (lo); v < (hi); ++v)
#define AED_SYNC()
#define AED_EPS 2.220446049250313e-16
#define AED_SMALL (2.2250738585072014e-308 / AED_EPS)
#define AED_RSQRT(x) (1.0 / sqrt(x))
static AedScratch g_aed;                 // host: one scratch object
#else
#define AED_FN __device__ __noinline__
#define AED_INL __device__ __forceinline__
#define AED_LANES(v, lo, hi) for (int v = (lo) + lane; v < (hi); v += 32)
#define AED_SYNC() __syncwarp()
#define AED_EPS EPS_D
#define AED_SMALL SMALL_D
#define AED_RSQRT(x) rsqrt(x)
// one scratch object per CTA at file scope: the (non-inlined) window routine addresses it with LDS / STS
// instead of generic loads through a pointer
__shared__ AedScratch g_aed;
#endif

// rotation with a real cosine: [c s; -conj(s) c] [x; y] = [r; 0]   (zlartg; make_rot_rc of common.cuh)
AED_INL void aed_rot(cplx x, cplx y, double& c, cplx& s, cplx& r) {
    const double nx = cabs2(x), ny = cabs2(y), n2 = nx + ny;
    if (ny == 0.0) { c = 1.0; s = mk(0, 0); r = x; }
    else if (nx > 0.0) {
        const cplx xy = cmulc(x, y);                      // off the rsqrt chain
        const double w = AED_RSQRT(nx) * AED_RSQRT(n2);   // 1 / (|x| n)
        c = nx * w;
        s = cscale(xy, w);
        r = cscale(x, n2 * w);
    } else {
        const double iry = AED_RSQRT(ny);
        c = 0.0; s = mk(y.x * iry, -y.y * iry); r = mk(ny * iry, 0.0);
    }
}
// rows:  [p; q] <- [c s; -conj(s) c] [p; q]
AED_INL void aed_rows(double c, cplx s, cplx& p, cplx& q) {
    const cplx top = cfma(s, q, cscale(p, c));
    const cplx bot = csub(cscale(q, c), cmulc(p, s));
    p = top; q = bot;
}
// columns:  [a b] <- [a b] [c s; -conj(s) c]^H  =  [c a + conj(s) b,  -s a + c b]
AED_INL void aed_cols(double c, cplx s, cplx& a, cplx& b) {
    const cplx n0 = cfmac(b, s, cscale(a, c));
    const cplx n1 = cfms(s, a, cscale(b, c));
    a = n0; b = n1;
}

// eigenvalue of [[a11 a12] [a21 a22]] nearer to a22
AED_INL cplx aed_shift2(cplx a11, cplx a12, cplx a21, cplx a22) {
    const cplx hd = cscale(csub(a11, a22), 0.5);
    const cplx disc = csqrt_(cadd(cmul(hd, hd), cmul(a12, a21)));
    const cplx mid = cscale(cadd(a11, a22), 0.5);
    const cplx e1 = cadd(mid, disc), e2 = csub(mid, disc);
    return (cabs1(csub(e2, a22)) < cabs1(csub(e1, a22))) ? e2 : e1;
}

// zlahqr's test for a negligible subdiagonal entry T(k, k-1) (standard + Ahues-Tisseur)
AED_INL bool aed_negligible(const AedScratch* S, int nw, int k) {
    const double sub = cabs1(S->T[k][k - 1]);
    if (sub <= AED_SMALL) return true;
    const cplx hkk = S->T[k][k], hk1 = S->T[k - 1][k - 1];
    double tst = cabs1(hk1) + cabs1(hkk);
    if (tst == 0.0) {
        if (k - 2 >= 0) tst += cabs1(S->T[k - 1][k - 2]);
        if (k + 1 <= nw - 1) tst += cabs1(S->T[k + 1][k]);
    }
    if (sub <= AED_EPS * tst) {
        const double up = cabs1(S->T[k - 1][k]);
        const double ab = fmax(sub, up), ba = fmin(sub, up);
        const double dd = cabs1(csub(hk1, hkk));
        const double aa = fmax(cabs1(hkk), dd), bb = fmin(cabs1(hkk), dd);
        const double sm = aa + ab;
        if (ba * (ab / sm) <= fmax(AED_SMALL, AED_EPS * (bb * (aa / sm)))) return true;
    }
    return false;
}

// One similarity rotation of the window in the plane (k, k+1): rows k, k+1 of the columns
// jlo .. nw-1, then columns k, k+1 of the rows 0 .. ihi, and columns k, k+1 of Z.
AED_INL void aed_apply(AedScratch* S, int nw, int k, int jlo, int ihi, double c, cplx s, int lane) {
    AED_LANES(j, jlo, nw) aed_rows(c, s, S->T[k][j], S->T[k + 1][j]);
    AED_SYNC();
    AED_LANES(i, 0, 2 * nw) {
        if (i < nw) { if (i <= ihi) aed_cols(c, s, S->T[i][k], S->T[i][k + 1]); }
        else aed_cols(c, s, S->Z[i - nw][k], S->Z[i - nw][k + 1]);
    }
    AED_SYNC();
}

// H: packed column-major Hessenberg matrix (hcol_off), active block [l, u], window kw = u-nw+1 .. u.
AED_FN void aed_window(const cplx* H, int l, int u, int nw, int lane)
{
    AedScratch* const S = &g_aed;
    const int kw = u - nw + 1;
    AED_LANES(idx, 0, nw * nw) {
        const int j = idx / nw, i = idx - j * nw;
        S->T[i][j] = (i <= j + 1) ? H[hcol_off(kw + j) + kw + i] : mk(0, 0);
        S->Z[i][j] = mk(i == j ? 1.0 : 0.0, 0.0);
    }
    const cplx spike = (kw > l) ? H[hcol_off(kw - 1) + kw] : mk(0, 0);
    AED_SYNC();
#ifdef AED_PROFILE
    long long pc0 = clock64(), pc1; int pn_rot = 0, pn_swap = 0, pn_reh = 0;
#define AED_PMARK(i) { pc1 = clock64(); if (lane == 0) S->prof[i] += pc1 - pc0; pc0 = pc1; }
#define AED_PCNT(v) ++v
#else
#define AED_PMARK(i)
#define AED_PCNT(v)
#endif

    // ---- 1. Schur form of the window (implicit single-shift QR, zlahqr tests) -----------------
    int iu = nw - 1, its = 0, total = 0, ok = 1;
    while (iu >= 1) {
        // largest k <= iu with a negligible subdiagonal entry (device: lane k tests entry k, one ballot)
        int il = 0;
#ifdef AED_HOST
        for (int k = iu; k >= 1 && il == 0; --k) if (aed_negligible(S, nw, k)) il = k;
#else
        {
            const bool neg = (lane >= 1 && lane <= iu) ? aed_negligible(S, nw, lane) : false;
            const unsigned m = __ballot_sync(0xffffffffu, neg);
            il = m ? 31 - __clz(m) : 0;
        }
#endif
        AED_SYNC();                                       // every lane has read the subdiagonal
        AED_PMARK(0)
        if (il > 0) { AED_LANES(q, 0, 1) S->T[il][il - 1] = mk(0, 0); }
        if (il == iu) { --iu; its = 0; AED_SYNC(); continue; }
        ++its; ++total;
        if (its > 30) { ok = 0; break; }
        cplx sig = aed_shift2(S->T[iu - 1][iu - 1], S->T[iu - 1][iu], S->T[iu][iu - 1], S->T[iu][iu]);
        if (its == 10 || its == 20) sig = cadd(S->T[iu][iu], mk(0.75 * cabs1(S->T[iu][iu - 1]), 0));
        AED_SYNC();
        AED_PMARK(1)
        for (int k = il; k < iu; ++k) {
            const cplx x = (k == il) ? csub(S->T[il][il], sig) : S->T[k][k - 1];
            const cplx y = (k == il) ? S->T[il + 1][il] : S->T[k + 1][k - 1];
            double c; cplx s, r;
            aed_rot(x, y, c, s, r);
            AED_SYNC();                                   // x, y read by every lane before they change
            AED_PMARK(2)
            if (k > il) { AED_LANES(q, 0, 1) { S->T[k][k - 1] = r; S->T[k + 1][k - 1] = mk(0, 0); } }
            aed_apply(S, nw, k, k, (k + 2 < iu) ? k + 2 : iu, c, s, lane);
            AED_PMARK(3)
            AED_PCNT(pn_rot);
        }
    }
    S->ok = ok; S->iters = total;
#ifdef AED_PROFILE
    if (lane == 0) S->prof[5] += total;
#endif                         // (every lane stores the same value)
    if (!ok) { S->ns = nw; AED_SYNC(); return; }

    // ---- 2. spike test from the bottom, undeflatable eigenvalues to the top (zlaqr2, ztrexc) ---
    int ns = nw, ilst = 0;
    const double sp1 = cabs1(spike);
    while (ilst < ns) {
        double foo = cabs1(S->T[ns - 1][ns - 1]);
        if (foo == 0.0) foo = sp1;
        if (sp1 * cabs1(S->Z[0][ns - 1]) <= fmax(AED_SMALL, AED_EPS * foo)) { --ns; continue; }
        for (int k = ns - 2; k >= ilst; --k) {
            const cplx t11 = S->T[k][k], t22 = S->T[k + 1][k + 1];
            double c; cplx s, r;
            aed_rot(S->T[k][k + 1], csub(t22, t11), c, s, r);
            AED_SYNC();
            AED_LANES(j, k + 2, nw) aed_rows(c, s, S->T[k][j], S->T[k + 1][j]);
            AED_LANES(i, 0, 2 * nw) {
                if (i < nw) { if (i < k) aed_cols(c, s, S->T[i][k], S->T[i][k + 1]); }
                else aed_cols(c, s, S->Z[i - nw][k], S->Z[i - nw][k + 1]);
            }
            AED_LANES(q, 0, 1) { S->T[k][k] = t22; S->T[k + 1][k + 1] = t11; }
            AED_SYNC();
            AED_PCNT(pn_swap);
        }
        ++ilst;
    }
    AED_LANES(j, 0, ns) S->shift[j] = S->T[j][j];
    S->ns = ns;
#ifdef AED_PROFILE
    if (lane == 0) { S->prof[4] += pn_rot; S->prof[7] += 1; }
#endif
    if (ns == nw) { AED_SYNC(); return; }                 // nothing deflated: the caller leaves H alone

    // ---- 3. spike vector and leading ns x ns part back to Hessenberg form ----------------------
    AED_LANES(j, 0, nw) S->sv[j] = (j < ns) ? cmulc(spike, S->Z[0][j]) : mk(0, 0);
    AED_SYNC();
    for (int j = ns - 1; j >= 1; --j) {
        {
            double c; cplx s, r;
            aed_rot(S->sv[j - 1], S->sv[j], c, s, r);
            AED_SYNC();
            AED_LANES(q, 0, 1) { S->sv[j - 1] = r; S->sv[j] = mk(0, 0); }
            aed_apply(S, nw, j - 1, j - 1, (j + 1 < ns - 1) ? j + 1 : ns - 1, c, s, lane);
            AED_PCNT(pn_reh);
        }
        for (int kk = j; kk + 1 <= ns - 1; ++kk) {         // chase the bulge T(kk+1, kk-1) out of the bottom
            double c; cplx s, r;
            aed_rot(S->T[kk][kk - 1], S->T[kk + 1][kk - 1], c, s, r);
            AED_SYNC();
            AED_LANES(q, 0, 1) { S->T[kk][kk - 1] = r; S->T[kk + 1][kk - 1] = mk(0, 0); }
            aed_apply(S, nw, kk, kk, (kk + 2 < ns - 1) ? kk + 2 : ns - 1, c, s, lane);
            AED_PCNT(pn_reh);
        }
    }
    AED_LANES(q, 0, 1) S->spike_new = (ns > 0) ? S->sv[0] : mk(0, 0);
    AED_SYNC();
#ifdef AED_PROFILE
    if (lane == 0) S->prof[6] += pn_reh;
#endif
}

// Two other formulations of this routine were built, checked on the host and measured on B200, and lost
// (DESIGN.md 4.5): aed_window_flat (one loop around one rotation step) and aed_window_reg (T and Z in
// registers, rotations through shuffles).  They live in aed_variants.cuh and are compiled only into the A/B
// builds (-DAXS_AED_BUILD=1|2|3) and into the host test program.
#if defined(AED_HOST) || (defined(AXS_AED_BUILD) && AXS_AED_BUILD != 0)
#include "aed_variants.cuh"
#endif
