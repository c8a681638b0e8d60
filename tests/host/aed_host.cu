// TEST INFRASTRUCTURE: host build of csrc/aed.cuh (the very source the kernel k_hqr_mb runs per warp)
// plus a sequential model of the kernel's batch loop, so that the aggressive-early-deflation logic
// can be checked against LAPACK without a GPU (tests/test_host.py::test_aed_*).
//   build: nvcc -x cu -DAED_HOST -O2 -o aed_host aed_host.cu          (host code only, no kernel)
//   usage: aed_host window <in> <out>     one AED call on the trailing window of [l, u]
//          aed_host solve  <in> <out>     whole eigenvalue solve: deflation scan, AED, multishift sweeps
// file formats (binary, little endian):
//   in : int32 n, l, u, nw, nb ; n*n complex128 column-major (upper Hessenberg)
//   out (window): int32 ns, ok, iters ; T[AED_MAX*AED_MAX] row-major, Z[...] row-major, shift[AED_MAX], spike_new
//   out (solve) : int32 macro, batches, aed_calls, aed_defl, failed ; n complex128 eigenvalues
// AED_REG=1 (environment) runs the register-resident window aed_window_reg on an emulated warp
// (warp_emul.h: 32 fibers, shuffles / ballots as rendezvous) instead of the shared-memory aed_window.
#include "../../axisym-safe-python_b200/csrc/common.cuh"
#include "warp_emul.h"
static inline cplx wemu_shfl_c(int lane, cplx v, int src, int tag) { wemu::shfl2(lane, v.x, v.y, src, tag); return v; }
#define AEDR_SHFL(v, src) wemu_shfl_c(lane, (v), (src), __LINE__)
#define AEDR_BALLOT(p) wemu::ballot(lane, (p), __LINE__)
#define AEDR_CLZ(m) __builtin_clz((unsigned)(m))
#define AEDW_SYNC() wemu::syncwarp(lane, __LINE__)
#define AEDW_BALLOT(p) wemu::ballot(lane, (p), __LINE__)
#define AEDW_CLZ(m) __builtin_clz((unsigned)(m))
#include "../../axisym-safe-python_b200/csrc/aed.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define MB_D 3
static long g_iters = 0;
static int g_stage[4] = {0, 0, 0, 0}, g_mstage[4] = {0, 0, 0, 0};
static double g_work = 0.0;   // sum over batches of macro steps x active size (cost model of the windowed large-path QR)
#define MB_MAX 8
typedef std::vector<cplx> Packed;
static int use_reg() { static int v = getenv("AED_REG") ? atoi(getenv("AED_REG")) : 0; return v; }
// one window call: AED_REG=0 the shared-memory routine as sequential code; 1 the register routine and
// 2 the one-loop shared-memory routine (aed_window_flat), both on the emulated warp
static void window_call(const cplx* H, int l, int u, int nw)
{
    if (use_reg() == 1) wemu::run([=](int lane) { aed_window_reg(H, l, u, nw, lane); });
    else if (use_reg() == 2) wemu::run([=](int lane) { aed_window_flat(H, l, u, nw, lane); });
    else aed_window(H, l, u, nw, 0);
}
static inline cplx& HP(Packed& H, int i, int j) { return H[hcol_off(j) + i]; }

// one implicit single-shift QR sweep on the active block [l, u]: what ONE bulge of k_hqr_mb does
// (phase R: rows k, k+1 of the columns k-1 .. u; phase C: columns k, k+1 of the rows l .. k+1, the
// bulge H(k+2, k) lives in a register).  Bulges of a batch commute, so a batch = nb such sweeps.
static void sweep(Packed& H, int l, int u, cplx sig)
{
    cplx x = csub(HP(H, l, l), sig), y = HP(H, l + 1, l);
    for (int k = l; k < u; ++k) {
        double c; cplx s, r;
        aed_rot(x, y, c, s, r);
        if (k > l) HP(H, k, k - 1) = r;
        for (int j = k; j <= u; ++j) aed_rows(c, s, HP(H, k, j), HP(H, k + 1, j));
        for (int i = l; i <= k + 1; ++i) aed_cols(c, s, HP(H, i, k), HP(H, i, k + 1));
        if (k + 2 <= u) {
            const cplx h = HP(H, k + 2, k + 1);
            y = cmulc(h, s);                             // new bulge H(k+2, k) = h conj(s)
            HP(H, k + 2, k + 1) = cscale(h, c);
            x = HP(H, k + 1, k);
        }
    }
}

static bool negligible(Packed& H, int n, int k)
{
    const double sub = cabs1(HP(H, k, k - 1));
    if (sub <= AED_SMALL) return true;
    const cplx hkk = HP(H, k, k), hk1 = HP(H, k - 1, k - 1);
    double tst = cabs1(hk1) + cabs1(hkk);
    if (tst == 0.0) {
        if (k - 2 >= 0) tst += cabs1(HP(H, k - 1, k - 2));
        if (k + 1 <= n - 1) tst += cabs1(HP(H, k + 1, k));
    }
    if (sub <= AED_EPS * tst) {
        const double up = cabs1(HP(H, k - 1, k));
        const double ab = fmax(sub, up), ba = fmin(sub, up);
        const double dd = cabs1(csub(hk1, hkk));
        const double aa = fmax(cabs1(hkk), dd), bb = fmin(cabs1(hkk), dd);
        const double sm = aa + ab;
        if (ba * (ab / sm) <= fmax(AED_SMALL, AED_EPS * (bb * (aa / sm)))) return true;
    }
    return false;
}

// mirror of the while loop of k_hqr_mb (eig_small.cu), m_stop = 0
static void solve(Packed& H, int n, int aed_nw, int bmax, std::vector<cplx>& eig, int* st)
{
    AedScratch S;
    cplx s_shift[AED_MAX > MB_MAX ? AED_MAX : MB_MAX];
    int u = n - 1, its = 0, failed = 0, n_macro = 0, n_batch = 0, n_aed = 0, n_aed_defl = 0;
    bool have_shifts = false;
    int ns_aed = 0;
    while (u >= 0) {
        int l = 0;
        for (int k = u; k >= 1; --k) if (negligible(H, n, k)) { l = k; break; }
        if (l > 0) HP(H, l, l - 1) = mk(0, 0);
        if (l == u) { eig[u] = HP(H, u, u); --u; its = 0; have_shifts = false; continue; }
        ++its;
        if (its > 120) { eig[u] = HP(H, u, u); ++failed; --u; its = 0; have_shifts = false; continue; }
        have_shifts = false;
        static int p_nibble = getenv("AED_NIBBLE") ? atoi(getenv("AED_NIBBLE")) : 14;
        static int p_minm = getenv("AED_MINM") ? atoi(getenv("AED_MINM")) : 3;
        static int p_every = getenv("AED_EVERY") ? atoi(getenv("AED_EVERY")) : 1;
        static int p_umin = getenv("AED_UMIN") ? atoi(getenv("AED_UMIN")) : 0;   // AED only while u >= p_umin (stage mask of the kernel)
        static int batch_ctr = 0;
        if (aed_nw >= 2 && u >= p_umin && u - l + 1 >= p_minm && (batch_ctr++ % p_every) == 0) {
            const int nw = (aed_nw < u - l + 1) ? aed_nw : u - l + 1;
            const int kw = u - nw + 1;
            window_call(H.data(), l, u, nw); S = g_aed;
            ++n_aed; g_stage[u >= 116 ? 0 : u >= 93 ? 1 : u >= 79 ? 2 : 3]++; g_iters += S.iters; if (getenv("AED_TRACE")) fprintf(stderr, "aed u=%d l=%d nw=%d iters=%d ns=%d\n", u, l, nw, S.iters, S.ns);
            const int ok = S.ok, ns = S.ns, nd = nw - ns;
            if (ok && nd > 0) {
                for (int i = l; i < kw; ++i) {
                    cplx out[AED_MAX];
                    for (int j = 0; j < nw; ++j) {
                        cplx acc = mk(0, 0);
                        for (int q = 0; q < nw; ++q) acc = cfma(HP(H, i, kw + q), S.Z[q][j], acc);
                        out[j] = acc;
                    }
                    for (int j = 0; j < nw; ++j) HP(H, i, kw + j) = out[j];
                }
                for (int j = 0; j < nw; ++j)
                    for (int i = 0; i < nw; ++i)
                        if (i <= j + 1) HP(H, kw + i, kw + j) = S.T[i][j];
                if (kw > l) HP(H, kw, kw - 1) = S.spike_new;
                for (int j = ns; j < nw; ++j) eig[kw + j] = S.T[j][j];
                n_aed_defl += nd;
                u = kw + ns - 1;
                its = 0;
            }
            if (ok) {
                for (int j = 0; j < ns; ++j) s_shift[j] = S.shift[j];
                have_shifts = ns > 0;
                ns_aed = ns;
            }
            if (ok && nd > 0) {
                if (u < l + 1 || (p_nibble > 0 && nd * 100 >= p_nibble * nw)) continue;
            }
        }
        const int m = u - l + 1;
        if (m < 2) continue;
        int nb = (m - 2) / MB_D + 1;
        if (nb > bmax) nb = bmax;
        if (have_shifts && its % 10 != 0) {
            if (nb > ns_aed) nb = ns_aed;
            cplx tmp[AED_MAX];
            const cplx huu = HP(H, u, u);
            for (int t = 0; t < ns_aed; ++t) {
                const double dme = cabs1(csub(s_shift[t], huu));
                int rank = 0;
                for (int q = 0; q < ns_aed; ++q) {
                    const double dq = cabs1(csub(s_shift[q], huu));
                    rank += (dq < dme || (dq == dme && q < t)) ? 1 : 0;
                }
                tmp[rank] = s_shift[t];
            }
            for (int t = 0; t < ns_aed; ++t) s_shift[t] = tmp[t];
        } else {
            for (int t = 0; t < (nb + 1) / 2; ++t) {
                const int i = u - 2 * t;
                const cplx a11 = HP(H, i - 1, i - 1), a12 = HP(H, i - 1, i), a21 = HP(H, i, i - 1), a22 = HP(H, i, i);
                const cplx hd = cscale(csub(a11, a22), 0.5);
                const cplx disc = csqrt_(cadd(cmul(hd, hd), cmul(a12, a21)));
                const cplx mid = cscale(cadd(a11, a22), 0.5);
                cplx e1 = cadd(mid, disc), e2 = csub(mid, disc);
                if (cabs1(csub(e2, a22)) < cabs1(csub(e1, a22))) { const cplx tt = e1; e1 = e2; e2 = tt; }
                if (its % 10 == 0) {
                    const double ex = 0.75 * cabs1(HP(H, u, u - 1));
                    e1 = cadd(e1, mk(ex, 0)); e2 = cadd(e2, mk(0, ex));
                }
                s_shift[2 * t] = e1;
                if (2 * t + 1 < MB_MAX) s_shift[2 * t + 1] = e2;
            }
        }
        for (int b = 0; b < nb; ++b) sweep(H, l, u, s_shift[b]);
        n_macro += (m - 1) + (nb - 1) * MB_D;
        g_mstage[u >= 116 ? 0 : u >= 93 ? 1 : u >= 79 ? 2 : 3] += (m - 1) + (nb - 1) * MB_D;
        g_work += (double)((m - 1) + (nb - 1) * MB_D) * m;
        ++n_batch;
    }
    st[0] = n_macro; st[1] = n_batch; st[2] = n_aed; st[3] = n_aed_defl; st[4] = failed;
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: aed_host window|solve in out\n"); return 2; }
    FILE* fi = fopen(argv[2], "rb");
    if (!fi) return 3;
    int hdr[5];
    if (fread(hdr, sizeof(int), 5, fi) != 5) return 4;
    const int n = hdr[0], l = hdr[1], u = hdr[2], nw = hdr[3], nb = hdr[4];
    std::vector<cplx> dense((size_t)n * n);
    if (fread(dense.data(), sizeof(cplx), (size_t)n * n, fi) != (size_t)n * n) return 5;
    fclose(fi);
    Packed H(hcol_size(n) + 2);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i <= j + 1 && i < n; ++i) H[hcol_off(j) + i] = dense[(size_t)j * n + i];
    FILE* fo = fopen(argv[3], "wb");
    if (!fo) return 6;
    if (!strcmp(argv[1], "window")) {
        AedScratch S;
        memset(&S, 0, sizeof S);
        window_call(H.data(), l, u, nw); S = g_aed;
        int o[3] = {S.ns, S.ok, S.iters};
        fwrite(o, sizeof(int), 3, fo);
        for (int i = 0; i < AED_MAX; ++i) fwrite(S.T[i], sizeof(cplx), AED_MAX, fo);
        for (int i = 0; i < AED_MAX; ++i) fwrite(S.Z[i], sizeof(cplx), AED_MAX, fo);
        fwrite(S.shift, sizeof(cplx), AED_MAX, fo);
        fwrite(&S.spike_new, sizeof(cplx), 1, fo);
    } else {
        std::vector<cplx> eig(n);
        int st[5];
        solve(H, n, nw, nb, eig, st);
        if (getenv("AED_STAGES")) fprintf(stderr, "WORK %.4g  warp collectives (complex shuffles + ballots) %ld\n", g_work, wemu::g.collectives);
        if (getenv("AED_STAGES")) fprintf(stderr, "STAGES macro %d %d %d %d aed %d %d %d %d iters %ld\n", g_mstage[0], g_mstage[1], g_mstage[2], g_mstage[3], g_stage[0], g_stage[1], g_stage[2], g_stage[3], g_iters);
        fwrite(st, sizeof(int), 5, fo);
        fwrite(eig.data(), sizeof(cplx), n, fo);
    }
    fclose(fo);
    return 0;
}
