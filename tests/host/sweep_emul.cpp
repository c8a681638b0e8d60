// TEST INFRASTRUCTURE: the eigen-solve of ONE frequency point through the product's own kernel sources on
// the emulated thread block of cta_emul.h - what the GPU does for every point of the sweep, without a GPU:
//   k_balance (S given dense, zgebal scaling)  ->  k_reduce_oc (on-chip Householder Hessenberg reduction; 2N < 96:
//   k_reduce, the L2-resident one)
//   ->  k_hqr_mb cascade (eigenvalues)  ->  k_modes (inverse iteration)  ->  k_backnorm (back-transformation,
//   un-balancing, B-normalisation, psi_hat layout).
// 16 <= 2N <= 128 runs BOTH back-transformations on the same k_modes output: the library's default there, the
// blocked compact-WY kernels k_wy_tfactor + k_backnorm_wy on FP64 tensor-core tiles (DMMA and cp.async through
// the models of tests/host/shim), and k_backnorm (the per-reflector instance for 128 < 2N <= 166 and behind
// backnorm_wy = 0); psi is the default path's, the largest relative difference of the two is reported.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/sweep_emul tests/host/sweep_emul.cpp
//   usage: sweep_emul <in> <out>
//   in : int32 N, hb, 0, 0, 0 ; S[(2N)^2] complex128 column-major ; Cb[N][2hb+1] complex128 ; K6d[N] complex128
//   out: int32 info, launches ; float64 max |psi_wy - psi_slab| / max |psi| (-1: one path only) ; scale[2N] float64 ;
//        k[2N] complex128 ; psi[2N][2N] complex128 (psi_hat layout)
// The launch logic mirrors axs_launch_reduce_oc (eig_reduce.cu), axs_launch_hqr / hqr_cascade and
// axs_launch_modes (eig_small.cu); the kernels are the product source.
#include "shim/cuda_runtime.h"
#include "shim/cuda_pipeline.h"
#include "shim/dmma.h"
#include "../../axisym-safe-python_b200/csrc/reduce_oc.cuh"
#include "../../axisym-safe-python_b200/csrc/reduce_l2.cuh"
#include "../../axisym-safe-python_b200/csrc/hqr_mb.cuh"
#include "../../axisym-safe-python_b200/csrc/modes.cuh"
#include "../../axisym-safe-python_b200/csrc/backnorm.cuh"
#include "../../axisym-safe-python_b200/csrc/backx_kernels.cuh"
#include <vector>

static int g_launches = 0;
static void fits(size_t bytes) { if (bytes > sizeof(cta::g.dyn_smem)) { fprintf(stderr, "shared memory: %zu bytes\n", bytes); exit(7); } }

static int capacity(int per_sm)
{
    int m = 0;
    while (hqr_mb_smem(m + 1) + 2048 <= (size_t)(227 * 1024 / per_sm)) ++m;
    return m;
}
template <int NT, int MINB, int NITEMS, bool OPT>
static void launch_mb(int n, int first, int m_stop, AxsWork wk, cplx* eig, int* info)
{
    ++g_launches;
    cta::run(NT, 0, [&] { k_hqr_mb<NT, MINB, NITEMS, OPT, false>(n, MB_MAX, first, m_stop, 0, wk.Hc, wk.Hp, eig, info, wk.ucur, wk.stats); });
}
static void cascade(int n, int m_cur, bool from_Hc, AxsWork wk, cplx* eig, int* info)
{
    const int m3 = capacity(3), m4 = capacity(4), m5 = from_Hc ? capacity(5) : 0;
    int first = from_Hc ? 1 : 0;
    if (m_cur <= 35 && MB_MAX * (m_cur + 1) <= 3 * (128 - 32)) { launch_mb<128, 8, 3, false>(n, first, 0, wk, eig, info); return; }
    if (m3 == 0 || m_cur > m3) {
        const int stop = (m3 >= 8 && m3 < m_cur) ? m3 : 0;
        launch_mb<512, 2, 3, false>(n, first, stop, wk, eig, info);
        if (!stop) return;
        first = 0; m_cur = stop;
    }
    if (m4 == 0 || m_cur > m4) {
        const int stop = (m4 >= 8 && m4 < m_cur) ? m4 : 0;
        launch_mb<384, 3, 3, false>(n, first, stop, wk, eig, info);
        if (!stop) return;
        first = 0; m_cur = stop;
    }
    if (!(m5 >= 8 && m_cur <= m5)) {
        const int stop = (m5 >= 8 && m5 < m_cur) ? m5 : 0;
        launch_mb<256, 4, 3, false>(n, first, stop, wk, eig, info);
        if (!stop) return;
        first = 0; m_cur = stop;
    }
    launch_mb<256, 5, 3, false>(n, first, 0, wk, eig, info);
}

template <int NWARP, int R, bool PAD>
static void launch_oc(int n, AxsWork wk)
{
    const size_t cap = 227 * 1024, fixed = reduce_oc_fixed<NWARP, R>();
    const size_t colb = (size_t)(PAD ? 32 * R : n) * sizeof(cplx);
    int ncs = (int)((cap - fixed) / colb);
    if (ncs > n - 1) ncs = n - 1;
    const int jg = n - ncs;
    fits(fixed + (size_t)ncs * colb);
    ++g_launches;
    cta::run(NWARP * 32, 0, [&] { k_reduce_oc<NWARP, R, PAD, false>(n, jg, wk); });
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: sweep_emul in out\n"); return 2; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    int hdr[5];
    if (fread(hdr, sizeof(int), 5, fi) != 5) return 4;
    const int N = hdr[0], hb = hdr[1], n = 2 * N, W = 2 * hb + 1;
    if (n < 4 || n > AXS_MAX_SMALL_N2) { fprintf(stderr, "the shared-memory eigen path serves 2N <= %d\n", AXS_MAX_SMALL_N2); return 8; }
    std::vector<cplx> S((size_t)n * n), Cb((size_t)N * W), K6d(N), eig(n), psi((size_t)n * n);
    if (fread(S.data(), sizeof(cplx), S.size(), fi) != S.size()) return 5;
    if (fread(Cb.data(), sizeof(cplx), Cb.size(), fi) != Cb.size()) return 5;
    if (fread(K6d.data(), sizeof(cplx), K6d.size(), fi) != K6d.size()) return 5;
    fclose(fi);
    std::vector<char> buf((size_t)axs_work_layout(n, 1, nullptr, nullptr) + 4096);
    AxsWork wk;
    axs_work_layout(n, 1, (void*)(((uintptr_t)buf.data() + 255) & ~(uintptr_t)255), &wk);
    int info = 0;
    // ---- axs_launch_reduce: k_balance, then the on-chip reduction (2N >= 96) or the L2-resident one ---------
    fits(balance_smem(n));
    ++g_launches;
    cta::run(BAL_THREADS, 0, [&] { k_balance(N, hb, nullptr, Cb.data(), nullptr, nullptr, K6d.data(), nullptr, S.data(), wk); });
    if (n < 96) {
        fits(reduce_smem(n));
        ++g_launches;
        cta::run(RED_THREADS, 0, [&] { k_reduce(N, hb, nullptr, Cb.data(), nullptr, nullptr, K6d.data(), nullptr, S.data(), wk, 1); });
    } else if (n <= 128) launch_oc<12, 4, true>(n, wk);
    else launch_oc<8, 6, false>(n, wk);
    // ---- axs_launch_hqr ----------------------------------------------------------------------------
    int m2 = capacity(2);
    if (m2 >= n || m2 < 8) m2 = 0;
    if (m2 == 0) cascade(n, n, true, wk, eig.data(), &info);
    else {
        fits(hqr_mb_smem(n));
        if (MB_MAX * n <= 3 * 384) launch_mb<512, 1, 3, true>(n, 1, m2, wk, eig.data(), &info);
        else launch_mb<512, 1, 3, false>(n, 1, m2, wk, eig.data(), &info);
        cascade(n, m2, false, wk, eig.data(), &info);
    }
    // ---- axs_launch_modes (k_modes + k_backnorm) -----------------------------------------------------
    int nw = 4;
    while (nw > 1 && modes_smem(n, nw) > 227 * 1024) --nw;
    fits(modes_smem(n, nw));
    for (int b = 0; b < (n + 8 * nw - 1) / (8 * nw); ++b) {
        ++g_launches;
        cta::run(32 * nw, b, [&] { k_modes(N, hb, Cb.data(), K6d.data(), eig.data(), psi.data(), wk); });
    }
    std::vector<cplx> X(psi), psi_slab;
    fits(backnorm_smem(n));
    for (int b = 0; b < (n + BN_COLS - 1) / BN_COLS; ++b) {
        ++g_launches;
        cta::run(256, b, [&] { k_backnorm(N, hb, Cb.data(), K6d.data(), eig.data(), psi.data(), wk); });
    }
    double wy_diff = -1.0;
    // ---- axs_launch_backnorm_wy (the default for 16 <= 2N <= 128) --------------------------------------
    const int nblk = (n - 2 + WY_NB - 1) / WY_NB;
    if (n <= WY_MAX_N && n >= 16 && (long long)nblk * WY_NB * WY_NB <= (long long)hcol_size(n)) {
        psi_slab = psi;
        psi = X;
        cplx* Tg = wk.Hp;                                  // the QR parking buffer is dead by now
        for (int b = 0; b < nblk; ++b) {
            ++g_launches;
            cta::run(32, b, [&] { k_wy_tfactor(n, wk, Tg, nblk); }, 0, nblk, 1);
        }
        fits(backnorm_wy_smem(n, 64));
        for (int b = 0; b < (n + 63) / 64; ++b) {
            ++g_launches;
            cta::run(256, b, [&] { k_backnorm_wy<64>(N, hb, Cb.data(), K6d.data(), eig.data(), psi.data(), wk, Tg, nblk); }, 0, (n + 63) / 64, 1);
        }
        double top = 0.0, dif = 0.0;
        for (size_t q = 0; q < psi.size(); ++q) {
            top = fmax(top, cabs1(psi[q]));
            dif = fmax(dif, cabs1(csub(psi[q], psi_slab[q])));
        }
        wy_diff = dif / top;
    }
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    int o[2] = {info, g_launches};
    fwrite(o, sizeof(int), 2, fo);
    fwrite(&wy_diff, sizeof(double), 1, fo);
    fwrite(wk.scale, sizeof(double), n, fo);
    fwrite(eig.data(), sizeof(cplx), n, fo);
    fwrite(psi.data(), sizeof(cplx), psi.size(), fo);
    fclose(fo);
    return 0;
}
