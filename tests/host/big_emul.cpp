// TEST INFRASTRUCTURE: one frequency point through the LARGE path (166 < 2N) on the emulated thread block of
// cta_emul.h - the product's kernel sources, launch by launch and CTA by CTA:
//   k_reduce_big (S given dense: |a|^2 tables + zgebal scaling)  ->  blocked compact-WY reduction on DMMA tiles
//   (bigred_kernels.cuh)  ->  k_hqr_big<false> (windowed multi-bulge QR, one CTA per matrix)  ->  k_hqr_mb cascade
//   ->  k_modes_col (inverse iteration, column oriented)  ->  back-transformation: k_backx_slab (the default while a
//   slab of mode vectors fits an SM) AND the blocked DMMA one (the default at cfg4 / cfg5 sizes)  ->  k_norm_big.
// The thread-block-cluster instance of the windowed QR needs distributed shared memory and is GPU-only.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/big_emul tests/host/big_emul.cpp
//   usage: big_emul <in> <out>
//   in : int32 N, hb, 0, 0, 0 ; S[(2N)^2] complex128 column-major ; Cb[N][2hb+1] ; K6d[N]
//   out: int32 info, launches ; float64 max |psi_blocked - psi_slab| / max |psi| ; scale[2N] ; k[2N] ; psi[2N][2N] (slab path)
// Launch sequences as in axs_launch_reduce_blocked (eig_bigred.cu), axs_launch_hqr_big, axs_launch_modes_big
// (eig_big.cu) and hqr_cascade (eig_small.cu), for one matrix.
#include "shim/cuda_runtime.h"
#include "shim/cuda_pipeline.h"
#include "shim/dmma.h"
#include "../../axisym-safe-python_b200/csrc/big_kernels.cuh"
#include "../../axisym-safe-python_b200/csrc/bigred_kernels.cuh"
#include "../../axisym-safe-python_b200/csrc/hqr_mb.cuh"
#include <vector>

static long g_launches = 0;
#define LAUNCH(gx, gy, gz, nt, ...) do { ++g_launches; cta::run_grid((gx), (gy), (gz), (nt), [&] { __VA_ARGS__; }); } while (0)
static void fits(size_t bytes) { if (bytes > sizeof(cta::g.dyn_smem)) { fprintf(stderr, "shared memory: %zu bytes\n", bytes); exit(7); } }

static int capacity(int per_sm)
{
    int m = 0;
    while (hqr_mb_smem(m + 1) + 2048 <= (size_t)(227 * 1024 / per_sm)) ++m;
    return m;
}
template <int NT, int MINB, int NITEMS, bool OPT>
static void launch_mb(int n, int first, int m_stop, cplx* Hp, cplx* eig, int* info, int* ucur)
{
    LAUNCH(1, 1, 1, NT, k_hqr_mb<NT, MINB, NITEMS, OPT, false>(n, MB_MAX, first, m_stop, 0, Hp, Hp, eig, info, ucur, nullptr));
}
// hqr_cascade(n, nf, m_cur, nullptr, Hp, ...) of eig_small.cu: the tail of the large path
static void cascade(int n, int m_cur, cplx* Hp, cplx* eig, int* info, int* ucur)
{
    const int m3 = capacity(3), m4 = capacity(4);
    if (m_cur <= 35 && MB_MAX * (m_cur + 1) <= 3 * (128 - 32)) { launch_mb<128, 8, 3, false>(n, 0, 0, Hp, eig, info, ucur); return; }
    if (m3 == 0 || m_cur > m3) {
        const int stop = (m3 >= 8 && m3 < m_cur) ? m3 : 0;
        launch_mb<512, 2, 3, false>(n, 0, stop, Hp, eig, info, ucur);
        if (!stop) return;
        m_cur = stop;
    }
    if (m4 == 0 || m_cur > m4) {
        const int stop = (m4 >= 8 && m4 < m_cur) ? m4 : 0;
        launch_mb<384, 3, 3, false>(n, 0, stop, Hp, eig, info, ucur);
        if (!stop) return;
        m_cur = stop;
    }
    launch_mb<256, 4, 3, false>(n, 0, 0, Hp, eig, info, ucur);      // (m5 = 0 without Hc_first: 256 x 4 finishes)
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    int h[5];
    if (fread(h, sizeof(int), 5, fi) != 5) return 4;
    const int N = h[0], hb = h[1], n = 2 * N, W = 2 * hb + 1, nf = 1, CG = 16, nb = BR_NB_MAX;
    if (n <= AXS_MAX_SMALL_N2 || n > 512) { fprintf(stderr, "this harness serves 166 < 2N <= 512\n"); return 8; }
    std::vector<cplx> S((size_t)n * n), Cb((size_t)N * W), K6d(N), eig(n), psi((size_t)n * n);
    if (fread(S.data(), sizeof(cplx), S.size(), fi) != S.size()) return 5;
    if (fread(Cb.data(), sizeof(cplx), Cb.size(), fi) != Cb.size()) return 5;
    if (fread(K6d.data(), sizeof(cplx), K6d.size(), fi) != K6d.size()) return 5;
    fclose(fi);
    std::vector<char> buf((size_t)axs_work_layout_big(n, nf, nullptr, nullptr) + 4096);
    AxsWorkBig wk;
    axs_work_layout_big(n, nf, (void*)(((uintptr_t)buf.data() + 255) & ~(uintptr_t)255), &wk);
    int info = 0;
    // ---- axs_launch_balance_big ------------------------------------------------------------------------------
    fits((size_t)n * sizeof(cplx));
    LAUNCH(nf, 1, 1, RB_NT, k_reduce_big(N, hb, nullptr, Cb.data(), nullptr, nullptr, K6d.data(), nullptr, S.data(), wk, 1));
    // ---- axs_launch_reduce_blocked ----------------------------------------------------------------------------
    for (int k0 = 0; k0 < n - 2; k0 += nb) {
        const int nbc = (nb < n - 2 - k0) ? nb : n - 2 - k0;
        const int rbn = (n - k0 - 1 + BR_GV_NT - 1) / BR_GV_NT;
        for (int j = 0; j < nbc; ++j) {
            LAUNCH(nf, 1, 1, BR_COL_NT, k_br_col(n, k0, j, CG, 0, wk));
            LAUNCH(rbn * CG, nf, 1, BR_GV_NT, k_br_gemv(n, k0, k0 + j, CG, wk));
        }
        LAUNCH(nf, 1, 1, BR_COL_NT, k_br_col(n, k0, nbc, CG, 1, wk));
        const int t0 = k0 + nbc, ntrail = n - t0;
        LAUNCH((k0 + 1 + BR_TM - 1) / BR_TM, 1, nf, 256, k_br_gemm<BR_YTOP>(n, k0, nbc, CG, wk, nullptr));
        LAUNCH((k0 + 1 + 255) / 256, nf, 1, 256, k_br_top(n, k0, nbc, CG, k0 / nb, wk));
        if (ntrail > 0) {
            const int ntn = (ntrail + BR_TN - 1) / BR_TN;
            LAUNCH((n + BR_TM - 1) / BR_TM, ntn, nf, 256, k_br_gemm<BR_RIGHT>(n, k0, nbc, CG, wk, nullptr));
            LAUNCH((ntrail + BR_TM - 1) / BR_TM, 1, nf, 256, k_br_gemm<BR_LEFTW>(n, k0, nbc, CG, wk, nullptr));
            LAUNCH((ntrail + 255) / 256, nf, 1, 256, k_br_tw(n, k0, nbc, CG, wk));
            LAUNCH((n - k0 - 1 + BR_TM - 1) / BR_TM, ntn, nf, 256, k_br_gemm<BR_LEFTA>(n, k0, nbc, CG, wk, nullptr));
        }
    }
    unsigned long long* hmax = (unsigned long long*)wk.msc;
    hmax[0] = 0;
    int gx = 2 * 148;
    if (gx > n) gx = n;
    LAUNCH(gx, nf, 1, 256, k_br_pack(n, wk, hmax));
    LAUNCH((n + 255) / 256 < 64 ? (n + 255) / 256 : 64, nf, 1, 256, k_br_pack_fin(n, wk, hmax));
    // ---- axs_launch_hqr_big (one CTA per matrix) + the shared-memory cascade ------------------------------------
    const int m2 = capacity(2);
    fits(hqr_big_smem());
    LAUNCH(nf, 1, 1, HB_NT, k_hqr_big<false>(n, m2, wk.Hc, eig.data(), &info, wk.ucur));
    cascade(n, m2, wk.Hc, eig.data(), &info, wk.ucur);
    // ---- axs_launch_modes_big -----------------------------------------------------------------------------------
    {
        const int M = 4, groups = (n + M - 1) / M;            // n <= 512: k_modes_col<256, 2, 4>
        const int grid = groups < AXS_BIG_MODE_GRID ? groups : AXS_BIG_MODE_GRID;
        fits((size_t)n * sizeof(cplx));
        LAUNCH(grid, 1, 1, 256, k_modes_col<256, 2, 4>(n, nf, eig.data(), psi.data(), wk));
    }
    std::vector<cplx> X(psi), psi_blocked;
    int Cs = 8;
    while (Cs >= 1 && (size_t)(Cs + 4) * n * sizeof(cplx) > 224 * 1024) Cs >>= 1;
    if (Cs < 1) return 9;
    fits((size_t)(Cs + 4) * n * sizeof(cplx));
    LAUNCH((n + Cs - 1) / Cs, nf, 1, 256, k_backx_slab(n, Cs, psi.data(), wk));
    LAUNCH((n + 31) / 32, nf, 1, 256, k_norm_big(N, hb, Cb.data(), K6d.data(), eig.data(), psi.data()));
    // the blocked back-transformation (axs_launch_backx_blocked) on the same Hessenberg-basis vectors
    {
        psi_blocked = X;
        int last = 0;
        while (last + nb < n - 2) last += nb;
        for (int k0 = last; k0 >= 0; k0 -= nb) {
            const int nbc = (nb < n - 2 - k0) ? nb : n - 2 - k0;
            LAUNCH((n + BR_TM - 1) / BR_TM, 1, nf, 256, k_br_gemm<BR_BXW>(n, k0, nbc, CG, wk, psi_blocked.data()));
            LAUNCH((n + 255) / 256, nf, 1, 256, k_bx_tw(n, nbc, CG, k0 / nb, wk));
            LAUNCH((n - k0 - 1 + BR_TM - 1) / BR_TM, (n + BR_TN - 1) / BR_TN, nf, 256, k_br_gemm<BR_BXA>(n, k0, nbc, CG, wk, psi_blocked.data()));
        }
        LAUNCH(256, nf, 1, 256, k_bx_scale(n, psi_blocked.data(), wk));
        LAUNCH((n + 31) / 32, nf, 1, 256, k_norm_big(N, hb, Cb.data(), K6d.data(), eig.data(), psi_blocked.data()));
    }
    double top = 0.0, dif = 0.0;
    for (size_t q = 0; q < psi.size(); ++q) {
        top = fmax(top, cabs1(psi[q]));
        dif = fmax(dif, cabs1(csub(psi[q], psi_blocked[q])));
    }
    const double rel = dif / top;
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    int o[2] = {info, (int)g_launches};
    fwrite(o, sizeof(int), 2, fo);
    fwrite(&rel, sizeof(double), 1, fo);
    fwrite(wk.scale, sizeof(double), n, fo);
    fwrite(eig.data(), sizeof(cplx), n, fo);
    fwrite(psi.data(), sizeof(cplx), psi.size(), fo);
    fclose(fo);
    return 0;
}
