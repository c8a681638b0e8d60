// TEST INFRASTRUCTURE: the blocked compact-WY Hessenberg reduction with DMMA trailing updates and the blocked
// back-transformation (csrc/bigred_kernels.cuh: the large-path instance of kernel 2, north-star kernel (2)) on
// the emulated thread block of cta_emul.h, launch by launch and CTA by CTA.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/bigred_emul tests/host/bigred_emul.cpp
//   usage: bigred_emul <in> <out>
//   in : int32 n, panel width, 0, 0, 0 ; A[n*n] complex128 column-major (the balanced matrix)
//   out: int32 launches, CTAs ; A (Hessenberg + reflectors below the subdiagonal), tau[n], Q[n][n] row-major = the
//        blocked back-transformation applied to the identity (scale = 1), packed Hc [hcol_size(n)]
// The launch sequences mirror axs_launch_reduce_blocked / axs_launch_backx_blocked (eig_bigred.cu) for one matrix;
// the balancing kernel in front of them (k_reduce_big, eig_big.cu) is not part of this harness.
#include "shim/cuda_runtime.h"
#include "shim/dmma.h"
#include "../../axisym-safe-python_b200/csrc/bigred_kernels.cuh"
#include <vector>

static long g_launches = 0, g_ctas = 0;
#define LAUNCH(gx, gy, gz, nt, ...) do { ++g_launches; g_ctas += (long)(gx) * (gy) * (gz); cta::run_grid((gx), (gy), (gz), (nt), [&] { __VA_ARGS__; }); } while (0)

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    int h[5];
    if (fread(h, sizeof(int), 5, fi) != 5) return 4;
    const int n = h[0], nf = 1;
    int nb = h[1] > 0 ? h[1] : 32;
    if (nb < 4) nb = 4;
    if (nb > BR_NB_MAX) nb = BR_NB_MAX;
    std::vector<cplx> A0((size_t)n * n);
    if (fread(A0.data(), sizeof(cplx), A0.size(), fi) != A0.size()) return 5;
    fclose(fi);
    std::vector<char> buf((size_t)axs_work_layout_big(n, nf, nullptr, nullptr) + 4096);
    AxsWorkBig wk;
    axs_work_layout_big(n, nf, (void*)(((uintptr_t)buf.data() + 255) & ~(uintptr_t)255), &wk);
    memcpy(wk.A, A0.data(), A0.size() * sizeof(cplx));
    for (int i = 0; i < n; ++i) { wk.scale[i] = 1.0; wk.tau[i] = mk(0, 0); }
    // ---- axs_launch_reduce_blocked -----------------------------------------------------------------------
    const size_t col_smem = (size_t)n * sizeof(cplx);
    if (col_smem > sizeof(cta::g.dyn_smem)) return 7;
    const int t_rows = (BR_NB_MAX * BR_NB_MAX + n - 1) / n, CG = 16;
    if (2 * BR_NB_MAX + CG + t_rows > AXS_BIG_YW) { fprintf(stderr, "scratch rows\n"); return 8; }
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
        LAUNCH((k0 + 1 + 255) / 256, nf, 1, 256, k_br_top(n, k0, nbc, CG, nb == BR_NB_MAX ? k0 / nb : -1, wk));
        if (ntrail > 0) {
            const int ntn = (ntrail + BR_TN - 1) / BR_TN;
            LAUNCH((n + BR_TM - 1) / BR_TM, ntn, nf, 256, k_br_gemm<BR_RIGHT>(n, k0, nbc, CG, wk, nullptr));
            LAUNCH((ntrail + BR_TM - 1) / BR_TM, 1, nf, 256, k_br_gemm<BR_LEFTW>(n, k0, nbc, CG, wk, nullptr));
            LAUNCH((ntrail + 255) / 256, nf, 1, 256, k_br_tw(n, k0, nbc, CG, wk));
            LAUNCH((n - k0 - 1 + BR_TM - 1) / BR_TM, ntn, nf, 256, k_br_gemm<BR_LEFTA>(n, k0, nbc, CG, wk, nullptr));
        }
    }
    std::vector<cplx> Aout(wk.A, wk.A + (size_t)n * n);      // (k_br_pack_fin reuses the scratch region, not A)
    unsigned long long* hmax = (unsigned long long*)wk.msc;
    hmax[0] = 0;
    int gx = 2 * 148;
    if (gx > n) gx = n;
    LAUNCH(gx, nf, 1, 256, k_br_pack(n, wk, hmax));
    LAUNCH((n + 255) / 256 < 64 ? (n + 255) / 256 : 64, nf, 1, 256, k_br_pack_fin(n, wk, hmax));
    std::vector<cplx> Hc(wk.Hc, wk.Hc + hcol_size(n));
    // ---- axs_launch_backx_blocked on X = I (needs panels of BR_NB_MAX: the T factors are kept only then) ----
    std::vector<cplx> X((size_t)n * n, mk(0, 0));
    for (int i = 0; i < n; ++i) X[(size_t)i * n + i] = mk(1, 0);
    int have_q = 0;
    if (nb == BR_NB_MAX) {
        have_q = 1;
        int last = 0;
        while (last + nb < n - 2) last += nb;
        for (int k0 = last; k0 >= 0; k0 -= nb) {
            const int nbc = (nb < n - 2 - k0) ? nb : n - 2 - k0;
            LAUNCH((n + BR_TM - 1) / BR_TM, 1, nf, 256, k_br_gemm<BR_BXW>(n, k0, nbc, CG, wk, X.data()));
            LAUNCH((n + 255) / 256, nf, 1, 256, k_bx_tw(n, nbc, CG, k0 / nb, wk));
            LAUNCH((n - k0 - 1 + BR_TM - 1) / BR_TM, (n + BR_TN - 1) / BR_TN, nf, 256, k_br_gemm<BR_BXA>(n, k0, nbc, CG, wk, X.data()));
        }
        LAUNCH(256, nf, 1, 256, k_bx_scale(n, X.data(), wk));
    }
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    int o[4] = {(int)g_launches, (int)g_ctas, have_q, 0};
    fwrite(o, sizeof(int), 4, fo);
    fwrite(Aout.data(), sizeof(cplx), Aout.size(), fo);
    fwrite(wk.tau, sizeof(cplx), n, fo);
    fwrite(X.data(), sizeof(cplx), X.size(), fo);
    fwrite(Hc.data(), sizeof(cplx), Hc.size(), fo);
    fclose(fo);
    return 0;
}
