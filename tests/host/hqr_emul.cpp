// TEST INFRASTRUCTURE: the multi-bulge QR kernel k_hqr_mb (csrc/hqr_mb.cuh, the kernel that takes 60 % of
// the sweep) compiled as plain C++ and run on the emulated thread block of cta_emul.h - every instance of
// the shrinking-window cascade, with and without the early-deflation window, on real SAFE matrices,
// without a GPU.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/hqr_emul tests/host/hqr_emul.cpp
//   usage: hqr_emul <in> <out>
//   in : int32 n, aed window (0 = off), aed stage mask, bulges, 0 ; n*n complex128 column-major (upper Hessenberg)
//   out: int32 info, macro steps, batches, AED calls, block barriers, launches ; n complex128 eigenvalues
// The launch logic below mirrors axs_launch_hqr / hqr_cascade of csrc/eig_small.cu (instance per window
// size, stage capacities from the shared-memory footprint); the kernel itself is the product source.
#include "shim/cuda_runtime.h"
#include "../../axisym-safe-python_b200/csrc/hqr_mb.cuh"
#include <vector>

static int g_aed_window = 0, g_mask = 15, g_launches = 0;

static int aed_of(int per_sm)
{
    const int bit = per_sm >= 4 ? 3 : per_sm - 1;
    return (g_aed_window >= 2 && ((g_mask >> bit) & 1)) ? (g_aed_window > AED_MAX ? AED_MAX : g_aed_window) : 0;
}
static int capacity(int per_sm)
{
    const size_t slack = aed_of(per_sm) >= 2 ? 5120 : 2048;
    int m = 0;
    while (hqr_mb_smem(m + 1) + slack <= (size_t)(227 * 1024 / per_sm)) ++m;
    return m;
}

template <int NT, int MINB, int NITEMS, bool OPT>
static void launch(int n, int bmax, int first, int m_stop, const cplx* Hc, cplx* Hp, cplx* eig, int* info, int* ucur, int* stats)
{
    const int aed = aed_of(MINB);
    if (hqr_mb_smem(first ? n : ucur[0] + 1) > sizeof(cta::g.dyn_smem)) { fprintf(stderr, "window does not fit\n"); exit(7); }
    ++g_launches;
    if (aed >= 2)
        cta::run(NT, 0, [&] { k_hqr_mb<NT, MINB, NITEMS, OPT, true>(n, bmax, first, m_stop, aed, Hc, Hp, eig, info, ucur, stats); });
    else
        cta::run(NT, 0, [&] { k_hqr_mb<NT, MINB, NITEMS, OPT, false>(n, bmax, first, m_stop, 0, Hc, Hp, eig, info, ucur, stats); });
}

// hqr_cascade of eig_small.cu (4 stages)
static void cascade(int n, int m_cur, const cplx* Hc_first, cplx* Hp, cplx* eig, int* info, int* ucur, int* stats, int bulges)
{
    const int m3 = capacity(3), m4 = capacity(4), m5 = Hc_first ? capacity(5) : 0;
    int first = Hc_first ? 1 : 0;
    const cplx* src = Hc_first ? Hc_first : Hp;
    if (m_cur <= 35 && bulges * (m_cur + 1) <= 3 * (128 - 32)) {
        launch<128, 8, 3, false>(n, bulges, first, 0, src, Hp, eig, info, ucur, stats);
        return;
    }
    if (m3 == 0 || m_cur > m3) {
        const int stop = (m3 >= 8 && m3 < m_cur) ? m3 : 0;
        launch<512, 2, 3, false>(n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (!stop) return;
        first = 0; src = Hp; m_cur = stop;
    }
    if (m4 == 0 || m_cur > m4) {
        const int stop = (m4 >= 8 && m4 < m_cur) ? m4 : 0;
        launch<384, 3, 3, false>(n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (!stop) return;
        first = 0; src = Hp; m_cur = stop;
    }
    if (!(m5 >= 8 && m_cur <= m5)) {
        const int stop = (m5 >= 8 && m5 < m_cur) ? m5 : 0;
        launch<256, 4, 3, false>(n, bulges, first, stop, src, Hp, eig, info, ucur, stats);
        if (!stop) return;
        first = 0; src = Hp; m_cur = stop;
    }
    launch<256, 5, 3, false>(n, bulges, first, 0, src, Hp, eig, info, ucur, stats);
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: hqr_emul in out\n"); return 2; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    int hdr[5];
    if (fread(hdr, sizeof(int), 5, fi) != 5) return 4;
    const int n = hdr[0];
    g_aed_window = hdr[1]; g_mask = hdr[2];
    int bulges = hdr[3];
    if (bulges < 1 || bulges > MB_MAX) bulges = MB_MAX;
    std::vector<cplx> dense((size_t)n * n);
    if (fread(dense.data(), sizeof(cplx), (size_t)n * n, fi) != (size_t)n * n) return 5;
    fclose(fi);
    std::vector<cplx> Hc(hcol_size(n) + 2), Hp(hcol_size(n) + 2), eig(n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i <= j + 1 && i < n; ++i) Hc[hcol_off(j) + i] = dense[(size_t)j * n + i];
    int info = 0, ucur = 0, stats[AXS_QR_STATS];
    memset(stats, 0, sizeof stats);
    // axs_launch_hqr of eig_small.cu
    int m2 = capacity(2);
    if (m2 >= n || m2 < 8) m2 = 0;
    if (m2 == 0)
        cascade(n, n, Hc.data(), Hp.data(), eig.data(), &info, &ucur, stats, bulges);
    else {
        if (bulges * n <= 3 * 384) launch<512, 1, 3, true>(n, bulges, 1, m2, Hc.data(), Hp.data(), eig.data(), &info, &ucur, stats);
        else launch<512, 1, 3, false>(n, bulges, 1, m2, Hc.data(), Hp.data(), eig.data(), &info, &ucur, stats);
        cascade(n, m2, nullptr, Hp.data(), eig.data(), &info, &ucur, stats, bulges);
    }
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    int o[6] = {info, stats[0], stats[1], stats[2], (int)cta::g.barriers, g_launches};
    fwrite(o, sizeof(int), 6, fo);
    fwrite(eig.data(), sizeof(cplx), n, fo);
    fclose(fo);
    return 0;
}
