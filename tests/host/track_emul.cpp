// TEST INFRASTRUCTURE: the tracking kernels (csrc/track_kernels.cuh: Y = B psi, the DMMA contraction
// P = psi_prev^T Y with its fused row arg-max, the column gather) on the emulated thread block of cta_emul.h.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/track_emul tests/host/track_emul.cpp
//   usage: track_emul <in> <out>
//   in : int32 N, hb, nf, tiled (1 = k_bpsi_tiled as the library chooses, 0 = k_bpsi), 0 ;
//        Cb[N][2hb+1] complex128 ; K6d[N] complex128 ; psi[nf][2N][2N] complex128 ; order[nf][2N] int32 ; k[nf][2N] complex128
//   out: arg[nf-1][2N] int32 ; amax[nf-1][2N] f64 ; k_sorted[nf][2N], psi_sorted[nf][2N][2N] complex128
// Launch geometry as in axs_launch_track / axs_launch_apply_order (track.cu).
#include "shim/cuda_runtime.h"
#include "shim/dmma.h"
#include "../../axisym-safe-python_b200/csrc/track_kernels.cuh"
#include <vector>

template <class T> static std::vector<T> rd(FILE* f, size_t n)
{
    std::vector<T> v(n ? n : 1);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(5); }
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    int h[5];
    if (fread(h, sizeof(int), 5, fi) != 5) return 4;
    const int N = h[0], hb = h[1], nf = h[2], tiled = h[3], n = 2 * N, W = 2 * hb + 1, npairs = nf - 1;
    auto Cb = rd<cplx>(fi, (size_t)N * W);
    auto K6d = rd<cplx>(fi, N);
    auto psi = rd<cplx>(fi, (size_t)nf * n * n);
    auto order = rd<int>(fi, (size_t)nf * n);
    auto k = rd<cplx>(fi, (size_t)nf * n);
    fclose(fi);
    std::vector<cplx> Y((size_t)nf * n * n), k_out((size_t)nf * n), psi_out((size_t)nf * n * n);
    std::vector<int> arg((size_t)npairs * n);
    std::vector<double> amax((size_t)npairs * n);
    // ---- axs_launch_track ----------------------------------------------------------------------------
    const size_t ubytes = (size_t)N * BP_COLS * sizeof(cplx), budget = (size_t)110 * 1024;
    int crows = ubytes < budget ? (int)((budget - ubytes) / ((size_t)W * sizeof(cplx))) : 0;
    if (crows > N) crows = N;
    if (crows >= 8 && tiled) {
        const int gx = (n + BP_COLS - 1) / BP_COLS;
        for (int f = 0; f < nf; ++f)
            for (int b = 0; b < gx; ++b)
                cta::run(256, b, [&] { k_bpsi_tiled(N, hb, crows, Cb.data(), K6d.data(), psi.data(), Y.data()); }, f, gx, nf);
    } else {
        int bx = (n * n + 255) / 256;
        if (bx > 64) bx = 64;
        for (int f = 0; f < nf; ++f)
            for (int b = 0; b < bx; ++b)
                cta::run(256, b, [&] { k_bpsi(N, hb, Cb.data(), K6d.data(), psi.data(), Y.data()); }, f, bx, nf);
    }
    const int gx = (n + 31) / 32;
    for (int p = 0; p < npairs; ++p)
        for (int b = 0; b < gx; ++b)
            cta::run(256, b, [&] { k_track_mma(n, psi.data(), Y.data(), arg.data(), amax.data()); }, p, gx, npairs);
    // ---- axs_launch_apply_order ------------------------------------------------------------------------
    {
        // geometry of the library: see axs_launch_apply_order (grid.x blocks of 256 threads over the n x n block, grid.y = nf)
        int bx = (n * n + 255) / 256;
        if (bx > 64) bx = 64;
        for (int f = 0; f < nf; ++f)
            for (int b = 0; b < bx; ++b)
                cta::run(256, b, [&] { k_apply_order(n, order.data(), k.data(), psi.data(), k_out.data(), psi_out.data()); }, f, bx, nf);
    }
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    fwrite(arg.data(), sizeof(int), arg.size(), fo);
    fwrite(amax.data(), sizeof(double), amax.size(), fo);
    fwrite(k_out.data(), sizeof(cplx), k_out.size(), fo);
    fwrite(psi_out.data(), sizeof(cplx), psi_out.size(), fo);
    fclose(fo);
    return 0;
}
