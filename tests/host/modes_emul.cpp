// TEST INFRASTRUCTURE: k_modes (csrc/modes.cuh: inverse iteration on the Hessenberg matrix, 15 % of the
// sweep) compiled as plain C++ and run on the emulated thread block of cta_emul.h, without a GPU.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/modes_emul tests/host/modes_emul.cpp
//   usage: modes_emul <in> <out>
//   in : int32 n, warps per CTA, 0, 0, 0 ; n*n complex128 column-major (upper Hessenberg) ; n complex128 eigenvalues
//   out: n*n complex128 row-major X (X[i][mode], the kernel's hand-over layout: unit max-norm vectors of H)
// The launch mirrors axs_launch_modes of csrc/eig_small.cu (grid = (ceil(n / (8 nw)), nf), block = 32 nw).
#include "shim/cuda_runtime.h"
#include "shim/cuda_pipeline.h"
#include "../../axisym-safe-python_b200/csrc/modes.cuh"
#include <vector>

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: modes_emul in out\n"); return 2; }
    FILE* fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    int hdr[5];
    if (fread(hdr, sizeof(int), 5, fi) != 5) return 4;
    const int n = hdr[0];
    int nw = hdr[1] > 0 ? hdr[1] : 4;
    std::vector<cplx> dense((size_t)n * n), eig(n), X((size_t)n * n);
    if (fread(dense.data(), sizeof(cplx), (size_t)n * n, fi) != (size_t)n * n) return 5;
    if (fread(eig.data(), sizeof(cplx), n, fi) != (size_t)n) return 5;
    fclose(fi);
    std::vector<cplx> Hr(hrow_size(n) + 2);
    for (int i = 0; i < n; ++i)
        for (int j = hrow_first(i); j < n; ++j) Hr[hrow_off(i, n) + j - hrow_first(i)] = dense[(size_t)j * n + i];
    while (nw > 1 && modes_smem(n, nw) > 227 * 1024) --nw;
    if (modes_smem(n, nw) > sizeof(cta::g.dyn_smem)) return 7;
    AxsWork wk;
    memset(&wk, 0, sizeof wk);
    wk.Hr = Hr.data();
    const int N = n / 2, nblocks = (n + 8 * nw - 1) / (8 * nw);
    for (int b = 0; b < nblocks; ++b)
        cta::run(32 * nw, b, [&] { k_modes(N, 0, nullptr, nullptr, eig.data(), X.data(), wk); }, 0);
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    fwrite(X.data(), sizeof(cplx), (size_t)n * n, fo);
    fclose(fo);
    return 0;
}
