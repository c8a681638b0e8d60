// TEST INFRASTRUCTURE: kernel 1 of the sweep (csrc/assemble_kernels.cuh: element quadrature, operator
// scatter, dense S(w) batch) compiled as plain C++ and run on the emulated thread block of cta_emul.h.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/assemble_emul tests/host/assemble_emul.cpp
//   usage: assemble_emul <in> <out>
//   in : int32 n_elem, n_tables, n_nodes, n_out, n_lm, N, hb, n_circ, nf, has_K1c ;
//        edesc_i[n_elem][8] int32 ; edesc_d[n_elem][12] f64 ; tables f64 ; nodes f64 ; lm int32 ; lm_off[n_elem] int32 ;
//        tclass[N] int32 ; K1c[N][2hb+1] complex128 (if has_K1c) ; omega[nf] f64
//   out: element matrices [n_out] complex128 ; K0s, Cb, Mb [N][2hb+1] ; K6d[N] ; S[nf][(2N)^2] column-major
// Launch geometry as in axs_launch_element_matrices / axs_launch_scatter / axs_launch_build_S (assemble.cu).
#include "shim/cuda_runtime.h"
#include "../../axisym-safe-python_b200/csrc/assemble_kernels.cuh"
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
    int h[10];
    if (fread(h, sizeof(int), 10, fi) != 10) return 4;
    const int n_elem = h[0], n_tables = h[1], n_nodes = h[2], n_out = h[3], n_lm = h[4], N = h[5], hb = h[6], n_circ = h[7],
              nf = h[8], has_k1c = h[9], W = 2 * hb + 1, n2 = 2 * N;
    auto edesc_i = rd<int>(fi, (size_t)n_elem * 8);
    auto edesc_d = rd<double>(fi, (size_t)n_elem * 12);
    auto tables = rd<double>(fi, n_tables);
    auto nodes = rd<double>(fi, n_nodes);
    auto lm = rd<int>(fi, n_lm);
    auto lm_off = rd<int>(fi, n_elem);
    auto tclass = rd<int>(fi, N);
    auto K1c = rd<cplx>(fi, has_k1c ? (size_t)N * W : 0);
    auto omega = rd<double>(fi, nf);
    fclose(fi);
    std::vector<cplx> out(n_out), K0s((size_t)N * W), Cb((size_t)N * W), Mb((size_t)N * W), K6d(N), S((size_t)nf * n2 * n2);
    const cplx zero = mk(0, 0);
    std::fill(out.begin(), out.end(), zero); std::fill(K0s.begin(), K0s.end(), zero);
    std::fill(Cb.begin(), Cb.end(), zero); std::fill(Mb.begin(), Mb.end(), zero); std::fill(K6d.begin(), K6d.end(), zero);
    const int warps = 4, grid = (n_elem + warps - 1) / warps;
    const size_t smem = warps * ((size_t)EL_MAX_M * EL_MAX_M * sizeof(double) + EL_MAX_M * sizeof(PointGeom) + EL_MAX_M * sizeof(double));
    if (smem > sizeof(cta::g.dyn_smem)) return 7;
    for (int b = 0; b < grid; ++b)
        cta::run(warps * 32, b, [&] { k_element_matrices(n_elem, edesc_i.data(), edesc_d.data(), tables.data(), nodes.data(), out.data()); }, 0, grid, 1);
    for (int b = 0; b < grid; ++b)
        cta::run(warps * 32, b, [&] { k_scatter_operators(n_elem, edesc_i.data(), out.data(), lm.data(), lm_off.data(), tclass.data(),
                                                          n_circ, N, hb, K0s.data(), Cb.data(), Mb.data(), K6d.data()); }, 0, grid, 1);
    int bx = (n2 * n2 + 255) / 256;
    if (bx > 64) bx = 64;
    for (int f = 0; f < nf; ++f)
        for (int b = 0; b < bx; ++b)
            cta::run(256, b, [&] { k_build_S(N, hb, K0s.data(), Cb.data(), Mb.data(), has_k1c ? K1c.data() : nullptr, K6d.data(),
                                             omega.data(), S.data()); }, f, bx, nf);
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    fwrite(out.data(), sizeof(cplx), out.size(), fo);
    fwrite(K0s.data(), sizeof(cplx), K0s.size(), fo);
    fwrite(Cb.data(), sizeof(cplx), Cb.size(), fo);
    fwrite(Mb.data(), sizeof(cplx), Mb.size(), fo);
    fwrite(K6d.data(), sizeof(cplx), K6d.size(), fo);
    fwrite(S.data(), sizeof(cplx), S.size(), fo);
    fclose(fo);
    return 0;
}
