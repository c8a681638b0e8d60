// TEST INFRASTRUCTURE: the kernels behind energy_ratio / energy_velocity / point_excited_waves and the subset
// mode of solve() (csrc/post_kernels.cuh) on the emulated thread block of cta_emul.h.
//   build: g++ -std=c++17 -O2 -I tests/host/shim -o tests/host/post_emul tests/host/post_emul.cpp
//   usage: post_emul <in> <out>
//   in : int32 N, nf, nnz_t, nnz_p, n_sel, n_mat, nnz_m, m ;
//        psi[nf][2N][2N] complex128 ; tclass[N] int32 ; (rt, ct int32 [nnz_t], vt complex128 [nnz_t]) ; (rp, cp, vp) ;
//        sel[n_sel] int32 ; mat_off[n_mat+1] int32 ; rows, cols int32 [nnz_m] ; vals complex128 [nnz_m] ;
//        fvec[2N] complex128 ; k[nf][2N] complex128 ; sigma[nf] f64
//   out: er[nf][2N] f64 ; forms[nf][n_sel][n_mat] complex128 ; proj[nf][2N] complex128 ; rank[nf][2N] int32 ;
//        psi with the columns of rank >= m zeroed
// Launch geometry as in axs_launch_energy_ratio / _quad_forms / _modal_projection / _select_modes (track.cu).
#include "shim/cuda_runtime.h"
#include "../../axisym-safe-python_b200/csrc/post_kernels.cuh"
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
    int h[8];
    if (fread(h, sizeof(int), 8, fi) != 8) return 4;
    const int N = h[0], nf = h[1], nnz_t = h[2], nnz_p = h[3], n_sel = h[4], n_mat = h[5], nnz_m = h[6], m = h[7], n = 2 * N;
    auto psi = rd<cplx>(fi, (size_t)nf * n * n);
    auto tclass = rd<int>(fi, N);
    auto rt = rd<int>(fi, nnz_t); auto ct = rd<int>(fi, nnz_t); auto vt = rd<cplx>(fi, nnz_t);
    auto rp = rd<int>(fi, nnz_p); auto cp = rd<int>(fi, nnz_p); auto vp = rd<cplx>(fi, nnz_p);
    auto sel = rd<int>(fi, n_sel);
    auto mat_off = rd<int>(fi, n_mat + 1);
    auto rows = rd<int>(fi, nnz_m); auto cols = rd<int>(fi, nnz_m); auto vals = rd<cplx>(fi, nnz_m);
    auto fvec = rd<cplx>(fi, n);
    auto k = rd<cplx>(fi, (size_t)nf * n);
    auto sigma = rd<double>(fi, nf);
    fclose(fi);
    std::vector<double> er((size_t)nf * n);
    std::vector<cplx> forms((size_t)nf * n_sel * n_mat), proj((size_t)nf * n);
    std::vector<int> rank((size_t)nf * n);
    cta::run_grid((n + 127) / 128, nf, 1, 128, [&] { k_energy_ratio(N, psi.data(), tclass.data(), nnz_t, rt.data(), ct.data(), vt.data(),
                                                                    nnz_p, rp.data(), cp.data(), vp.data(), er.data()); });
    cta::run_grid((n_sel + 127) / 128, nf, 1, 128, [&] { k_quad_forms(N, psi.data(), tclass.data(), n_sel, sel.data(), n_mat, mat_off.data(),
                                                                      rows.data(), cols.data(), vals.data(), forms.data()); });
    cta::run_grid((n + 127) / 128, nf, 1, 128, [&] { k_modal_projection(n, psi.data(), fvec.data(), proj.data()); });
    cta::run_grid(nf, 1, 1, 256, [&] { k_rank_modes(n, k.data(), sigma.data(), rank.data()); });
    if (m < n) {
        long long cells = (long long)n * n;
        int bx = (int)((cells + 255) / 256);
        if (bx > 64) bx = 64;
        cta::run_grid(bx, nf, 1, 256, [&] { k_zero_unselected(n, m, rank.data(), psi.data()); });
    }
    FILE* fo = fopen(argv[2], "wb");
    if (!fo) return 6;
    fwrite(er.data(), sizeof(double), er.size(), fo);
    fwrite(forms.data(), sizeof(cplx), forms.size(), fo);
    fwrite(proj.data(), sizeof(cplx), proj.size(), fo);
    fwrite(rank.data(), sizeof(int), rank.size(), fo);
    fwrite(psi.data(), sizeof(cplx), psi.size(), fo);
    fclose(fo);
    return 0;
}
