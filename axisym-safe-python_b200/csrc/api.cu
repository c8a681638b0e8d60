// extern "C" entry points of libaxisafe_b200.so (include/axisafe_b200.h).
#include "common.cuh"
#include "../../include/axisafe_b200.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

extern "C" {
int axs_launch_element_matrices(int, const int*, const double*, const double*, const double*, cplx*, cudaStream_t);
int axs_launch_scatter(int, const int*, const cplx*, const int*, const int*, const int*, int, int, int,
                       cplx*, cplx*, cplx*, cplx*, cudaStream_t);
int axs_launch_build_S(int, int, const cplx*, const cplx*, const cplx*, const cplx*, const cplx*,
                       const double*, int, cplx*, cudaStream_t);
int axs_launch_reduce(int, int, const cplx*, const cplx*, const cplx*, const cplx*, const cplx*,
                      const double*, int, const cplx*, AxsWork, cudaStream_t);
int axs_launch_hqr(int, int, cplx*, int*, AxsWork, cudaStream_t);
int axs_launch_modes(int, int, const cplx*, const cplx*, int, const cplx*, cplx*, AxsWork, cudaStream_t);
int axs_launch_unpack(int, int, AxsWork, cplx*, cudaStream_t);
int axs_launch_track(int, int, const cplx*, const cplx*, const cplx*, int, cplx*, int*, double*, cudaStream_t);
int axs_launch_apply_order(int, int, const int*, const cplx*, const cplx*, cplx*, cplx*, cudaStream_t);
int axs_launch_select_modes(int, int, const cplx*, const double*, int, int*, cplx*, cudaStream_t);
int axs_launch_reduce_big(int, int, const cplx*, const cplx*, const cplx*, const cplx*, const cplx*,
                          const double*, int, const cplx*, AxsWorkBig, cudaStream_t);
int axs_launch_hqr_big(int, int, cplx*, int*, AxsWorkBig, cudaStream_t);
int axs_launch_modes_big(int, int, const cplx*, const cplx*, int, const cplx*, cplx*, AxsWorkBig, cudaStream_t);
int axs_launch_unpack_big(int, int, AxsWorkBig, cplx*, cudaStream_t);
int axs_launch_quad_forms(int, int, const cplx*, const int*, int, const int*, int, const int*, const int*, const int*,
                          const cplx*, cplx*, cudaStream_t);
int axs_launch_modal_projection(int, int, const cplx*, const cplx*, cplx*, cudaStream_t);
int axs_launch_energy_ratio(int, int, const cplx*, const int*, int, const int*, const int*, const cplx*,
                            int, const int*, const int*, const cplx*, double*, cudaStream_t);
int axs_launch_copy_to_pinned(const void*, void*, long long, cudaStream_t);
}

#define AXS_ERR_TOO_LARGE 100001
#define AXS_GRID_Y 32768                /* frequencies per launch of the kernels that put nf in gridDim.y */

// ---- kernel-selection switches: environment read once, then axs_set_tuning ----------------------
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
static AxsTuning* tuning_mut(void)
{
    static AxsTuning t;
    static bool init = false;
    if (!init) {
        t.reduce_oc = env_int("AXS_REDUCE_OC", 12);
        t.reduce_prebal = env_int("AXS_REDUCE_PREBAL", 1);
        t.hqr_stages = env_int("AXS_HQR_STAGES", 4);
        t.hqr_bulges = env_int("AXS_HQR_BULGES", 8);
        t.hqr_small_direct = env_int("AXS_HQR_SMALL_DIRECT", 1);
        t.hqr_aed = env_int("AXS_HQR_AED", 0);      // measured on B200: halves the macro steps, but the serial window costs more than it saves
        t.hqr_aed_reg = env_int("AXS_HQR_AED_REG", 1);
        t.hqr_aed_stages = env_int("AXS_HQR_AED_STAGES", 15);
        t.backnorm_wy = env_int("AXS_BACKNORM_WY", 1);
        t.backnorm_wy_modes = env_int("AXS_BACKNORM_WY_MODES", 64);
        t.bpsi_tiled = env_int("AXS_BPSI_TILED", 1);
        t.big_blocked = env_int("AXS_BIG_BLOCKED", 1);
        t.big_panel = env_int("AXS_BIG_PANEL", 32);
        t.big_qr_cluster = env_int("AXS_BIG_QR_CLUSTER", 0);
        init = true;
    }
    return &t;
}
extern "C" const AxsTuning* axs_tuning(void) { return tuning_mut(); }

static int* tuning_field(const char* name)
{
    AxsTuning* t = tuning_mut();
    if (!name) return nullptr;
    if (!strcmp(name, "reduce_oc")) return &t->reduce_oc;
    if (!strcmp(name, "reduce_prebal")) return &t->reduce_prebal;
    if (!strcmp(name, "hqr_stages")) return &t->hqr_stages;
    if (!strcmp(name, "hqr_bulges")) return &t->hqr_bulges;
    if (!strcmp(name, "hqr_small_direct")) return &t->hqr_small_direct;
    if (!strcmp(name, "hqr_aed")) return &t->hqr_aed;
    if (!strcmp(name, "hqr_aed_reg")) return &t->hqr_aed_reg;
    if (!strcmp(name, "hqr_aed_stages")) return &t->hqr_aed_stages;
    if (!strcmp(name, "backnorm_wy")) return &t->backnorm_wy;
    if (!strcmp(name, "backnorm_wy_modes")) return &t->backnorm_wy_modes;
    if (!strcmp(name, "bpsi_tiled")) return &t->bpsi_tiled;
    if (!strcmp(name, "big_blocked")) return &t->big_blocked;
    if (!strcmp(name, "big_panel")) return &t->big_panel;
    if (!strcmp(name, "big_qr_cluster")) return &t->big_qr_cluster;
    return nullptr;
}
extern "C" int axs_set_tuning(const char* name, int value)
{
    int* p = tuning_field(name);
    if (!p) return -1;
    *p = value;
    return 0;
}
extern "C" int axs_get_tuning(const char* name, int* value)
{
    int* p = tuning_field(name);
    if (!p) return -1;
    if (!value) return -2;
    *value = *p;
    return 0;
}

// ---- dynamic shared-memory opt-in, once per (kernel, device) ---------------------------------------
#include <mutex>
extern "C" int axs_smem_once(const void* kernel, size_t bytes)
{
    struct Entry { const void* k; int dev; size_t bytes; };
    static Entry table[256];
    static int used = 0;
    static std::mutex mu;
    int dev = 0;
    AXS_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    Entry* e = nullptr;
    for (int i = 0; i < used; ++i)
        if (table[i].k == kernel && table[i].dev == dev) { e = &table[i]; break; }
    if (e && e->bytes >= bytes) return 0;
    AXS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (!e && used < 256) { e = &table[used++]; e->k = kernel; e->dev = dev; }
    if (e) e->bytes = bytes;
    return 0;
}

// workspace views advanced by f0 frequencies (launch chunking)
static AxsWork work_at(AxsWork w, int n2, long long f0)
{
    w.A += f0 * n2 * n2; w.Hc += f0 * hcol_size(n2); w.Hr += f0 * hrow_size(n2); w.tau += f0 * n2;
    w.scale += f0 * n2; w.ucur += f0; w.Hp += f0 * hcol_size(n2); w.stats += AXS_QR_STATS * f0;
    return w;
}
static AxsWorkBig workbig_at(AxsWorkBig w, int n2, long long f0)
{
    w.A += f0 * n2 * n2; w.Hc += f0 * hcol_size(n2); w.Hm += f0 * hcol_size(n2); w.tau += f0 * n2;
    w.scale += f0 * n2; w.ucur += f0; w.yw += f0 * AXS_BIG_YW * n2;
    return w;
}

extern "C" int axs_version(void) { return 100; }

extern "C" const char* axs_strerror(int code)
{
    static __thread char buf[96];
    if (code == 0) return "success";
    if (code == AXS_ERR_TOO_LARGE) return "problem size 2N exceeds the largest supported size (axs_max_n2)";
    if (code < 0) { snprintf(buf, sizeof buf, "invalid argument %d", -code); return buf; }
    return cudaGetErrorString((cudaError_t)code);
}

extern "C" int axs_max_small_n2(void) { return AXS_MAX_SMALL_N2; }
extern "C" int axs_max_n2(void) { return AXS_MAX_BIG_N2; }
static inline bool is_big(int N) { return 2 * N > AXS_MAX_SMALL_N2; }

extern "C" int axs_element_matrices(int n_elem, const int* edesc_i, const double* edesc_d,
                                    const double* tables, const double* nodes, axs_cplx* out, void* stream)
{
    if (n_elem <= 0) return -1;
    if (!edesc_i) return -2;
    if (!edesc_d) return -3;
    if (!tables) return -4;
    if (!nodes) return -5;
    if (!out) return -6;
    return axs_launch_element_matrices(n_elem, edesc_i, edesc_d, tables, nodes, (cplx*)out, (cudaStream_t)stream);
}

extern "C" int axs_assemble_operators(int n_elem, const int* edesc_i, const axs_cplx* elem, const int* lm,
                                      const int* lm_off, const int* tclass, int n_circ, int N, int hb,
                                      axs_cplx* K0s, axs_cplx* Cb, axs_cplx* Mb, axs_cplx* K6d, void* stream)
{
    if (n_elem <= 0) return -1;
    if (!edesc_i) return -2;
    if (!elem) return -3;
    if (!lm) return -4;
    if (!lm_off) return -5;
    if (!tclass) return -6;
    if (n_circ < 0) return -7;
    if (N <= 0) return -8;
    if (hb < 0) return -9;
    if (!K0s) return -10;
    if (!Cb) return -11;
    if (!Mb) return -12;
    if (!K6d) return -13;
    return axs_launch_scatter(n_elem, edesc_i, (const cplx*)elem, lm, lm_off, tclass, n_circ, N, hb,
                              (cplx*)K0s, (cplx*)Cb, (cplx*)Mb, (cplx*)K6d, (cudaStream_t)stream);
}

static int check_ops(int N, int hb, const void* K0s, const void* Cb, const void* Mb, const void* K6d)
{
    if (N <= 0) return -1;
    if (hb < 0) return -2;
    if (!K0s) return -3;
    if (!Cb) return -4;
    if (!Mb) return -5;
    if (!K6d) return -7;
    return 0;
}

extern "C" int axs_build_S(int N, int hb, const axs_cplx* K0s, const axs_cplx* Cb, const axs_cplx* Mb,
                           const axs_cplx* K1c, const axs_cplx* K6d, const double* omega, int nf,
                           axs_cplx* S, void* stream)
{
    int rc = check_ops(N, hb, K0s, Cb, Mb, K6d);
    if (rc) return rc;
    if (!omega) return -8;
    if (nf <= 0) return -9;
    if (!S) return -10;
    const long long n2 = 2 * N;
    for (int f0 = 0; f0 < nf; f0 += AXS_GRID_Y) {
        const int cnt = (nf - f0 < AXS_GRID_Y) ? nf - f0 : AXS_GRID_Y;
        rc = axs_launch_build_S(N, hb, (const cplx*)K0s, (const cplx*)Cb, (const cplx*)Mb, (const cplx*)K1c,
                                (const cplx*)K6d, omega + f0, cnt, (cplx*)S + f0 * n2 * n2, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" long long axs_eig_worksize(int N, int nf)
{
    if (N <= 0 || nf <= 0) return 0;
    if (is_big(N)) return axs_work_layout_big(2 * N, nf, nullptr, nullptr);
    return axs_work_layout(2 * N, nf, nullptr, nullptr);
}

extern "C" int axs_reduce(int N, int hb, const axs_cplx* K0s, const axs_cplx* Cb, const axs_cplx* Mb,
                          const axs_cplx* K1c, const axs_cplx* K6d, const double* omega, int nf,
                          const axs_cplx* S_in, void* work, long long work_bytes, void* stream)
{
    if (N <= 0) return -1;
    if (2 * N > AXS_MAX_BIG_N2) return AXS_ERR_TOO_LARGE;
    if (!S_in) {
        int rc = check_ops(N, hb, K0s, Cb, Mb, K6d);
        if (rc) return rc;
        if (!omega) return -8;
    }
    if (nf <= 0) return -9;
    if (!work) return -11;
    if (work_bytes < axs_eig_worksize(N, nf)) return -12;
    if (is_big(N)) {
        AxsWorkBig wb;
        axs_work_layout_big(2 * N, nf, work, &wb);
        return axs_launch_reduce_big(N, hb, (const cplx*)K0s, (const cplx*)Cb, (const cplx*)Mb, (const cplx*)K1c,
                                     (const cplx*)K6d, omega, nf, (const cplx*)S_in, wb, (cudaStream_t)stream);
    }
    AxsWork wk;
    axs_work_layout(2 * N, nf, work, &wk);
    return axs_launch_reduce(N, hb, (const cplx*)K0s, (const cplx*)Cb, (const cplx*)Mb, (const cplx*)K1c,
                             (const cplx*)K6d, omega, nf, (const cplx*)S_in, wk, (cudaStream_t)stream);
}

extern "C" int axs_hqr(int N, int nf, axs_cplx* k, int* info, void* work, long long work_bytes, void* stream)
{
    if (N <= 0) return -1;
    if (2 * N > AXS_MAX_BIG_N2) return AXS_ERR_TOO_LARGE;
    if (nf <= 0) return -2;
    if (!k) return -3;
    if (!info) return -4;
    if (!work) return -5;
    if (work_bytes < axs_eig_worksize(N, nf)) return -6;
    if (is_big(N)) {
        AxsWorkBig wb;
        axs_work_layout_big(2 * N, nf, work, &wb);
        return axs_launch_hqr_big(2 * N, nf, (cplx*)k, info, wb, (cudaStream_t)stream);
    }
    AxsWork wk;
    axs_work_layout(2 * N, nf, work, &wk);
    return axs_launch_hqr(2 * N, nf, (cplx*)k, info, wk, (cudaStream_t)stream);
}

extern "C" int axs_modes(int N, int hb, const axs_cplx* Cb, const axs_cplx* K6d, int nf,
                         const axs_cplx* k, axs_cplx* psi, void* work, long long work_bytes, void* stream)
{
    if (N <= 0) return -1;
    if (2 * N > AXS_MAX_BIG_N2) return AXS_ERR_TOO_LARGE;
    if (hb < 0) return -2;
    if (!Cb) return -3;
    if (!K6d) return -4;
    if (nf <= 0) return -5;
    if (!k) return -6;
    if (!psi) return -7;
    if (!work) return -8;
    if (work_bytes < axs_eig_worksize(N, nf)) return -9;
    const int n2 = 2 * N;
    const axs_cplx* kk = k;
    for (int f0 = 0; f0 < nf; f0 += AXS_GRID_Y) {        // kernels with nf in gridDim.y: <= 65535 per launch
        const int cnt = (nf - f0 < AXS_GRID_Y) ? nf - f0 : AXS_GRID_Y;
        int rc;
        if (is_big(N)) {
            AxsWorkBig wb;
            axs_work_layout_big(n2, nf, work, &wb);
            rc = axs_launch_modes_big(N, hb, (const cplx*)Cb, (const cplx*)K6d, cnt, (const cplx*)kk + (long long)f0 * n2,
                                      (cplx*)psi + (long long)f0 * n2 * n2, workbig_at(wb, n2, f0), (cudaStream_t)stream);
        } else {
            AxsWork wk;
            axs_work_layout(n2, nf, work, &wk);
            rc = axs_launch_modes(N, hb, (const cplx*)Cb, (const cplx*)K6d, cnt, (const cplx*)kk + (long long)f0 * n2,
                                  (cplx*)psi + (long long)f0 * n2 * n2, work_at(wk, n2, f0), (cudaStream_t)stream);
        }
        if (rc) return rc;
    }
    return 0;
}

extern "C" int axs_eig_sweep(int N, int hb, const axs_cplx* K0s, const axs_cplx* Cb, const axs_cplx* Mb,
                             const axs_cplx* K1c, const axs_cplx* K6d, const double* omega, int nf,
                             axs_cplx* k, axs_cplx* psi, int* info, void* work, long long work_bytes,
                             void* stream)
{
    int rc = axs_reduce(N, hb, K0s, Cb, Mb, K1c, K6d, omega, nf, nullptr, work, work_bytes, stream);
    if (rc) return rc;
    if (!k) return -10;
    if (!info) return -12;
    rc = axs_hqr(N, nf, k, info, work, work_bytes, stream);
    if (rc) return rc;
    if (psi) rc = axs_modes(N, hb, Cb, K6d, nf, k, psi, work, work_bytes, stream);
    return rc;
}

extern "C" int axs_get_reduction(int N, int nf, const void* work, double* scale, axs_cplx* H, void* stream)
{
    if (N <= 0) return -1;
    if (nf <= 0) return -2;
    if (!work) return -3;
    if (is_big(N)) {
        AxsWorkBig wb;
        axs_work_layout_big(2 * N, nf, (void*)work, &wb);
        if (scale)
            AXS_CUDA_OK(cudaMemcpyAsync(scale, wb.scale, (size_t)nf * 2 * N * sizeof(double),
                                        cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        if (H) return axs_launch_unpack_big(2 * N, nf, wb, (cplx*)H, (cudaStream_t)stream);
        return 0;
    }
    AxsWork wk;
    axs_work_layout(2 * N, nf, (void*)work, &wk);
    if (scale)
        AXS_CUDA_OK(cudaMemcpyAsync(scale, wk.scale, (size_t)nf * 2 * N * sizeof(double),
                                    cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (H) return axs_launch_unpack(2 * N, nf, wk, (cplx*)H, (cudaStream_t)stream);
    return 0;
}

extern "C" int axs_hqr_stats(int N, int nf, const void* work, int* stats, void* stream)
{
    if (N <= 0) return -1;
    if (nf <= 0) return -2;
    if (!work) return -3;
    if (!stats) return -4;
    if (is_big(N)) return -1;                              /* counters exist for the shared-memory path only */
    AxsWork wk;
    axs_work_layout(2 * N, nf, (void*)work, &wk);
    AXS_CUDA_OK(cudaMemcpyAsync(stats, wk.stats, (size_t)nf * AXS_QR_STATS * sizeof(int), cudaMemcpyDeviceToDevice,
                                (cudaStream_t)stream));
    return 0;
}

extern "C" int axs_copy_to_pinned(const void* src, void* h_pinned_dst, long long bytes, void* stream)
{
    if (!src) return -1;
    if (!h_pinned_dst) return -2;
    if (bytes < 0 || (bytes & 3)) return -3;
    return axs_launch_copy_to_pinned(src, h_pinned_dst, bytes, (cudaStream_t)stream);
}

extern "C" long long axs_track_worksize(int N, int npairs)
{
    if (N <= 0 || npairs <= 0) return 0;
    return (long long)(npairs + 1) * 4 * N * N * (long long)sizeof(cplx);
}

extern "C" int axs_track_pairs(int N, int hb, const axs_cplx* Cb, const axs_cplx* K6d, const axs_cplx* psi,
                               int npairs, int* arg, double* amax, void* work, long long work_bytes,
                               void* stream)
{
    if (N <= 0) return -1;
    if (2 * N > AXS_MAX_BIG_N2) return AXS_ERR_TOO_LARGE;
    if (hb < 0) return -2;
    if (!Cb) return -3;
    if (!K6d) return -4;
    if (!psi) return -5;
    if (npairs <= 0) return -6;
    if (!arg) return -7;
    if (!amax) return -8;
    if (!work) return -9;
    if (work_bytes < axs_track_worksize(N, npairs)) return -10;
    const long long n2 = 2 * N;
    for (int p0 = 0; p0 < npairs; p0 += AXS_GRID_Y) {
        const int cnt = (npairs - p0 < AXS_GRID_Y) ? npairs - p0 : AXS_GRID_Y;
        const int rc = axs_launch_track(N, hb, (const cplx*)Cb, (const cplx*)K6d, (const cplx*)psi + p0 * n2 * n2, cnt,
                                        (cplx*)work, arg + p0 * n2, amax + p0 * n2, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int axs_track_scan(const int* h_arg, const double* h_amax, int nf, int n2, int start, int* h_order)
{
    if (start <= 0) {
        for (int r = 0; r < n2; ++r) h_order[r] = r;
        start = 1;
    }
    for (int i = start; i < nf; ++i) {
        const int* a = h_arg + (long long)(i - 1) * n2;
        const double* m = h_amax + (long long)(i - 1) * n2;
        const int* prev = h_order + (long long)(i - 1) * n2;
        int* cur = h_order + (long long)i * n2;
        bool all_valid = true;
        for (int r = 0; r < n2; ++r) {
            const int src = prev[r];
            cur[r] = a[src];
            // np.round(x, 2) > 0.9   (numpy: rint(x * 100) / 100; q / 100.0 > 0.9 <=> q >= 91 for integral q,
            // because 90 / 100.0 is the double 0.9 itself) - no division in the loop
            if (!(rint(m[src] * 100.0) >= 91.0)) all_valid = false;
        }
        if (!all_valid) return i;      // caller applies the reference's set-based fix-up to row i
    }
    return nf;
}

extern "C" int axs_apply_order(int n2, int nf, const int* order, const axs_cplx* k, const axs_cplx* psi,
                               axs_cplx* k_out, axs_cplx* psi_out, void* stream)
{
    if (n2 <= 0) return -1;
    if (nf <= 0) return -2;
    if (!order) return -3;
    if (!k) return -4;
    if (!k_out) return -6;
    for (int f0 = 0; f0 < nf; f0 += AXS_GRID_Y) {
        const int cnt = (nf - f0 < AXS_GRID_Y) ? nf - f0 : AXS_GRID_Y;
        const long long o1 = (long long)f0 * n2, o2 = o1 * n2;
        const int rc = axs_launch_apply_order(n2, cnt, order + o1, (const cplx*)k + o1,
                                              psi ? (const cplx*)psi + o2 : nullptr, (cplx*)k_out + o1,
                                              psi_out ? (cplx*)psi_out + o2 : nullptr, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int axs_select_modes(int n2, int nf, const axs_cplx* k, const double* sigma, int m, int* rank,
                                axs_cplx* psi, void* stream)
{
    if (n2 <= 0) return -1;
    if (nf <= 0) return -2;
    if (!k) return -3;
    if (!sigma) return -4;
    if (m <= 0 || m > n2) return -5;
    if (!rank) return -6;
    for (int f0 = 0; f0 < nf; f0 += AXS_GRID_Y) {
        const int cnt = (nf - f0 < AXS_GRID_Y) ? nf - f0 : AXS_GRID_Y;
        const long long o1 = (long long)f0 * n2;
        const int rc = axs_launch_select_modes(n2, cnt, (const cplx*)k + o1, sigma + f0, m, rank + o1,
                                               psi ? (cplx*)psi + o1 * n2 : nullptr, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int axs_energy_ratio(int N, int nf, const axs_cplx* psi, const int* tclass,
                                int nnz_total, const int* rows_total, const int* cols_total, const axs_cplx* vals_total,
                                int nnz_pml, const int* rows_pml, const int* cols_pml, const axs_cplx* vals_pml,
                                double* er, void* stream)
{
    if (N <= 0) return -1;
    if (nf <= 0) return -2;
    if (!psi) return -3;
    if (!tclass) return -4;
    if (nnz_total <= 0 || !rows_total || !cols_total || !vals_total) return -5;
    if (nnz_pml < 0 || (nnz_pml > 0 && (!rows_pml || !cols_pml || !vals_pml))) return -9;
    if (!er) return -13;
    const long long n2 = 2 * N;
    for (int f0 = 0; f0 < nf; f0 += AXS_GRID_Y) {
        const int cnt = (nf - f0 < AXS_GRID_Y) ? nf - f0 : AXS_GRID_Y;
        const int rc = axs_launch_energy_ratio(N, cnt, (const cplx*)psi + f0 * n2 * n2, tclass, nnz_total, rows_total,
                                               cols_total, (const cplx*)vals_total, nnz_pml, rows_pml, cols_pml,
                                               (const cplx*)vals_pml, er + f0 * n2, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int axs_quadratic_forms(int N, int nf, const axs_cplx* psi, const int* tclass, int n_sel, const int* sel,
                                   int n_mat, const int* mat_off, const int* rows, const int* cols,
                                   const axs_cplx* vals, axs_cplx* out, void* stream)
{
    if (N <= 0) return -1;
    if (nf <= 0) return -2;
    if (!psi) return -3;
    if (!tclass) return -4;
    if (n_sel < 0 || n_sel > 2 * N) return -5;
    if (n_sel == 0) return 0;                              /* no mode selected: nothing to do (empty result) */
    if (n_mat <= 0) return -7;
    if (!mat_off) return -8;
    if (!rows) return -9;
    if (!cols) return -10;
    if (!vals) return -11;
    if (!out) return -12;
    const long long n2 = 2 * N;
    for (int f0 = 0; f0 < nf; f0 += AXS_GRID_Y) {
        const int cnt = (nf - f0 < AXS_GRID_Y) ? nf - f0 : AXS_GRID_Y;
        const int rc = axs_launch_quad_forms(N, cnt, (const cplx*)psi + f0 * n2 * n2, tclass, n_sel, sel, n_mat, mat_off,
                                             rows, cols, (const cplx*)vals, (cplx*)out + (long long)f0 * n_sel * n_mat,
                                             (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int axs_modal_projection(int N, int nf, const axs_cplx* psi, const axs_cplx* fvec, axs_cplx* out,
                                    void* stream)
{
    if (N <= 0) return -1;
    if (nf <= 0) return -2;
    if (!psi) return -3;
    if (!fvec) return -4;
    if (!out) return -5;
    const long long n2 = 2 * N;
    for (int f0 = 0; f0 < nf; f0 += AXS_GRID_Y) {
        const int cnt = (nf - f0 < AXS_GRID_Y) ? nf - f0 : AXS_GRID_Y;
        const int rc = axs_launch_modal_projection(2 * N, cnt, (const cplx*)psi + f0 * n2 * n2, (const cplx*)fvec,
                                                   (cplx*)out + f0 * n2, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}
