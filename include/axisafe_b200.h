/*
 * axisafe_b200.h - C ABI of the B200-native SAFE dispersion sweep.
 *
 * The reference (michalkalkowski/axisym-safe-python) is pure Python and has no
 * FFI; its boundary for this path is the Python call surface
 *     Mesh.assemble_matrices(n)            axisafe/mesh.py:290
 *     WaveElementAxisym.solve(f)           axisafe/solver.py:76
 * The entry points below are what a ctypes binding inside those two methods
 * calls (see INTEGRATION.md for the stub).  Each one cites the reference code
 * it replaces.
 *
 * Conventions
 *   - every pointer is a CUDA DEVICE pointer unless the name starts with h_;
 *   - complex numbers are interleaved (re, im) IEEE doubles ("cplx");
 *   - the caller allocates every buffer, nothing is retained between calls;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *     calls are asynchronous on that stream;
 *   - return 0 on success, -k if argument k (1-based) is invalid, >0 = CUDA
 *     error code (cudaError_t) - decode either with axs_strerror().
 *   - per-frequency convergence failures do not fail the call: info[f] holds
 *     the number of eigenvalues that did not converge (0 = ok).
 */
#ifndef AXISAFE_B200_H
#define AXISAFE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } axs_cplx;

int         axs_version(void);
const char* axs_strerror(int code);

/* Largest dense problem size (2N) handled by the shared-memory eigen kernels (one SM holds the
 * whole Hessenberg matrix); above it and up to axs_max_n2() the same entry points run the
 * global-memory instance of the kernels: a blocked compact-WY Hessenberg reduction with DMMA trailing
 * updates that moves the batch through the reduction in lockstep (csrc/eig_bigred.cu), a windowed
 * multi-bulge QR whose off-diagonal strips are shared by a thread-block cluster per matrix when the
 * batch is small, column-oriented inverse iteration and a blocked DMMA back-transformation
 * (csrc/eig_big.cu). */
int axs_max_small_n2(void);
int axs_max_n2(void);

/* ---- kernel 1a: Gauss-Lobatto quadrature of the element matrices ------------------
 * Replaces the per-element Python loops Element.calculate_matrices()
 *   elements.py:148-229 (SLAX6) :293-330 (ALAX6) :447-586 (SLAX6_core)
 *   :683-737 (ALAX6_core) :810-894 (SLAX6_PML) :962-1007 (ALAX6_PML)
 * driven from mesh.py:488-489.  One warp per element.
 *   edesc_i[e][8] : kind(0 solid,1 fluid), family(0 GLL,1 GLJ), pml(0 none,1 exp,2 poly),
 *                   m (nodes), table_off, nodes_off, out_off, reserved
 *   edesc_d[e][12]: lambda, mu, eta, rho, c_p^2, pml_start, pml_thk, pml_a, pml_b, 0,0,0
 *   tables        : per element at table_off: ksi[m], w[m], D[m*m] (D[a*m+q] = l_a'(ksi_q))
 *   nodes         : radii of the element's nodes at nodes_off
 *   out           : per element at out_off (in cplx): 10 matrices nd x nd row-major,
 *                   nd = 3m (solid) or m (fluid), order K1 K2 K3 K4 K5 K6 KF1 KF2 KF3 M
 *                   (full size, axis conditions NOT applied; M of non-PML solids carries the
 *                   reference's per-point complex64 rounding, elements.py:99,:217,:397,:529)
 */
int axs_element_matrices(int n_elem, const int* edesc_i, const double* edesc_d,
                         const double* tables, const double* nodes,
                         axs_cplx* out, void* stream);

/* ---- kernel 1b: scatter-assembly + T-transform into banded operators --------------
 * Replaces mesh.py:518-619 (scatter), solver.py:59-74 (T-transform) and the static
 * parts of solver.py:115-133:
 *     K0s = K1 + n^2 K5 + n K2hat      Cb = n K4 + K3hat     Mb = M    K6d = diag(K6)
 * with Khat[r,c] = i T_r conj(T_c) K[r,c].  Banded row-major, width W = 2*hb+1,
 * entry (r,c) at [r*W + c-r+hb].  Output buffers must be zeroed by the caller.
 *   lm[e] at lm_off[e] : local->global DOF map over the FULL element DOFs, -1 = dropped
 *   tclass[N]          : 0 if Tdiag = 1, 1 if Tdiag = i
 */
int axs_assemble_operators(int n_elem, const int* edesc_i, const axs_cplx* elem,
                           const int* lm, const int* lm_off, const int* tclass,
                           int n_circ, int N, int hb,
                           axs_cplx* K0s, axs_cplx* Cb, axs_cplx* Mb, axs_cplx* K6d,
                           void* stream);

/* ---- kernel 1c: frequency-batched dense companion matrix ---------------------------
 * Replaces solver.py:115-134 (B, Binv = spla.inv(B), A(w), S = Binv A):
 *     S(w) = [[-K6^-1 Cb, -K6^-1 (K0s - w^2 Mb + i w K1c)], [I, 0]]
 * S is [nf][n2][n2] column-major, n2 = 2N.  K1c is banded like K0s (may be NULL).
 */
int axs_build_S(int N, int hb, const axs_cplx* K0s, const axs_cplx* Cb, const axs_cplx* Mb,
                const axs_cplx* K1c, const axs_cplx* K6d, const double* omega, int nf,
                axs_cplx* S, void* stream);

/* ---- kernels 2+3: batched dense eigen-solve of S(w) --------------------------------
 * Replaces val, vec = scipy.linalg.eig(S) (LAPACK zgeev, solver.py:136) and the
 * B-normalisation solver.py:142:
 *   stage A  build S(w), zgebal-style scaling, Householder Hessenberg   (axs_reduce)
 *   stage B  multi-bulge implicit single-shift QR with aggressive early
 *            deflation on the Hessenberg matrix                         (axs_hqr)
 *   stage C  inverse iteration, back-transformation, psi = v/sqrt(v^T B v) (axs_modes)
 * axs_eig_sweep runs A, B, C for nf frequencies.  Workspace: axs_eig_worksize() bytes.
 *   k    [nf][n2]          eigenvalues in the solver's native order
 *   psi  [nf][n2][n2]      row-major, psi[f][r][j] = component r of mode j, rows [0,N) = k*u,
 *                          rows [N,2N) = u (same layout as psi_hat, solver.py:106-109);
 *                          may be NULL (eigenvalues only)
 *   info [nf]
 */
long long axs_eig_worksize(int N, int nf);
int axs_eig_sweep(int N, int hb, const axs_cplx* K0s, const axs_cplx* Cb, const axs_cplx* Mb,
                  const axs_cplx* K1c, const axs_cplx* K6d, const double* omega, int nf,
                  axs_cplx* k, axs_cplx* psi, int* info, void* work, long long work_bytes,
                  void* stream);

/* The three stages separately (used by the parity tests; same workspace layout).
 * axs_reduce accepts an explicit dense input S_in ([nf][n2][n2] column-major) instead of
 * the banded operators when S_in != NULL.  */
int axs_reduce(int N, int hb, const axs_cplx* K0s, const axs_cplx* Cb, const axs_cplx* Mb,
               const axs_cplx* K1c, const axs_cplx* K6d, const double* omega, int nf,
               const axs_cplx* S_in, void* work, long long work_bytes, void* stream);
int axs_hqr(int N, int nf, axs_cplx* k, int* info, void* work, long long work_bytes, void* stream);
int axs_modes(int N, int hb, const axs_cplx* Cb, const axs_cplx* K6d, int nf,
              const axs_cplx* k, axs_cplx* psi, void* work, long long work_bytes, void* stream);
/* Debug/parity accessors into the workspace: copy scale factors D [nf][n2] (double) and the
 * Hessenberg matrix H [nf][n2][n2] (dense column-major, zeros below the subdiagonal). */
int axs_get_reduction(int N, int nf, const void* work, double* scale, axs_cplx* H, void* stream);
/* Counters of the QR stage left in the workspace by axs_hqr (shared-memory path, 2N <= axs_max_small_n2()):
 * stats [nf][16] (device, int32) = macro steps, shift batches, aggressive-early-deflation calls,
 * eigenvalues deflated inside those calls - the iteration counts LAPACK's zhseqr (zlaqr0/zlaqr3
 * behind solver.py:136) would report through its own counters - followed, for each of the four
 * cascade stages, by the kilo-cycles one thread spent in the AED window, applying it, and sweeping. */
int axs_hqr_stats(int N, int nf, const void* work, int* stats, void* stream);

/* Kernel-selection switches (A/B measurements and the variant parity tests).  Every switch is read
 * ONCE from the environment variable AXS_<NAME> when the library is first used; afterwards it can be
 * changed with axs_set_tuning.  Names: reduce_oc, reduce_prebal, hqr_stages, hqr_bulges,
 * hqr_small_direct, hqr_aed, hqr_aed_reg, hqr_aed_stages, backnorm_wy, backnorm_wy_modes, bpsi_tiled, big_blocked,
 * big_panel, big_qr_cluster.  Return 0, or -1 for an unknown name.  Not thread safe against concurrent launches. */
int axs_set_tuning(const char* name, int value);
int axs_get_tuning(const char* name, int* value);

/* ---- mode tracking ------------------------------------------------------------------
 * Replaces the two dense products of solver.py:157 and the row arg-max of :159-161:
 * for pair p (p = 0..npairs-1) with psi_prev = psi[p], psi_cur = psi[p+1],
 *     P = psi_prev^T (B psi_cur),  B = [[0, K6], [K6, Cb]]
 *     arg[p][a] = argmax_b |P[a,b]|,   amax[p][a] = max_b |P[a,b]|
 * psi holds npairs+1 consecutive frequencies in the solver's native (unsorted) order; the
 * workspace (axs_track_worksize bytes) receives Y = B psi.  The sequential permutation scan (solver.py:159-182) is
 * axs_track_scan (host).
 */
long long axs_track_worksize(int N, int npairs);
int axs_track_pairs(int N, int hb, const axs_cplx* Cb, const axs_cplx* K6d,
                    const axs_cplx* psi, int npairs, int* arg, double* amax,
                    void* work, long long work_bytes, void* stream);

/* HOST function.  Composes the per-pair arg-max tables into the reference's column order:
 * order[0] = identity (or h_order[start-1] when start > 0), order[i][r] =
 * arg[i-1][order[i-1][r]]; rows with round(amax,2) <= 0.9 need the reference's Python-set
 * fix-up: the function stops there and returns that frequency index (caller fills
 * order[i] and resumes with start = i+1).  Returns nf when done.  */
int axs_track_scan(const int* h_arg, const double* h_amax, int nf, int n2, int start,
                   int* h_order);

/* Small device -> PINNED host copy executed by the SMs (a kernel storing through the unified address space)
 * instead of the copy engine, so that the few hundred KB of tracking tables the host scan waits for
 * (solver.py:159-182 needs them in order) do not queue behind the bulk download of psi_hat on the same
 * D2H engine.  h_pinned_dst must be page-locked host memory (cudaHostAlloc / torch pin_memory); bytes a
 * multiple of 4.  Asynchronous on `stream`. */
int axs_copy_to_pinned(const void* src, void* h_pinned_dst, long long bytes, void* stream);

/* Column gather after tracking: k_out[f][c] = k[f][order[f][c]],
 * psi_out[f][r][c] = psi[f][r][order[f][c]]  (solver.py:176-182). psi/psi_out may be NULL. */
int axs_apply_order(int n2, int nf, const int* order, const axs_cplx* k, const axs_cplx* psi,
                    axs_cplx* k_out, axs_cplx* psi_out, void* stream);

/* ---- subset of the spectrum (the no_of_waves != 'full' mode) -------
 * Replaces scipy.sparse.linalg.eigs(S, k=m, sigma=w/central_wavespeed, which='LM') of
 * solver.py:137-140 - shift-invert Arnoldi returns the m eigenvalues nearest to sigma; with the
 * whole spectrum already on the device that is a selection:
 *     rank[f][j] = position of native mode j of frequency f when sorted by |k[f][j] - sigma[f]|
 *                  (ties by native index),
 * and, when psi != NULL, the mode-shape columns of rank >= m are zeroed in place so that
 * axs_track_pairs (arg-max over columns) only sees the selected modes.  psi is [nf][n2][n2]. */
int axs_select_modes(int n2, int nf, const axs_cplx* k, const double* sigma, int m, int* rank,
                     axs_cplx* psi, void* stream);

/* ---- energy ratio ------------------------------------------
 * Replaces the dense batched products of WaveElementAxisym.energy_ratio, solver.py:248-266:
 *     q = conj(T) psi[N:],  E_tot = q^H M_total q,  E_pml = q^H M_pml q,  er = |E_pml| / |E_tot|
 * (the common factor w^2/4 cancels).  Both mass matrices are passed as COO triplets over the
 * global DOFs with the reference's fluid "multiplier" (-1 on fluid columns, solver.py:234-239)
 * already applied.  psi is [nf][2N][2N] in the tracked column order; er is [nf][2N]. */
int axs_energy_ratio(int N, int nf, const axs_cplx* psi, const int* tclass,
                     int nnz_total, const int* rows_total, const int* cols_total, const axs_cplx* vals_total,
                     int nnz_pml, const int* rows_pml, const int* cols_pml, const axs_cplx* vals_pml,
                     double* er, void* stream);

/* ---- energy velocity -----------------------------------------
 * Replaces the dense batched products of WaveElementAxisym.energy_velocity, solver.py:374-408:
 * the quadratic forms  out[f][s][m] = q^H X_m q,  q = conj(T) psi[N:, sel[s]]  for n_mat sparse
 * matrices X_m (the core-subset M, K1, K2, K5, KF, KF^T, KFt^T, KFt, K6 with the fluid
 * multiplier applied) given as concatenated COO triplets, matrix m = entries
 * [mat_off[m], mat_off[m+1]).  sel (may be NULL = all 2N modes) lists the mode columns
 * (propagating_indices).  The host combines the forms with k, n and w into e_k, e_p, p. */
int axs_quadratic_forms(int N, int nf, const axs_cplx* psi, const int* tclass, int n_sel, const int* sel,
                        int n_mat, const int* mat_off, const int* rows, const int* cols,
                        const axs_cplx* vals, axs_cplx* out, void* stream);

/* ---- forced response -------------------------------------------
 * Replaces np.matmul(psi_hat^T, force_vector) of WaveElementAxisym.point_excited_waves,
 * solver.py:428-435:  out[f][j] = sum_r psi[f][r][j] * fvec[r],  fvec = [0; f_ext] of length 2N.
 * The scalar factor i sinc(n theta_0 / pi) / (2 pi r_f) is applied by the host. */
int axs_modal_projection(int N, int nf, const axs_cplx* psi, const axs_cplx* fvec, axs_cplx* out,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif
