// The multi-bulge QR kernel of the shared-memory eigen path (one CTA per matrix).  A header of its own so
// that tests/host/hqr_emul.cpp can compile this very source as host code and run it on an emulated CTA
// (threads as fibers, __syncthreads as a rendezvous: tests/host/cta_emul.h) without a GPU.
#pragma once
#include "common.cuh"
#include "aed.cuh"
#ifndef AXS_AED_BUILD
#define AXS_AED_BUILD 0                    /* product build: the shared-memory window routine (eig_small.cu) */
#endif
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
// =========================================================================================
// k_hqr_mb
// =========================================================================================
// k_hqr_mb: implicit single-shift QR with several bulges chased concurrently ("small-bulge
// multishift").  A batch takes nb shifts (eigenvalues of the trailing 2x2 diagonal blocks of
// the active window, closed form), starts bulge b at macro step b*MB_D and advances every
// live bulge by one position per macro step:
//     phase R  rows (k, k+1) <- G_k rows,   one thread per (bulge, column)
//     phase C  cols (k, k+1) <- cols G_k^H, one thread per (bulge, row); the thread that owns
//              row k+1 also updates row k+2 (the new bulge) and generates G_{k+1}
// Left and right multiplications commute and bulges are >= 3 rows apart, so the result is
// exactly that of nb successive single-shift sweeps, at ~1/nb of the barrier-bound latency.
#define MB_MAX 8
#define MB_D 3
#define MB_THREADS 512
#define MB_WORKERS (MB_THREADS - 32)
#define MB_ITEMS 3                      // ceil(MB_MAX * (AXS_MAX_SMALL_N2 + 1) / MB_THREADS)
#define AED_ITEMS 3                     // ceil(window rows above x AED_MAX / threads) for every stage instance

// Staging: the active window only shrinks, so the sweep is cut into stages with a smaller
// shared-memory footprint each; a stage stops once the window is <= m_stop, parks the leading
// block and its size in global memory, and the next stage resumes with more CTAs per SM
// (the kernel is latency bound, so co-resident matrices translate into throughput).
// OPT (stage 1, one CTA per SM, where the leader chain suffers from contention on its
// sub-partition): the warps that share the leader warp's SM sub-partition take no phase-C items.
// Rotations have a real cosine (zlartg form, common.cuh).  Variants that were measured and
// rejected in round 1 (explicit-shift wavefronts, register carry, leader-owned diagonal block,
// complex-cosine rotations, the windowed warp-specialised stage 1) are in the git history
// (commit 3cb38ed) and in DESIGN.md section 4.5; they are no longer compiled into the library.
// AED = false instances do not reference the window scratch (2.6 KB of static shared memory), so their
// stage capacities are the round-1 ones (117 / 95 / 82 instead of 116 / 93 / 79).
template <int NT, int MINB, int NITEMS, bool OPT = false, bool AED = false>
__global__ void __launch_bounds__(NT, MINB)
k_hqr_mb(int n, int bmax, int first_stage, int m_stop, int aed_nw, const cplx* __restrict__ Hc_all,
         cplx* __restrict__ Hp_all, cplx* __restrict__ eig_all, int* __restrict__ info, int* __restrict__ ucur,
         int* __restrict__ stats)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    const int f = blockIdx.x, tid = threadIdx.x;
    const int hsz = hcol_size(n);
    cplx* H = (cplx*)smem_raw;
    __shared__ int s_l;
    __shared__ cplx s_shift[AED_MAX > MB_MAX ? AED_MAX : MB_MAX];
    __shared__ cplx s_ra[2][MB_MAX], s_rb[2][MB_MAX];
    __shared__ cplx s_rr[2][MB_MAX];
#define s_aed g_aed                                       /* file-scope shared scratch of csrc/aed.cuh */

    const cplx* Hg = (first_stage ? Hc_all : Hp_all) + (long long)f * hsz;
    cplx* Hpark = Hp_all + (long long)f * hsz;
    int u = first_stage ? n - 1 : ucur[f];
    if (u < 0) return;                                    // finished in an earlier stage
    for (int i = tid; i < hcol_size(u + 1); i += NT) H[i] = Hg[i];
    cplx* eig = eig_all + (long long)f * n;
    __syncthreads();

#ifdef AED_PROFILE
    if constexpr (AED) if (tid < 8) g_aed.prof[tid] = 0;
#endif
    const bool aed_reg = (aed_nw & 0x100) != 0;           // register-resident window routine (AXS_AED_BUILD == 2 only)
    (void)aed_reg;
    aed_nw &= 0xff;
    int its = 0, failed = 0;
    int n_macro = 0, n_batch = 0, n_aed = 0, n_aed_defl = 0;
    long long clk[4] = {0, 0, 0, 0}, clk0 = 0;            // thread 0: cycles in the AED window / its application / the sweeps
#define AED_CLK(i) if (tid == 0) { const long long c_ = clock64(); if (i) clk[i] += c_ - clk0; clk0 = c_; }
    bool have_shifts = false;                             // s_shift[0..ns_aed) hold the AED window's eigenvalues
    int ns_aed = 0;
    while (u >= m_stop) {
        // ---- deflation scan ----------------------------------------------------------------
        if (tid == 0) s_l = 0;
        __syncthreads();
        for (int k = 1 + tid; k <= u; k += NT) {
            const double sub = cabs1(HC(H, k, k - 1));
            bool neg = false;
            if (sub <= SMALL_D) neg = true;
            else {
                const cplx hkk = HC(H, k, k), hk1k1 = HC(H, k - 1, k - 1);
                double tst = cabs1(hk1k1) + cabs1(hkk);
                if (tst == 0.0) {
                    if (k - 2 >= 0) tst += cabs1(HC(H, k - 1, k - 2));
                    if (k + 1 <= n - 1) tst += cabs1(HC(H, k + 1, k));
                }
                if (sub <= EPS_D * tst) {
                    const double up = cabs1(HC(H, k - 1, k));
                    const double ab = fmax(sub, up), ba = fmin(sub, up);
                    const double dd = cabs1(csub(hk1k1, hkk));
                    const double aa = fmax(cabs1(hkk), dd), bb = fmin(cabs1(hkk), dd);
                    const double sm = aa + ab;
                    if (ba * (ab / sm) <= fmax(SMALL_D, EPS_D * (bb * (aa / sm)))) neg = true;
                }
            }
            if (neg) atomicMax(&s_l, k);
        }
        __syncthreads();
        const int l = s_l;
        if (l > 0 && tid == 0) HC(H, l, l - 1) = mk(0, 0);
        if (l == u) {
            if (tid == 0) eig[u] = HC(H, u, u);
            --u; its = 0; have_shifts = false;
            __syncthreads();
            continue;
        }
        ++its;
        if (its > 120) {
            if (tid == 0) eig[u] = HC(H, u, u);
            ++failed; --u; its = 0; have_shifts = false;
            __syncthreads();
            continue;
        }
        // ---- aggressive early deflation on the trailing window (zlaqr2 semantics) ------------
        have_shifts = false;
        if constexpr (AED) {
        if (aed_nw >= 2 && u - l + 1 >= 3) {
            __syncthreads();                              // H(l, l-1) = 0 is visible
            const int nw = (aed_nw < u - l + 1) ? aed_nw : u - l + 1;
            const int kw = u - nw + 1;
            AED_CLK(0)
            if (tid < 32) {
#if AXS_AED_BUILD == 1
                aed_window_reg(H, l, u, nw, tid);         // register-resident routine only
#elif AXS_AED_BUILD == 3
                aed_window_flat(H, l, u, nw, tid);        // shared-memory routine as one loop around one rotation step
#elif AXS_AED_BUILD == 0
                aed_window(H, l, u, nw, tid);             // shared-memory routine only (product build)
#else
                if (aed_reg) aed_window_reg(H, l, u, nw, tid);
                else aed_window(H, l, u, nw, tid);
#endif
            }
            __syncthreads();
            AED_CLK(1)
            ++n_aed;
            const int aed_ok = s_aed.ok;
            const int ns = s_aed.ns, nd = nw - ns;        // undeflated / deflated window eigenvalues
            if (aed_ok && nd > 0) {
                // rows of the active block above the window: H[l:kw, kw:u+1] <- H[l:kw, kw:u+1] Z
                // (one thread per output entry, results held back until every input has been read)
                const int nra = kw - l;
                cplx upd[AED_ITEMS];
#pragma unroll
                for (int q = 0; q < AED_ITEMS; ++q) {
                    const int idx = tid + q * NT;
                    upd[q] = mk(0, 0);
                    if (idx < nra * nw) {
                        const int j = idx / nra, i = l + (idx - j * nra);
                        cplx acc = mk(0, 0);
                        for (int qq = 0; qq < nw; ++qq) acc = cfma(HC(H, i, kw + qq), s_aed.Z[qq][j], acc);
                        upd[q] = acc;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < AED_ITEMS; ++q) {
                    const int idx = tid + q * NT;
                    if (idx < nra * nw) {
                        const int j = idx / nra, i = l + (idx - j * nra);
                        HC(H, i, kw + j) = upd[q];
                    }
                }
                // the window itself: T (Hessenberg in its leading ns x ns part, triangular below) and the spike
                for (int idx = tid; idx < nw * nw; idx += NT) {
                    const int j = idx / nw, i = idx - j * nw;
                    if (i <= j + 1) HC(H, kw + i, kw + j) = s_aed.T[i][j];
                }
                if (tid == 0 && kw > l) HC(H, kw, kw - 1) = s_aed.spike_new;
                for (int j = ns + tid; j < nw; j += NT) eig[kw + j] = s_aed.T[j][j];
                n_aed_defl += nd;
                u = kw + ns - 1;
                its = 0;
            }
            if (aed_ok) {
                for (int j = tid; j < ns; j += NT) s_shift[j] = s_aed.shift[j];
                have_shifts = ns > 0;
                ns_aed = ns;
            }
            __syncthreads();
            AED_CLK(2)
            if (aed_ok && nd > 0) {
                // LAPACK's "nibble": a productive window is followed by another one, not by a sweep
                if (u < l + 1 || nd * 100 >= 14 * nw) continue;
            }
        }
        }                                                 // if constexpr (AED)
        const int m = u - l + 1;
        if (m < 2) continue;
        int nb = (m - 2) / MB_D + 1;
        if (nb > bmax) nb = bmax;
        // ---- shifts ---------------------------------------------------------------------------
        if (have_shifts && its % 10 != 0) {
            // the undeflated eigenvalues of the AED window, the ones nearest to H(u, u) first
            if (nb > ns_aed) nb = ns_aed;
            __shared__ cplx s_tmp[AED_MAX];
            const cplx huu = HC(H, u, u);
            if (tid < ns_aed) {
                const cplx me = s_shift[tid];
                const double dme = cabs1(csub(me, huu));
                int rank = 0;
                for (int q = 0; q < ns_aed; ++q) {
                    const double dq = cabs1(csub(s_shift[q], huu));
                    rank += (dq < dme || (dq == dme && q < tid)) ? 1 : 0;
                }
                s_tmp[rank] = me;
            }
            __syncthreads();
            if (tid < ns_aed) s_shift[tid] = s_tmp[tid];
        } else if (tid < (nb + 1) / 2) {
            // eigenvalues of the trailing 2x2 diagonal blocks
            const int i = u - 2 * tid;
            const cplx a11 = HC(H, i - 1, i - 1), a12 = HC(H, i - 1, i), a21 = HC(H, i, i - 1), a22 = HC(H, i, i);
            const cplx hd = cscale(csub(a11, a22), 0.5);
            const cplx disc = csqrt_(cadd(cmul(hd, hd), cmul(a12, a21)));
            const cplx mid = cscale(cadd(a11, a22), 0.5);
            cplx e1 = cadd(mid, disc), e2 = csub(mid, disc);
            if (cabs1(csub(e2, a22)) < cabs1(csub(e1, a22))) { const cplx tmp = e1; e1 = e2; e2 = tmp; }
            if (its % 10 == 0) {                         // exceptional shifts against stagnation
                const double ex = 0.75 * cabs1(HC(H, u, u - 1));
                e1 = cadd(e1, mk(ex, 0)); e2 = cadd(e2, mk(0, ex));
            }
            s_shift[2 * tid] = e1;
            if (2 * tid + 1 < MB_MAX) s_shift[2 * tid + 1] = e2;
        }
        __syncthreads();
        if (tid == 0) {                                  // rotation that introduces bulge 0
            cplx a, b, r;
            make_rot_rc(csub(HC(H, l, l), s_shift[0]), HC(H, l + 1, l), a, b, r);
            s_ra[0][0] = a; s_rb[0][0] = b; s_rr[0][0] = r;
        }
        __syncthreads();
        const int nsteps = (m - 1) + (nb - 1) * MB_D;
        const int CW = m + 1;                            // columns l-1 .. u
        // per-thread work items are fixed for the whole batch: (bulge, column) for the row
        // rotations, (bulge, row) for the column rotations; only k = k0 + t moves.
        int r_k0[NITEMS], r_col[NITEMS], r_off[NITEMS], c_k0[NITEMS], c_row[NITEMS];
        // phase-C workers: all worker warps, or (OPT) only those on other sub-partitions than the leader's
        const int wq = tid >> 5;
        const int NCW = OPT ? (NT / 32 - NT / 128) * 32 : (NT - 32);
        const int cidx = OPT ? (((wq & 3) == 3) ? (1 << 30) : (((wq >> 2) * 3 + (wq & 3)) * 32 + (tid & 31)))
                             : ((tid < (NT - 32)) ? tid : (1 << 30));
#pragma unroll
        for (int q = 0; q < NITEMS; ++q) {
            const int item = (tid < (NT - 32)) ? tid + q * (NT - 32) : (1 << 30);   // last warp: leaders only
            r_k0[q] = c_k0[q] = 1 << 28;                 // "never active"
            r_col[q] = 0; r_off[q] = 0; c_row[q] = 0;
            if (item < nb * CW) {
                const int b = item / CW, col = l - 1 + (item - b * CW);
                if (col >= l) { r_k0[q] = l - b * MB_D; r_col[q] = col; r_off[q] = hcol_off(col); }
                // column l-1 never receives a rotation (k-1 >= l whenever the slot is used)
            }
            const int citem = (cidx < (1 << 29)) ? cidx + q * NCW : (1 << 30);
            if (citem < nb * m) {
                const int b = citem / m;
                c_k0[q] = l - b * MB_D; c_row[q] = l + (citem - b * m);
            }
        }
        n_macro += nsteps; ++n_batch;
        AED_CLK(0)
        for (int t = 0; t < nsteps; ++t) {
            const int cur = t & 1, nxt = cur ^ 1;
            // ---- phase R: row rotations ---------------------------------------------------------
#pragma unroll
            for (int q = 0; q < NITEMS; ++q) {
                const int k = r_k0[q] + t, col = r_col[q];
                if (k >= l && k <= u - 1 && col >= k - 1 && k < (1 << 27)) {
                    const int b = (l - r_k0[q]) / MB_D;
                    cplx* pk = H + r_off[q] + k;
                    if (col == k - 1) { *pk = s_rr[cur][b]; }
                    else {
                        const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                        const cplx p = pk[0], qv = pk[1];
                        cplx top, bot;
                        rot_rows<true>(ga, gb, p, qv, top, bot);
                        pk[0] = top; pk[1] = bot;
                    }
                }
            }
            __syncthreads();
            // ---- phase C: column rotations, next rotation ---------------------------------------
            // The last warp owns the critical chain of every bulge (rows k+1, k+2 -> next rotation)
            // so that it overlaps with the bulk updates done by the other warps.
            if (tid >= (NT - 32)) {
                const int b = tid - ((NT - 32));
                const int k = l + t - b * MB_D;
                if (b < nb && k >= l && k <= u - 1) {
                    const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                    cplx* p0 = H + hcol_off(k) + k + 1;
                    cplx* p1 = H + hcol_off(k + 1) + k + 1;
                    const cplx c0 = *p0, c1 = *p1;
                    const cplx h = (k + 2 <= u) ? p1[1] : mk(0, 0);         // H(k+2, k+1), before the stores
                    cplx n0, n1;
                    rot_cols<true>(ga, gb, c0, c1, n0, n1);
                    *p0 = n0;
                    *p1 = n1;
                    if (k + 2 <= u) {
                        const cplx y = cmulc(h, gb);                        // new bulge H(k+2, k) = h conj(b)
                        p1[1] = cscale(h, ga.x);
                        cplx a2, b2, r2;
                        make_rot_rc(n0, y, a2, b2, r2);
                        s_ra[nxt][b] = a2; s_rb[nxt][b] = b2; s_rr[nxt][b] = r2;
                    }
                }
                if (b == nb && (t + 1) % MB_D == 0 && (t + 1) / MB_D < nb) {   // bulge (t+1)/D starts next step
                    const int bn = (t + 1) / MB_D;
                    cplx a2, b2, r2;
                    make_rot_rc(csub(HC(H, l, l), s_shift[bn]), HC(H, l + 1, l), a2, b2, r2);
                    s_ra[nxt][bn] = a2; s_rb[nxt][bn] = b2; s_rr[nxt][bn] = r2;
                }
            }
#pragma unroll
            for (int q = 0; q < NITEMS; ++q) {
                const int k = c_k0[q] + t, r = c_row[q];
                if (k >= l && k <= u - 1 && r <= k) {
                    const int b = (l - c_k0[q]) / MB_D;
                    const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                    cplx* p0 = H + hcol_off(k) + r;
                    cplx* p1 = H + hcol_off(k + 1) + r;
                    const cplx c0 = *p0, c1 = *p1;
                    cplx n0, n1;
                    rot_cols<true>(ga, gb, c0, c1, n0, n1);
                    *p0 = n0; *p1 = n1;
                }
            }
            __syncthreads();
        }
        AED_CLK(3)
    }
#ifdef AED_PROFILE
    if constexpr (AED) if (tid == 0 && first_stage && f == 0 && aed_nw >= 2)
        printf("AED_PROFILE stage1 f0: calls %lld | cycles scan %lld shift %lld rot %lld apply %lld | steps %lld sweeps %lld rehess %lld\n",
               g_aed.prof[7], g_aed.prof[0], g_aed.prof[1], g_aed.prof[2], g_aed.prof[3], g_aed.prof[4], g_aed.prof[5], g_aed.prof[6]);
#endif
    if (tid == 0) {
        info[f] = first_stage ? failed : info[f] + failed;
        ucur[f] = u;
        if (stats) {                                      // per-matrix counters (axs_hqr_stats): macro steps,
            int* sp = stats + AXS_QR_STATS * (long long)f;   // batches, AED calls, eigenvalues deflated by AED,
            const int stg = first_stage ? 0 : (NT == 512 ? 1 : NT == 384 ? 2 : 3);   // kilo-cycles: window / apply / sweeps per stage
            if (first_stage) for (int q = 0; q < AXS_QR_STATS; ++q) sp[q] = 0;
            sp[0] += n_macro; sp[1] += n_batch; sp[2] += n_aed; sp[3] += n_aed_defl;
            sp[4 + 3 * stg] += (int)(clk[1] >> 10); sp[5 + 3 * stg] += (int)(clk[2] >> 10); sp[6 + 3 * stg] += (int)(clk[3] >> 10);
        }
    }
    if (u >= 0)                                           // park the remaining leading block
        for (int i = tid; i < hcol_size(u + 1); i += NT) Hpark[i] = H[i];
}

static size_t hqr_mb_smem(int n) { return (size_t)hcol_size(n) * sizeof(cplx); }
