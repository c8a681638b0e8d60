// Stage 1 of the QR cascade as a warp-specialised, windowed multi-bulge chase on chip
// (eigenvalues of the Hessenberg matrix: zhseqr behind scipy.linalg.eig, reference solver.py:136).
//
// Same mathematics as k_hqr_mb (implicit single-shift QR, up to 8 bulges in flight, real-cosine
// rotations, standard + Ahues-Tisseur deflation), different schedule.  k_hqr_mb applies every
// rotation to the whole active block between two block barriers per macro step, so the LSU time of
// the bulk, its FP64 time and the latency of the rotation chain add up.  Here the
// chase is cut into segments of HW_W macro steps:
//   D team (leader warp on its own SM sub-partition + 6 worker warps) runs the macro steps on the
//          DIAGONAL WINDOW only (rows / columns kmin .. kmax+2 of the segment, ~40 x 40), with team
//          barriers (224 threads), and records the rotations;
//   O team (6 warps) applies the recorded rotations to everything outside the window - row rotations
//          to the columns right of it, column rotations to the rows above it - in bulge-major order
//          with the running element in a register (1 LDS + 1 STS per rotation, no barrier inside).
//          The next HW_W columns (the part of the next window) go first, in (column, bulge)
//          lockstep; the rest overlaps with the D team's next segment.
// Rotations of different bulges commute unless they share a row / column, and then the bulge ahead
// always acts first in the sequential algorithm, so bulge-major order is legal (dev/proto_window_qr.py).
#include "common.cuh"
#include <stdlib.h>
#include <stdio.h>

#define HW_NT 512
#define HW_W 16
#define HW_MB 8
#define HW_D 3
#define HW_DTEAM 224                   // leader warp + 6 worker warps
#define HW_OTEAM 192
#define HW_SLOTS 2                     // window items per D worker

__device__ __forceinline__ void hw_bar(int id, int count) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

__global__ void __launch_bounds__(HW_NT, 1)
k_hqr_win(int n, int bmax, int m_stop, const cplx* __restrict__ Hc_all, cplx* __restrict__ Hp_all,
          cplx* __restrict__ eig_all, int* __restrict__ info, int* __restrict__ ucur)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int f = blockIdx.x, tid = threadIdx.x, wq = tid >> 5, lane = tid & 31;
    const int hsz = hcol_size(n);
    cplx* H = (cplx*)smem_raw;
    __shared__ int s_l;
    __shared__ cplx s_shift[HW_MB];
    __shared__ cplx s_ra[2][HW_MB], s_rb[2][HW_MB], s_rr[2][HW_MB];
    __shared__ cplx s_rot[2][HW_W][HW_MB][2];

    // roles: warp 3 = leader (sub-partition 3 otherwise idle), warps 0-2,4-6 = D workers,
    // warps 8-10,12-14 = O team, warps 7, 11, 15 only take part in the block-wide steps
    const bool leader = (wq == 3);
    const bool dwork = (wq < 7) && ((wq & 3) != 3);
    const bool oteam = (wq >= 8) && ((wq & 3) != 3);
    const int didx = ((wq >> 2) * 3 + (wq & 3)) * 32 + lane;          // 0..191 for D workers
    const int oidx = (((wq - 8) >> 2) * 3 + (wq & 3)) * 32 + lane;    // 0..191 for the O team

    const cplx* Hg = Hc_all + (long long)f * hsz;
    cplx* Hpark = Hp_all + (long long)f * hsz;
    int u = n - 1;
    for (int i = tid; i < hsz; i += HW_NT) H[i] = Hg[i];
    cplx* eig = eig_all + (long long)f * n;
    __syncthreads();

    int its = 0, failed = 0;
#ifdef OC_TIMING
    long long tw[6] = {0, 0, 0, 0, 0, 0}, tw0 = clock64(); int nseg_tot = 0, nstep_tot = 0;
#define HW_MARK(i) { const long long t1c = clock64(); tw[i] += t1c - tw0; tw0 = t1c; }
#else
#define HW_MARK(i)
#endif
    while (u >= m_stop) {
        HW_MARK(5)
        // ---- deflation scan (as k_hqr_mb) ------------------------------------------------------
        if (tid == 0) s_l = 0;
        __syncthreads();
        for (int k = 1 + tid; k <= u; k += HW_NT) {
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
            --u; its = 0;
            __syncthreads();
            continue;
        }
        ++its;
        if (its > 120) {
            if (tid == 0) eig[u] = HC(H, u, u);
            ++failed; --u; its = 0;
            __syncthreads();
            continue;
        }
        const int m = u - l + 1;
        int nb = (m - 2) / HW_D + 1;
        if (nb > bmax) nb = bmax;
        if (tid < (nb + 1) / 2) {
            const int i = u - 2 * tid;
            const cplx a11 = HC(H, i - 1, i - 1), a12 = HC(H, i - 1, i), a21 = HC(H, i, i - 1), a22 = HC(H, i, i);
            const cplx hd = cscale(csub(a11, a22), 0.5);
            const cplx disc = csqrt_(cadd(cmul(hd, hd), cmul(a12, a21)));
            const cplx mid = cscale(cadd(a11, a22), 0.5);
            cplx e1 = cadd(mid, disc), e2 = csub(mid, disc);
            if (cabs1(csub(e2, a22)) < cabs1(csub(e1, a22))) { const cplx tmp = e1; e1 = e2; e2 = tmp; }
            if (its % 10 == 0) {
                const double ex = 0.75 * cabs1(HC(H, u, u - 1));
                e1 = cadd(e1, mk(ex, 0)); e2 = cadd(e2, mk(0, ex));
            }
            s_shift[2 * tid] = e1;
            if (2 * tid + 1 < HW_MB) s_shift[2 * tid + 1] = e2;
        }
        __syncthreads();
        if (tid == 0) {
            cplx a, b, r;
            make_rot_rc(csub(HC(H, l, l), s_shift[0]), HC(H, l + 1, l), a, b, r);
            s_ra[0][0] = a; s_rb[0][0] = b; s_rr[0][0] = r;
        }
        __syncthreads();
        const int nsteps = (m - 1) + (nb - 1) * HW_D;
        const int nseg = (nsteps + HW_W - 1) / HW_W;

        if (leader || dwork) {
            // =================================== D team ===========================================
            for (int s = 0; s < nseg; ++s) {
                HW_MARK(5)
                if (s > 0) hw_bar(2, HW_DTEAM + HW_OTEAM);            // near strip of segment s-1 is done
                HW_MARK(0)
                const int t0 = s * HW_W, t1 = min(nsteps, t0 + HW_W);
                const int kmin = max(l, l + t0 - HW_D * (nb - 1)), kmax = min(u - 1, l + t1 - 1);
                const int wlo = kmin, whi = min(u, kmax + 2);
                const int WWc = whi - wlo + 2, WWr = whi - wlo + 1;   // columns wlo-1..whi, rows wlo..whi
                int r_b[HW_SLOTS], r_col[HW_SLOTS], r_off[HW_SLOTS], c_b[HW_SLOTS], c_row[HW_SLOTS];
#pragma unroll
                for (int q = 0; q < HW_SLOTS; ++q) {
                    const int item = dwork ? didx + q * 192 : (1 << 30);
                    r_b[q] = c_b[q] = 1 << 20;                        // "never active"
                    r_col[q] = 0; r_off[q] = 0; c_row[q] = 0;
                    if (item < nb * WWc) {
                        const int b = item / WWc, col = wlo - 1 + (item - b * WWc);
                        if (col >= l) { r_b[q] = b; r_col[q] = col; r_off[q] = hcol_off(col); }
                    }
                    if (item < nb * WWr) {
                        const int b = item / WWr;
                        c_b[q] = b; c_row[q] = wlo + (item - b * WWr);
                    }
                }
                for (int t = t0; t < t1; ++t) {
                    const int cur = t & 1, nxt = cur ^ 1;
                    // ---- phase R inside the window ---------------------------------------------
#pragma unroll
                    for (int q = 0; q < HW_SLOTS; ++q) {
                        const int b = r_b[q], k = l + t - HW_D * b, col = r_col[q];
                        if (b < nb && k >= l && k <= u - 1 && col >= k - 1) {
                            cplx* pk = H + r_off[q] + k;
                            if (col == k - 1) *pk = s_rr[cur][b];
                            else {
                                const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                                const cplx p = pk[0], qv = pk[1];
                                cplx top, bot;
                                rot_rows<true>(ga, gb, p, qv, top, bot);
                                pk[0] = top; pk[1] = bot;
                            }
                        }
                    }
                    hw_bar(1, HW_DTEAM);
                    // ---- phase C inside the window, rotation chain ----------------------------------
                    if (leader) {
                        const int b = lane;
                        const int k = l + t - b * HW_D;
                        if (b < nb && k >= l && k <= u - 1) {
                            const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                            s_rot[s & 1][t - t0][b][0] = ga; s_rot[s & 1][t - t0][b][1] = gb;
                            cplx* p0 = H + hcol_off(k) + k + 1;
                            cplx* p1 = H + hcol_off(k + 1) + k + 1;
                            const cplx c0 = *p0, c1 = *p1;
                            const cplx h = (k + 2 <= u) ? p1[1] : mk(0, 0);
                            cplx n0, n1;
                            rot_cols<true>(ga, gb, c0, c1, n0, n1);
                            *p0 = n0; *p1 = n1;
                            if (k + 2 <= u) {
                                const cplx y = cmulc(h, gb);
                                p1[1] = cscale(h, ga.x);
                                cplx a2, b2, r2;
                                make_rot_rc(n0, y, a2, b2, r2);
                                s_ra[nxt][b] = a2; s_rb[nxt][b] = b2; s_rr[nxt][b] = r2;
                            }
                        }
                        if (b == nb && (t + 1) % HW_D == 0 && (t + 1) / HW_D < nb) {
                            const int bn = (t + 1) / HW_D;
                            cplx a2, b2, r2;
                            make_rot_rc(csub(HC(H, l, l), s_shift[bn]), HC(H, l + 1, l), a2, b2, r2);
                            s_ra[nxt][bn] = a2; s_rb[nxt][bn] = b2; s_rr[nxt][bn] = r2;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < HW_SLOTS; ++q) {
                        const int b = c_b[q], k = l + t - HW_D * b, r = c_row[q];
                        if (b < nb && k >= l && k <= u - 1 && r <= k) {
                            const cplx ga = s_ra[cur][b], gb = s_rb[cur][b];
                            cplx* p0 = H + hcol_off(k) + r;
                            cplx* p1 = H + hcol_off(k + 1) + r;
                            const cplx c0 = *p0, c1 = *p1;
                            cplx n0, n1;
                            rot_cols<true>(ga, gb, c0, c1, n0, n1);
                            *p0 = n0; *p1 = n1;
                        }
                    }
                    hw_bar(1, HW_DTEAM);
                }
                HW_MARK(1)
#ifdef OC_TIMING
                ++nseg_tot; nstep_tot += t1 - t0;
#endif
                hw_bar(2, HW_DTEAM + HW_OTEAM);                       // segment s recorded
                HW_MARK(2)
            }
        } else if (oteam) {
            // =================================== O team ===========================================
            for (int s = 0; s < nseg; ++s) {
                HW_MARK(5)
                hw_bar(2, HW_DTEAM + HW_OTEAM);                       // wait for segment s
                HW_MARK(0)
                const int t0 = s * HW_W, t1 = min(nsteps, t0 + HW_W);
                const int kmin = max(l, l + t0 - HW_D * (nb - 1)), kmax = min(u - 1, l + t1 - 1);
                const int wlo = kmin, whi = min(u, kmax + 2);
                const int cN1 = min(u, whi + HW_W);                   // near strip: columns whi+1 .. cN1
                // ---- near strip: (column, bulge) lockstep, 4 columns x 8 bulges per warp -------------
                if (oidx < 128) {
                    const int b = lane & 7, c = whi + 1 + (oidx >> 5) * 4 + (lane >> 3);
                    const bool col_ok = (c <= cN1) && (b < nb);
                    cplx carry = mk(0, 0);
                    bool has = false;
                    for (int t = t0; t < t1; ++t) {
                        const int k = l + t - HW_D * b;
                        if (col_ok && k >= l && k <= u - 1) {
                            const cplx ga = s_rot[s & 1][t - t0][b][0], gb = s_rot[s & 1][t - t0][b][1];
                            cplx* pk = H + hcol_off(c) + k;
                            const cplx p = has ? carry : pk[0];
                            const cplx qv = pk[1];
                            cplx top, bot;
                            rot_rows<true>(ga, gb, p, qv, top, bot);
                            pk[0] = top;
                            const bool more = (t + 1 < t1) && (k + 1 <= u - 1);
                            if (more) { carry = bot; has = true; } else { pk[1] = bot; has = false; }
                        }
                        __syncwarp();
                    }
                }
                HW_MARK(1)
                if (s + 1 < nseg) hw_bar(2, HW_DTEAM + HW_OTEAM);     // next segment may start
                HW_MARK(2)
                // ---- far strips: one column (row rotations) or one row (column rotations) per thread ----
                const int cF0 = whi + HW_W + 1;
                const int nfc = max(0, u - cF0 + 1), nfr = wlo - l;
                for (int job = oidx; job < nfc + nfr; job += HW_OTEAM) {
                    if (job < nfc) {
                        const int c = cF0 + job;
                        cplx* colp = H + hcol_off(c);
                        for (int b = 0; b < nb; ++b) {
                            cplx carry = mk(0, 0);
                            int kc = -1;                              // row of the carried element, -1 = none
                            for (int t = t0; t < t1; ++t) {
                                const int k = l + t - HW_D * b;
                                if (k >= l && k <= u - 1) {
                                    const cplx ga = s_rot[s & 1][t - t0][b][0], gb = s_rot[s & 1][t - t0][b][1];
                                    const cplx p = (kc == k) ? carry : colp[k];
                                    const cplx qv = colp[k + 1];
                                    cplx top, bot;
                                    rot_rows<true>(ga, gb, p, qv, top, bot);
                                    colp[k] = top;
                                    carry = bot; kc = k + 1;
                                }
                            }
                            if (kc >= 0) colp[kc] = carry;
                        }
                    } else {
                        const int r = l + (job - nfc);
                        for (int b = 0; b < nb; ++b) {
                            cplx carry = mk(0, 0);
                            int kc = -1;                              // column of the carried element
                            for (int t = t0; t < t1; ++t) {
                                const int k = l + t - HW_D * b;
                                if (k >= l && k <= u - 1) {
                                    const cplx ga = s_rot[s & 1][t - t0][b][0], gb = s_rot[s & 1][t - t0][b][1];
                                    const cplx c0 = (kc == k) ? carry : H[hcol_off(k) + r];
                                    const cplx c1 = H[hcol_off(k + 1) + r];
                                    cplx n0, n1;
                                    rot_cols<true>(ga, gb, c0, c1, n0, n1);
                                    H[hcol_off(k) + r] = n0;
                                    carry = n1; kc = k + 1;
                                }
                            }
                            if (kc >= 0) H[hcol_off(kc) + r] = carry;
                        }
                    }
                }
                HW_MARK(3)
            }
        }
        HW_MARK(5)
        __syncthreads();
        HW_MARK(4)
    }
#ifdef OC_TIMING
    if (f == 0 && (tid == 96 || tid == 0 || tid == 256 || tid == 256 + 191 - 64))
        printf("hqr_win tid %d: [D: waitA %lld steps %lld waitB %lld | O: waitB %lld near %lld waitA %lld far %lld] batchsync %lld other %lld, %d segs %d steps\n",
               tid, tw[0], tw[1], tw[2], tw[0], tw[1], tw[2], tw[3], tw[4], tw[5], nseg_tot, nstep_tot);
#endif
    if (tid == 0) { info[f] = failed; ucur[f] = u; }
    if (u >= 0)
        for (int i = tid; i < hcol_size(u + 1); i += HW_NT) Hpark[i] = H[i];
}

// returns -1 if this instance does not apply
extern "C" int axs_launch_hqr_win(int n, int nf, int bulges, int m_stop, const cplx* Hc, cplx* Hp, cplx* eig,
                                  int* info, int* ucur, cudaStream_t st)
{
    if (n > 144 || n < 64 || bulges > HW_MB) return -1;
    const size_t smem = (size_t)hcol_size(n) * sizeof(cplx);
    AXS_CUDA_OK(cudaFuncSetAttribute(k_hqr_win, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_hqr_win<<<nf, HW_NT, smem, st>>>(n, bulges, m_stop, Hc, Hp, eig, info, ucur);
    return (int)cudaGetLastError();
}
