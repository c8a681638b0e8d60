// A/B formulations of the early-deflation window routine of aed.cuh (same scratch object, same results):
// NOT part of the product build (AXS_AED_BUILD = 0).  Both run on the host through tests/host/warp_emul.h
// (32 fibers as lanes, collectives as rendezvous, randomised lane order as a race check) and were measured
// on B200 (tests/run_ab_aed2.sh, run_ab_aed3.sh; numbers in DESIGN.md 4.5):
//   aed_window_flat  42.5 - 49.8 ms per 2000 points against 37.9 - 42.8 ms for aed_window (windows 6 and 8)
//   aed_window_reg   55.8 ms (window 8); in a kernel that also carries aed_window: 201 ms
#pragma once
// =========================================================================================
// aed_window_flat: the shared-memory routine above as ONE loop around ONE rotation step.  A small
// warp-uniform state machine (the same one as in aed_window_reg below) selects the next rotation of the
// Schur sweeps (phase 0), of the spike test with its swaps (phase 1) and of the return to Hessenberg form
// (phase 2); the step itself is aed_window's: generate, overwrite up to two entries, rows, columns + Z, with
// three warp barriers.  Same arithmetic per entry as aed_window; the point is the code size - a single warp
// has nobody to hide its instruction fetches behind, and the loop it spins in should fit the 6 KB L0
// instruction cache of its SM sub-partition.
// Warp-style macros (identical on the host, where tests/host/warp_emul.h runs the 32 lanes as fibers):
// (the window has at most 8 x 8 entries: a lane owns at most ONE index of a range of <= 32, which spares the
// compiler-unrolled strided loops of AED_LANES)
#define AEDW_ONE(v, lo, hi) for (int v = (lo) + lane, once_ = 1; once_ && v < (hi); once_ = 0)
#ifndef AEDW_SYNC
#define AEDW_SYNC() __syncwarp()
#define AEDW_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define AEDW_CLZ(m) __clz((int)(m))
#endif

AED_FN void aed_window_flat(const cplx* H, int l, int u, int nw, int lane)
{
    AedScratch* const S = &g_aed;
    const int kw = u - nw + 1;
#pragma unroll 1
    for (int idx = lane; idx < nw * nw; idx += 32) {
        const int j = idx / nw, i = idx - j * nw;
        S->T[i][j] = (i <= j + 1) ? H[hcol_off(kw + j) + kw + i] : mk(0, 0);
        S->Z[i][j] = mk(i == j ? 1.0 : 0.0, 0.0);
    }
    const cplx spike = (kw > l) ? H[hcol_off(kw - 1) + kw] : mk(0, 0);
    const double sp1 = cabs1(spike);
    AEDW_SYNC();
    cplx sig = mk(0, 0);
    int phase = 0, k = -1;                              // k < 0: the phase has to choose its next group of rotations
    int iu = nw - 1, il = 0, its = 0, total = 0, ok = 1;     // phase 0
    int ns = nw, ilst = 0;                              // phase 1
    int jv = 0;                                         // phase 2: spike entry being annihilated

    for (;;) {
        // ---- control (warp-uniform; every lane reads the same shared entries) -------------------------
        int rk, rjlo, rihi, e1i = -1, e1j = 0, e2i = -1, e2j = 0, sva = -1;
        bool e1_is_r = false;
        cplx x, y, v1 = mk(0, 0), v2 = mk(0, 0);
        if (phase == 0) {
            if (k < 0) {
                if (iu < 1) { phase = 1; continue; }
                const bool neg = (lane >= 1 && lane <= iu) ? aed_negligible(S, nw, lane) : false;
                const unsigned m = AEDW_BALLOT(neg);          // (also: every lane has read the subdiagonal)
                il = m ? 31 - AEDW_CLZ(m) : 0;
                if (il > 0 && lane == 0) S->T[il][il - 1] = mk(0, 0);
                if (il == iu) { --iu; its = 0; AEDW_SYNC(); continue; }
                ++its; ++total;
                if (its > 30) { ok = 0; break; }
                sig = aed_shift2(S->T[iu - 1][iu - 1], S->T[iu - 1][iu], S->T[iu][iu - 1], S->T[iu][iu]);
                if (its == 10 || its == 20) sig = cadd(S->T[iu][iu], mk(0.75 * cabs1(S->T[iu][iu - 1]), 0));
                k = il;
            }
            const int gc = (k == il) ? il : k - 1;
            x = S->T[k][gc]; y = S->T[k + 1][gc];
            if (k == il) x = csub(x, sig);
            else { e1i = k; e1j = k - 1; e1_is_r = true; e2i = k + 1; e2j = k - 1; }
            rk = k; rjlo = k; rihi = (k + 2 < iu) ? k + 2 : iu;
            ++k;
            if (k >= iu) k = -1;
        } else if (phase == 1) {
            if (k < 0) {
                if (ilst >= ns) {
                    AEDW_ONE(j, 0, ns) S->shift[j] = S->T[j][j];      // undeflated window eigenvalues = next shifts
                    if (ns == nw) break;                       // nothing deflated: the caller leaves H alone
                    AEDW_ONE(j, 0, nw) S->sv[j] = (j < ns) ? cmulc(spike, S->Z[0][j]) : mk(0, 0);
                    AEDW_SYNC();
                    phase = 2; jv = ns - 1; k = -1;
                    continue;
                }
                double foo = cabs1(S->T[ns - 1][ns - 1]);
                if (foo == 0.0) foo = sp1;
                if (sp1 * cabs1(S->Z[0][ns - 1]) <= fmax(AED_SMALL, AED_EPS * foo)) { --ns; continue; }
                k = ns - 2;                                    // move T(ns-1, ns-1) up to position ilst by adjacent swaps
                if (k < ilst) { ++ilst; k = -1; continue; }
            }
            {
                const cplx t11 = S->T[k][k], t22 = S->T[k + 1][k + 1];
                x = S->T[k][k + 1]; y = csub(t22, t11);
                e1i = k; e1j = k; v1 = t22; e2i = k + 1; e2j = k + 1; v2 = t11;
            }
            rk = k; rjlo = k + 2; rihi = k - 1;                // rows: columns >= k+2, columns: rows < k
            --k;
            if (k < ilst) { ++ilst; k = -1; }
        } else {
            if (jv < 1) break;
            if (k < 0) {                                       // annihilate spike entry jv against entry jv-1
                x = S->sv[jv - 1]; y = S->sv[jv];
                sva = jv - 1;
                rk = jv - 1; rjlo = jv - 1; rihi = (jv + 1 < ns - 1) ? jv + 1 : ns - 1;
                k = jv;
            } else {                                           // chase the bulge T(k+1, k-1) out of the bottom
                x = S->T[k][k - 1]; y = S->T[k + 1][k - 1];
                e1i = k; e1j = k - 1; e1_is_r = true; e2i = k + 1; e2j = k - 1;
                rk = k; rjlo = k; rihi = (k + 2 < ns - 1) ? k + 2 : ns - 1;
                ++k;
            }
            if (k + 1 > ns - 1) { --jv; k = -1; }
        }
        // ---- the rotation step ------------------------------------------------------------------------
        AEDW_SYNC();                                      // x, y and the control inputs read by every lane before they change
        double c; cplx s, r;
        aed_rot(x, y, c, s, r);
        if (lane == 0) {
            if (e1_is_r) v1 = r;
            if (e1i >= 0) S->T[e1i][e1j] = v1;
            if (e2i >= 0) S->T[e2i][e2j] = v2;
            if (sva >= 0) { S->sv[sva] = r; S->sv[sva + 1] = mk(0, 0); }
        }
        AEDW_ONE(j, rjlo, nw) aed_rows(c, s, S->T[rk][j], S->T[rk + 1][j]);
        AEDW_SYNC();
        AEDW_ONE(ii, 0, 2 * nw) {
            if (ii < nw) { if (ii <= rihi) aed_cols(c, s, S->T[ii][rk], S->T[ii][rk + 1]); }
            else aed_cols(c, s, S->Z[ii - nw][rk], S->Z[ii - nw][rk + 1]);
        }
        AEDW_SYNC();
    }
    if (lane == 0) {
        S->ok = ok; S->iters = total; S->ns = ok ? ns : nw;
        S->spike_new = (ok && ns > 0 && ns < nw) ? S->sv[0] : mk(0, 0);
    }
    AEDW_SYNC();
}

// =========================================================================================
// Register-resident window (aed_window_reg): the same four steps, but T and Z never touch shared
// memory while the warp works on them.  Lane L = 4 i + p owns the entries (i, 2p) and (i, 2p + 1) of T
// and of Z in registers:
//   * a row rotation (k, k+1) is ONE exchange between the lanes of row k and of row k+1 (lane +- 4)
//     followed by a branch-free update of the own entries,
//   * a column rotation (k, k+1) needs no exchange for even k (both columns sit in the same lane) and one
//     exchange between neighbouring lanes for odd k,
//   * the rotation itself is generated redundantly by every lane from two broadcast entries, so nothing
//     has to be sent back,
//   * the deflation scan gathers the three diagonals into lanes 1..iu (one ballot), the shift comes from
//     four broadcast entries.
// The dependent chain of a rotation step is two shuffles + the rotation + two 6-instruction updates
// instead of a shared-memory round trip with three warp barriers; control flow is warp-uniform throughout
// (every lane derives the same decisions from the same broadcast values), which is what the host build
// checks by running this very source on an emulated warp (tests/host/warp_emul.h: 32 fibers, collectives
// as rendezvous; AEDR_SHFL / AEDR_BALLOT / AEDR_CLZ are then supplied by the includer).
// Results go to the same scratch object (T, Z, shift, spike_new, ns, ok, iters) as aed_window's.
#ifndef AEDR_SHFL
__device__ __forceinline__ cplx aedr_shfl_dev(cplx v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
#define AEDR_SHFL(v, src) aedr_shfl_dev((v), (src))
#define AEDR_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define AEDR_CLZ(m) __clz((int)(m))
#endif

// entry (ii, jj) of T / Z for warp-uniform ii, jj: broadcast from its owner lane
#define AEDR_GET(ii, jj) AEDR_SHFL((((jj) & 1) ? t1 : t0), 4 * (ii) + ((jj) >> 1))
#define AEDR_GETZ(ii, jj) AEDR_SHFL((((jj) & 1) ? z1 : z0), 4 * (ii) + ((jj) >> 1))
#define AEDR_SET(ii, jj, v) do { if (i == (ii) && p == ((jj) >> 1)) { if ((jj) & 1) t1 = (v); else t0 = (v); } } while (0)

// One similarity rotation in the plane (k, k+1): rows k, k+1 of the columns >= jlo, then columns k, k+1
// of the rows <= ihi of T and of every row of Z (aed_apply for the register layout).  The row exchange
// does not depend on the rotation, so it is issued BEFORE the rotation is generated (AEDR_ROWX) and its
// latency hides behind the rsqrt chain; q0 / q1 are the partner row's entries (entries of columns < jlo
// may be stale by then - they are not used).
#define AEDR_ROWX(k, q0, q1)                                                                            \
    const int rowx_src_ = (i == (k)) ? lane + 4 : ((i == (k) + 1) ? lane - 4 : lane);                      \
    const cplx q0 = AEDR_SHFL(t0, rowx_src_), q1 = AEDR_SHFL(t1, rowx_src_)

AED_INL void aedr_apply(cplx& t0, cplx& t1, cplx& z0, cplx& z1, int lane, int k, int jlo, int ihi, double c, cplx s,
                        cplx q0, cplx q1)
{
    const int i = lane >> 2, p = lane & 3;
    {
        const bool top = (i == k), bot = (i == k + 1);
        // top: c own + s other;  bottom: c own - conj(s) other
        const cplx se = top ? s : mk(-s.x, s.y);
        if (top || bot) {
            if (2 * p >= jlo) t0 = cfma(se, q0, cscale(t0, c));
            if (2 * p + 1 >= jlo) t1 = cfma(se, q1, cscale(t1, c));
        }
    }
    if ((k & 1) == 0) {
        if (p == (k >> 1)) {
            if (i <= ihi) {
                const cplx n0 = cfmac(t1, s, cscale(t0, c)), n1 = cfms(s, t0, cscale(t1, c));
                t0 = n0; t1 = n1;
            }
            const cplx m0 = cfmac(z1, s, cscale(z0, c)), m1 = cfms(s, z0, cscale(z1, c));
            z0 = m0; z1 = m1;
        }
    } else {
        // column k is the odd entry of pair (k-1)/2, column k+1 the even entry of pair (k+1)/2
        const bool isa = (p == (k >> 1)), isb = (p == ((k + 1) >> 1));
        const int src = isa ? lane + 1 : (isb ? lane - 1 : lane);
        const cplx rt = AEDR_SHFL(isa ? t1 : t0, src), rz = AEDR_SHFL(isa ? z1 : z0, src);
        if (isa) {
            if (i <= ihi) t1 = cfmac(rt, s, cscale(t1, c));
            z1 = cfmac(rz, s, cscale(z1, c));
        } else if (isb) {
            if (i <= ihi) t0 = cfms(s, rt, cscale(t0, c));
            z0 = cfms(s, rz, cscale(z0, c));
        }
    }
}

// The routine is ONE loop around ONE rotation step (row exchange, rotation, two entry updates, similarity
// update): a small warp-uniform state machine selects the next rotation of the Schur sweeps (phase 0), of the
// spike test with its swaps (phase 1) and of the return to Hessenberg form (phase 2).  A single warp has
// nobody to hide its instruction fetches behind, so the code it loops over has to stay small (the four
// inlined copies of the step of a straightforward transcription ran 8 x slower on B200 than this form's
// predecessor in shared memory - DESIGN.md 4.5).
AED_FN void aed_window_reg(const cplx* H, int l, int u, int nw, int lane)
{
    AedScratch* const S = &g_aed;
    const int kw = u - nw + 1;
    const int i = lane >> 2, p = lane & 3, j0 = 2 * p, j1 = 2 * p + 1;
    cplx t0 = (i < nw && j0 < nw && i <= j0 + 1) ? H[hcol_off(kw + j0) + kw + i] : mk(0, 0);
    cplx t1 = (i < nw && j1 < nw && i <= j1 + 1) ? H[hcol_off(kw + j1) + kw + i] : mk(0, 0);
    cplx z0 = mk(i == j0 ? 1.0 : 0.0, 0.0), z1 = mk(i == j1 ? 1.0 : 0.0, 0.0);
    const cplx spike = (kw > l) ? H[hcol_off(kw - 1) + kw] : mk(0, 0);
    const double sp1 = cabs1(spike);
    cplx sv = mk(0, 0), sig = mk(0, 0);                 // spike vector entry of lane j (phase 2); shift of the sweep
    int phase = 0, k = -1;                              // k < 0: the phase has to choose its next group of rotations
    int iu = nw - 1, il = 0, its = 0, total = 0, ok = 1;     // phase 0
    int ns = nw, ilst = 0;                              // phase 1
    int jv = 0;                                         // phase 2: spike entry being annihilated

    for (;;) {
        // ---- control (warp-uniform): the next rotation = plane rk, rows from column rjlo, columns up to row rihi,
        //      generated from (x, y); up to two entries (e1, e2) are overwritten before the similarity update
        int rk, rjlo, rihi, e1i = -1, e1j = 0, e2i = -1, e2j = 0, sva = -1;
        bool e1_is_r = false;
        cplx x, y, v1 = mk(0, 0), v2 = mk(0, 0);
        if (phase == 0) {
            if (k < 0) {
                if (iu < 1) { phase = 1; continue; }
                // lane q (1 <= q <= iu) tests T(q, q-1): the three diagonals are gathered into the lanes
                {
                    const int q = (lane < nw) ? lane : nw - 1, qm = (q > 0) ? q - 1 : 0;
                    const int s_d = 4 * q + (q >> 1), s_sub = 4 * q + (qm >> 1), s_sup = 4 * qm + (q >> 1);
                    const cplx d_a = AEDR_SHFL(t0, s_d), d_b = AEDR_SHFL(t1, s_d);
                    const cplx b_a = AEDR_SHFL(t0, s_sub), b_b = AEDR_SHFL(t1, s_sub);
                    const cplx p_a = AEDR_SHFL(t0, s_sup), p_b = AEDR_SHFL(t1, s_sup);
                    const cplx hkk = (q & 1) ? d_b : d_a;             // T(q, q)
                    const cplx hsub = (qm & 1) ? b_b : b_a;           // T(q, q-1)
                    const cplx hsup = (q & 1) ? p_b : p_a;            // T(q-1, q)
                    const cplx hk1 = AEDR_SHFL(hkk, (lane + 31) & 31);      // T(q-1, q-1)
                    const cplx hsubm = AEDR_SHFL(hsub, (lane + 31) & 31);   // T(q-1, q-2)
                    const cplx hsubp = AEDR_SHFL(hsub, (lane + 1) & 31);    // T(q+1, q)
                    bool neg = false;
                    if (lane >= 1 && lane <= iu) {
                        const double sub = cabs1(hsub);
                        if (sub <= AED_SMALL) neg = true;
                        else {
                            double tst = cabs1(hk1) + cabs1(hkk);
                            if (tst == 0.0) {
                                if (q - 2 >= 0) tst += cabs1(hsubm);
                                if (q + 1 <= nw - 1) tst += cabs1(hsubp);
                            }
                            if (sub <= AED_EPS * tst) {
                                const double up = cabs1(hsup);
                                const double ab = fmax(sub, up), ba = fmin(sub, up);
                                const double dd = cabs1(csub(hk1, hkk));
                                const double aa = fmax(cabs1(hkk), dd), bb = fmin(cabs1(hkk), dd);
                                const double sm = aa + ab;
                                if (ba * (ab / sm) <= fmax(AED_SMALL, AED_EPS * (bb * (aa / sm)))) neg = true;
                            }
                        }
                    }
                    const unsigned m = AEDR_BALLOT(neg);
                    il = m ? 31 - AEDR_CLZ(m) : 0;
                }
                if (il > 0) AEDR_SET(il, il - 1, mk(0, 0));
                if (il == iu) { --iu; its = 0; continue; }
                ++its; ++total;
                if (its > 30) { ok = 0; break; }
                {
                    const cplx a11 = AEDR_GET(iu - 1, iu - 1), a12 = AEDR_GET(iu - 1, iu);
                    const cplx a21 = AEDR_GET(iu, iu - 1), a22 = AEDR_GET(iu, iu);
                    sig = aed_shift2(a11, a12, a21, a22);
                    if (its == 10 || its == 20) sig = cadd(a22, mk(0.75 * cabs1(a21), 0));
                }
                k = il;
            }
            // rotation k of the sweep [il, iu]: the first one introduces the bulge, the others chase it
            const int gc = (k == il) ? il : k - 1;
            x = AEDR_GET(k, gc); y = AEDR_GET(k + 1, gc);
            if (k == il) x = csub(x, sig);
            else { e1i = k; e1j = k - 1; e1_is_r = true; e2i = k + 1; e2j = k - 1; }
            rk = k; rjlo = k; rihi = (k + 2 < iu) ? k + 2 : iu;
            ++k;
            if (k >= iu) k = -1;
        } else if (phase == 1) {
            if (k < 0) {
                if (ilst >= ns) {
                    // the undeflated window eigenvalues are the next shifts
                    if (i == j0 && j0 < ns) S->shift[j0] = t0;
                    if (i == j1 && j1 < ns) S->shift[j1] = t1;
                    if (ns == nw) break;                       // nothing deflated: the caller leaves H alone
                    {
                        const int j = lane & 7;
                        const cplx za = AEDR_SHFL(z0, j >> 1), zb = AEDR_SHFL(z1, j >> 1);      // Z(0, j)
                        if (lane < 8 && lane < ns) sv = cmulc(spike, (j & 1) ? zb : za);
                    }
                    phase = 2; jv = ns - 1; k = -1;
                    continue;
                }
                double foo = cabs1(AEDR_GET(ns - 1, ns - 1));
                if (foo == 0.0) foo = sp1;
                if (sp1 * cabs1(AEDR_GETZ(0, ns - 1)) <= fmax(AED_SMALL, AED_EPS * foo)) { --ns; continue; }
                k = ns - 2;                                    // move T(ns-1, ns-1) up to position ilst by adjacent swaps
                if (k < ilst) { ++ilst; k = -1; continue; }
            }
            {
                const cplx t11 = AEDR_GET(k, k), t22 = AEDR_GET(k + 1, k + 1);
                x = AEDR_GET(k, k + 1); y = csub(t22, t11);
                e1i = k; e1j = k; v1 = t22; e2i = k + 1; e2j = k + 1; v2 = t11;
            }
            rk = k; rjlo = k + 2; rihi = k - 1;                // rows: columns >= k+2, columns: rows < k
            --k;
            if (k < ilst) { ++ilst; k = -1; }
        } else {
            if (jv < 1) break;
            if (k < 0) {                                       // annihilate spike entry jv against entry jv-1
                x = AEDR_SHFL(sv, jv - 1); y = AEDR_SHFL(sv, jv);
                sva = jv - 1;
                rk = jv - 1; rjlo = jv - 1; rihi = (jv + 1 < ns - 1) ? jv + 1 : ns - 1;
                k = jv;
            } else {                                           // chase the bulge T(k+1, k-1) out of the bottom
                x = AEDR_GET(k, k - 1); y = AEDR_GET(k + 1, k - 1);
                e1i = k; e1j = k - 1; e1_is_r = true; e2i = k + 1; e2j = k - 1;
                rk = k; rjlo = k; rihi = (k + 2 < ns - 1) ? k + 2 : ns - 1;
                ++k;
            }
            if (k + 1 > ns - 1) { --jv; k = -1; }
        }
        // ---- the rotation step ------------------------------------------------------------------------
        AEDR_ROWX(rk, q0, q1);
        double c; cplx s, r;
        aed_rot(x, y, c, s, r);
        if (e1_is_r) v1 = r;
        AEDR_SET(e1i, e1j, v1);
        AEDR_SET(e2i, e2j, v2);
        if (lane == sva) sv = r;
        if (lane == sva + 1 && sva >= 0) sv = mk(0, 0);
        aedr_apply(t0, t1, z0, z1, lane, rk, rjlo, rihi, c, s, q0, q1);
    }
    if (lane == 0) {
        S->ok = ok; S->iters = total; S->ns = ok ? ns : nw;
        S->spike_new = (ok && ns > 0 && ns < nw) ? sv : mk(0, 0);
    }
    S->T[i][j0] = t0; S->T[i][j1] = t1;
    S->Z[i][j0] = z0; S->Z[i][j1] = z1;
    AED_SYNC();
}
