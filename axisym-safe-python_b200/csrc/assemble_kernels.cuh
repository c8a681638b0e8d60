// k_element_matrices, k_scatter_operators, k_build_S: kernel 1 of the SAFE sweep (see assemble.cu for the
// launchers and the overview).  A header of its own so that tests/host/assemble_emul.cpp can compile this very
// source as host code and run it on an emulated thread block (tests/host/cta_emul.h) without a GPU.
#pragma once
#include "common.cuh"
// the dynamic shared memory of the CTA (the host emulation supplies its own definition: a plain buffer)
#ifndef AXS_DYNAMIC_SMEM
#define AXS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define EL_MAX_M 64          // nodes per element supported by the shared-memory staging
#define TWO_PI 6.283185307179586

struct PointGeom {           // per quadrature point
    cplx invr;               // 1/r   (1/r~ in a PML, 0 on the axis)
    cplx ginv;               // 1/gamma (1 outside a PML)
    cplx wt;                 // 2 pi w J r  (r~ gamma in a PML, J*rat in a core element)
    int axis;
};

// strain-operator columns: L e_c, L_r e_c, L_theta e_c, L_z e_c as (row index, sign)
__device__ __forceinline__ void bvec1(int c, cplx dirac_invr, cplx nr_g, bool axis, cplx nr, cplx* b) {
    // b1 = L e_c * (delta/r) + L_r e_c * (N_r/gamma);  on the axis (L + L_r) e_c * N_r
    for (int s = 0; s < 6; ++s) b[s] = mk(0, 0);
    cplx radial = axis ? nr : nr_g;
    cplx hoop = axis ? nr : dirac_invr;
    if (c == 0) { b[0] = radial; b[1] = hoop; }
    else if (c == 1) { b[5] = csub(radial, hoop); }
    else { b[4] = radial; }
}
__device__ __forceinline__ void bvec2(int c, cplx dirac_invr, bool axis, cplx nr, cplx* b) {
    for (int s = 0; s < 6; ++s) b[s] = mk(0, 0);
    cplx v = axis ? nr : dirac_invr;
    if (c == 0) b[5] = v; else if (c == 1) b[1] = v; else b[3] = v;
}
__device__ __forceinline__ void bvec3(int c, double dirac, cplx* b) {
    for (int s = 0; s < 6; ++s) b[s] = mk(0, 0);
    cplx v = mk(dirac, 0);
    if (c == 0) b[4] = v; else if (c == 1) b[3] = v; else b[2] = v;
}
// y = C b for isotropic C = (1 + i eta) iso(lambda, mu)
__device__ __forceinline__ void apply_C(cplx lam, cplx mu, const cplx* b, cplx* y) {
    cplx tr = cadd(cadd(b[0], b[1]), b[2]);
    cplx ltr = cmul(lam, tr);
    cplx mu2 = cscale(mu, 2.0);
    for (int s = 0; s < 3; ++s) y[s] = cadd(ltr, cmul(mu2, b[s]));
    for (int s = 3; s < 6; ++s) y[s] = cmul(mu, b[s]);
}
__device__ __forceinline__ cplx dot6(const cplx* a, const cplx* b) {
    cplx acc = mk(0, 0);
    for (int s = 0; s < 6; ++s) acc = cfma(a[s], b[s], acc);
    return acc;
}

__global__ void __launch_bounds__(128)
k_element_matrices(int n_elem, const int* __restrict__ edesc_i, const double* __restrict__ edesc_d,
                   const double* __restrict__ tables, const double* __restrict__ nodes,
                   cplx* __restrict__ out)
{
    AXS_DYNAMIC_SMEM(smem_raw);
    const int warps_per_cta = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * warps_per_cta + warp;
    if (e >= n_elem) return;
    const int* di = edesc_i + e * 8;
    const double* dd = edesc_d + e * 12;
    const int kind = di[0], family = di[1], pml = di[2], m = di[3];
    const double* ksi = tables + di[4];
    const double* wq = ksi + m;
    const double* D = wq + m;
    const double* rn = nodes + di[5];
    cplx* o = out + (long long)di[6];
    const int nd = (kind == 0 ? 3 : 1) * m;
    const long long msz = (long long)nd * nd;

    // shared staging for this warp: Nr[m*m] (real), geom[m]
    size_t per_warp = (size_t)EL_MAX_M * EL_MAX_M * sizeof(double) + EL_MAX_M * sizeof(PointGeom)
                      + EL_MAX_M * sizeof(double);
    double* Nr = (double*)(smem_raw + warp * per_warp);
    PointGeom* pg = (PointGeom*)(Nr + EL_MAX_M * EL_MAX_M);
    double* Jq = (double*)(pg + EL_MAX_M);

    const double rho = dd[3];
    for (int q = lane; q < m; q += 32) {
        double J = 0.0;
        for (int a = 0; a < m; ++a) J = fma(D[a * m + q], rn[a], J);
        double r = rn[q];
        PointGeom g;
        g.axis = (family == 1) && (fabs(r) < 5e-8);         // np.round(r, 7) == 0
        if (pml == 0) {
            g.ginv = mk(1.0, 0.0);
            if (family == 1) {
                double rat = g.axis ? J : r / (1.0 + ksi[q]);
                g.invr = g.axis ? mk(0, 0) : mk(1.0 / r, 0.0);
                g.wt = mk(((TWO_PI * wq[q]) * J) * rat, 0.0);
            } else {
                g.invr = mk(1.0 / r, 0.0);
                g.wt = mk(((TWO_PI * wq[q]) * J) * r, 0.0);
            }
        } else {
            const double start = dd[5], thk = dd[6], ga = dd[7], gb = dd[8];
            cplx gam, rt;
            if (pml == 1) {                                   // exponential profile
                double s = (r - start) / thk;
                double ea = exp(ga * s), eb = exp(gb * s);
                gam = mk(ea, -(eb - 1.0));
                rt = mk(start + thk / ga * (ea - 1.0), -thk / gb * (eb - 1.0) + s * thk);
            } else {                                          // quadratic profile, gamma = mean value
                double s = fabs(r - start) / thk;
                cplx gm1 = mk(ga - 1.0, gb);
                gam = cadd(mk(1.0, 0.0), cscale(gm1, 3.0 * s * s));
                double d3 = (start - r) * (start - r) * (start - r) / (thk * thk);
                rt = cadd(cscale(gm1, -d3), mk(r, 0.0));
            }
            g.invr = cdiv(mk(1.0, 0.0), rt);
            g.ginv = cdiv(mk(1.0, 0.0), gam);
            g.wt = cmul(cmul(mk((TWO_PI * wq[q]) * J, 0.0), rt), gam);
        }
        pg[q] = g;
        Jq[q] = J;
    }
    __syncwarp();
    for (int idx = lane; idx < m * m; idx += 32) {
        int q = idx % m;
        Nr[idx] = D[idx] / Jq[q];                             // Nr[a*m+q] = dN_a/dr at point q
    }
    __syncwarp();

    // zero the whole output block first (slots that stay zero for this element kind)
    for (long long idx = lane; idx < 10 * msz; idx += 32) o[idx] = mk(0, 0);
    __syncwarp();

    if (kind == 0) {
        const double eta = dd[2];
        const cplx lam = mk(dd[0], dd[0] * eta), mu = mk(dd[1], dd[1] * eta);
        for (int idx = lane; idx < nd * nd; idx += 32) {
            const int i = idx / nd, j = idx % nd;
            const int a = i / 3, ca = i % 3, b = j / 3, cb = j % 3;
            cplx k1 = mk(0, 0), k5 = k1, k6 = k1, f1 = k1, f2 = k1, f3 = k1, f1t = k1, f2t = k1, f3t = k1;
            for (int q = 0; q < m; ++q) {
                const PointGeom g = pg[q];
                const bool ia = (a == q), jb = (b == q);
                if (!ia && !jb && !g.axis) {
                    // only the N_r x N_r part of K1 survives away from both nodes
                    if ((ca == cb)) {
                        // b1_i = Lr e_ca Nr_a/g, b1_j likewise: (Lr e_c)^T C (Lr e_c) = C00 (c=0) or mu
                        cplx cc = (ca == 0) ? cadd(lam, cscale(mu, 2.0)) : mu;
                        cplx v = cmul(cscale(cmul(g.ginv, g.ginv), Nr[a * m + q] * Nr[b * m + q]), cc);
                        k1 = cfma(g.wt, v, k1);
                    }
                    continue;
                }
                cplx bi1[6], bi2[6], bi3[6], bj1[6], bj2[6], bj3[6], c1[6], c2[6], c3[6];
                cplx nra = mk(Nr[a * m + q], 0), nrb = mk(Nr[b * m + q], 0);
                bvec1(ca, ia ? g.invr : mk(0, 0), cmul(nra, g.ginv), g.axis, nra, bi1);
                bvec2(ca, ia ? g.invr : mk(0, 0), g.axis, nra, bi2);
                bvec3(ca, ia ? 1.0 : 0.0, bi3);
                bvec1(cb, jb ? g.invr : mk(0, 0), cmul(nrb, g.ginv), g.axis, nrb, bj1);
                bvec2(cb, jb ? g.invr : mk(0, 0), g.axis, nrb, bj2);
                bvec3(cb, jb ? 1.0 : 0.0, bj3);
                apply_C(lam, mu, bj1, c1);
                apply_C(lam, mu, bj2, c2);
                apply_C(lam, mu, bj3, c3);
                k1 = cfma(g.wt, dot6(bi1, c1), k1);
                f2 = cfma(g.wt, dot6(bi1, c2), f2);
                f1 = cfma(g.wt, dot6(bi1, c3), f1);
                k5 = cfma(g.wt, dot6(bi2, c2), k5);
                f3 = cfma(g.wt, dot6(bi2, c3), f3);
                k6 = cfma(g.wt, dot6(bi3, c3), k6);
                f2t = cfma(g.wt, dot6(bi2, c1), f2t);       // KF2[j,i]
                f1t = cfma(g.wt, dot6(bi3, c1), f1t);       // KF1[j,i]
                f3t = cfma(g.wt, dot6(bi3, c2), f3t);       // KF3[j,i]
            }
            o[0 * msz + idx] = k1;
            o[1 * msz + idx] = csub(f2t, f2);                 // K2 = KF2^T - KF2
            o[2 * msz + idx] = csub(f1t, f1);                 // K3 = KF1^T - KF1
            o[3 * msz + idx] = cadd(f3t, f3);                 // K4 = KF3^T + KF3
            o[4 * msz + idx] = k5;
            o[5 * msz + idx] = k6;
            o[6 * msz + idx] = f1;
            o[7 * msz + idx] = f2;
            o[8 * msz + idx] = f3;
        }
        // mass matrix: diagonal; non-PML solids reproduce the reference's complex64 accumulation
        for (int i = lane; i < nd; i += 32) {
            const int q = i / 3;
            cplx mval;
            if (pml == 0) {
                double last = (family == 1) ? (pg[q].axis ? Jq[q] : rn[q] / (1.0 + ksi[q])) : rn[q];
                double term = (((TWO_PI * wq[q]) * rho) * Jq[q]) * last;
                mval = mk((double)__double2float_rn(term), 0.0);
            } else {
                mval = cscale(pg[q].wt, rho);
            }
            o[9 * msz + (long long)i * nd + i] = mval;
        }
    } else {
        const double cp2 = dd[4];
        for (int idx = lane; idx < nd * nd; idx += 32) {
            const int a = idx / nd, b = idx % nd;
            cplx k1 = mk(0, 0), k5 = mk(0, 0);
            for (int q = 0; q < m; ++q) {
                const PointGeom g = pg[q];
                cplx sc = cscale(g.wt, -rho);
                double nn = Nr[a * m + q] * Nr[b * m + q];
                k1 = cfma(sc, cscale(cmul(g.ginv, g.ginv), nn), k1);
                if (g.axis) k5 = cfma(sc, mk(nn, 0), k5);
                else if (a == q && b == q) k5 = cfma(sc, cmul(g.invr, g.invr), k5);
            }
            o[0 * msz + idx] = k1;
            o[4 * msz + idx] = k5;
            if (a == b) {
                cplx sc = cscale(pg[a].wt, -rho);
                o[5 * msz + idx] = sc;
                o[9 * msz + idx] = mk(sc.x / cp2, sc.y / cp2);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_cadd(cplx* p, cplx v) {
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}

__global__ void __launch_bounds__(128)
k_scatter_operators(int n_elem, const int* __restrict__ edesc_i, const cplx* __restrict__ elem,
                    const int* __restrict__ lm, const int* __restrict__ lm_off,
                    const int* __restrict__ tclass, int n_circ, int N, int hb,
                    cplx* K0s, cplx* Cb, cplx* Mb, cplx* K6d)
{
    const int warps_per_cta = blockDim.x >> 5;
    const int e = blockIdx.x * warps_per_cta + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= n_elem) return;
    const int* di = edesc_i + e * 8;
    const int nd = (di[0] == 0 ? 3 : 1) * di[3];
    const long long msz = (long long)nd * nd;
    const cplx* o = elem + (long long)di[6];
    const int* map = lm + lm_off[e];
    const int W = 2 * hb + 1;
    const double dn = (double)n_circ, dn2 = dn * dn;
    for (int idx = lane; idx < nd * nd; idx += 32) {
        const int gi = map[idx / nd], gj = map[idx % nd];
        if (gi < 0 || gj < 0) continue;
        const int ti = tclass[gi], tj = tclass[gj];
        // Khat[r,c] = i T_r conj(T_c) K[r,c]:  (1,1)->i  (1,i)->1  (i,1)->-1  (i,i)->i
        cplx tf = (ti == tj) ? mk(0, 1) : (ti == 0 ? mk(1, 0) : mk(-1, 0));
        cplx k0 = cadd(cadd(o[idx], cscale(o[4 * msz + idx], dn2)), cscale(cmul(tf, o[1 * msz + idx]), dn));
        cplx cc = cadd(cscale(o[3 * msz + idx], dn), cmul(tf, o[2 * msz + idx]));
        const long long pos = (long long)gi * W + (gj - gi + hb);
        atomic_cadd(K0s + pos, k0);
        atomic_cadd(Cb + pos, cc);
        atomic_cadd(Mb + pos, o[9 * msz + idx]);
        if (gi == gj) atomic_cadd(K6d + gi, o[5 * msz + idx]);
    }
}

// ----------------------------------------------------------------------------------------
// S[f] column-major n2 x n2.  grid = (ceil(n2*n2 / (256*4)), nf), block 256.
__device__ __forceinline__ cplx band_at(const cplx* B, int W, int hb, int r, int c) {
    int d = c - r;
    return (d < -hb || d > hb) ? mk(0, 0) : B[(long long)r * W + d + hb];
}

__device__ __forceinline__ cplx S_entry(int N, int hb, const cplx* K0s, const cplx* Cb, const cplx* Mb,
                                        const cplx* K1c, const cplx* K6d, double w, int r, int c)
{
    const int W = 2 * hb + 1;
    if (r >= N) return mk((r - N == c) ? 1.0 : 0.0, 0.0);
    cplx num;
    if (c < N) {
        if (c - r < -hb || c - r > hb) return mk(0, 0);
        num = band_at(Cb, W, hb, r, c);
    } else {
        int cc = c - N;
        if (cc - r < -hb || cc - r > hb) return mk(0, 0);
        num = band_at(K0s, W, hb, r, cc);
        cplx mm = band_at(Mb, W, hb, r, cc);
        num = cfma(mk(-w * w, 0), mm, num);
        if (K1c) num = cfma(mk(0, w), band_at(K1c, W, hb, r, cc), num);
    }
    return cneg(cdiv(num, K6d[r]));
}

__global__ void __launch_bounds__(256)
k_build_S(int N, int hb, const cplx* __restrict__ K0s, const cplx* __restrict__ Cb,
          const cplx* __restrict__ Mb, const cplx* __restrict__ K1c, const cplx* __restrict__ K6d,
          const double* __restrict__ omega, cplx* __restrict__ S)
{
    const int n2 = 2 * N;
    const int f = blockIdx.y;
    const double w = omega[f];
    cplx* Sf = S + (long long)f * n2 * n2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n2 * n2; idx += gridDim.x * blockDim.x) {
        int c = idx / n2, r = idx % n2;
        Sf[idx] = S_entry(N, hb, K0s, Cb, Mb, K1c, K6d, w, r, c);
    }
}

