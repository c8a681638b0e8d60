"""NumPy prototype of the windowed multi-bulge QR used by csrc/eig_big.cu (k_hqr_big).

A chain of nb single-shift bulges (3 rows apart) is chased through a diagonal window in
chunks of s macro steps ("phase D", sequential semantics of k_hqr_mb), the rotations are
recorded, and the off-diagonal strips are updated afterwards bulge-major ("phase O").
Validates window bounds, ordering and convergence against numpy.linalg.eigvals.
Not imported by the product or the tests.
"""
import numpy as np

EPS = np.finfo(float).eps
SMALL = np.finfo(float).tiny / EPS


def cabs1(z):
    return abs(z.real) + abs(z.imag)


def make_rot(x, y):
    r2 = abs(x) ** 2 + abs(y) ** 2
    if r2 > 0:
        ir = 1 / np.sqrt(r2)
        return np.conj(x) * ir, np.conj(y) * ir, r2 * ir
    return 1.0 + 0j, 0j, 0.0


def chase_batch(H, l, u, shifts, s=24, stats=None):
    nb = len(shifts)
    m = u - l + 1
    nsteps = (m - 1) + (nb - 1) * 3
    rot = [None] * nb                       # current rotation of every live bulge
    rot[0] = make_rot(H[l, l] - shifts[0], H[l + 1, l])
    for t0 in range(0, nsteps, s):
        t1 = min(t0 + s, nsteps)
        ks = [l + t - 3 * b for t in range(t0, t1) for b in range(nb) if l <= l + t - 3 * b <= u - 1]
        kmin, kmax = min(ks), max(ks)
        wlo, whi = max(l, kmin - 1), min(u, kmax + 2)
        if stats is not None:
            stats.append(whi - wlo + 1)
        rec = {}
        # ---- phase D ---------------------------------------------------------------------
        for t in range(t0, t1):
            nxt = {}
            for b in range(nb):
                k = l + t - 3 * b
                if not (l <= k <= u - 1):
                    continue
                a, bb, r = rot[b]
                rec[(t, b)] = (a, bb)
                # R: rows k, k+1 for columns k-1 .. whi
                if k - 1 >= l:
                    assert k - 1 >= wlo
                    H[k, k - 1] = r
                    H[k + 1, k - 1] = 0
                cols = np.arange(k, whi + 1)
                p, q = H[k, cols].copy(), H[k + 1, cols].copy()
                H[k, cols] = a * p + bb * q
                H[k + 1, cols] = np.conj(a) * q - np.conj(bb) * p
            for b in range(nb):
                k = l + t - 3 * b
                if not (l <= k <= u - 1):
                    continue
                a, bb, r = rot[b]
                # C: columns k, k+1 for rows wlo .. min(k+2, u)
                top = min(k + 2, u)
                assert top <= whi
                rows = np.arange(wlo, top + 1)
                c0, c1 = H[rows, k].copy(), H[rows, k + 1].copy()
                H[rows, k] = c0 * np.conj(a) + c1 * np.conj(bb)
                H[rows, k + 1] = c1 * a - c0 * bb
                if k + 2 <= u:
                    nxt[b] = make_rot(H[k + 1, k], H[k + 2, k])
            for b, g in nxt.items():
                rot[b] = g
            if (t + 1) % 3 == 0 and (t + 1) // 3 < nb:
                bn = (t + 1) // 3
                rot[bn] = make_rot(H[l, l] - shifts[bn], H[l + 1, l])
        # ---- phase O (bulge-major) ---------------------------------------------------------
        for b in range(nb):
            for t in range(t0, t1):
                if (t, b) not in rec:
                    continue
                k = l + t - 3 * b
                a, bb = rec[(t, b)]
                cols = np.arange(whi + 1, u + 1)
                if len(cols):
                    p, q = H[k, cols].copy(), H[k + 1, cols].copy()
                    H[k, cols] = a * p + bb * q
                    H[k + 1, cols] = np.conj(a) * q - np.conj(bb) * p
                rows = np.arange(l, wlo)
                if len(rows):
                    c0, c1 = H[rows, k].copy(), H[rows, k + 1].copy()
                    H[rows, k] = c0 * np.conj(a) + c1 * np.conj(bb)
                    H[rows, k + 1] = c1 * a - c0 * bb


def deflation_l(H, u):
    l = 0
    for k in range(u, 0, -1):
        sub = cabs1(H[k, k - 1])
        neg = False
        if sub <= SMALL:
            neg = True
        else:
            tst = cabs1(H[k - 1, k - 1]) + cabs1(H[k, k])
            if sub <= EPS * tst:
                up = cabs1(H[k - 1, k])
                ab, ba = max(sub, up), min(sub, up)
                dd = cabs1(H[k - 1, k - 1] - H[k, k])
                aa, bb = max(cabs1(H[k, k]), dd), min(cabs1(H[k, k]), dd)
                sm = aa + ab
                if ba * (ab / sm) <= max(SMALL, EPS * (bb * (aa / sm))):
                    neg = True
        if neg:
            l = k
            break
    return l


def qr_eigs(H, nbmax=8, s=24):
    H = H.copy()
    n = H.shape[0]
    eig = np.zeros(n, complex)
    u, its, batches, wsz = n - 1, 0, 0, []
    while u >= 0:
        l = deflation_l(H, u)
        if l > 0:
            H[l, l - 1] = 0
        if l == u:
            eig[u] = H[u, u]
            u -= 1
            its = 0
            continue
        its += 1
        m = u - l + 1
        nb = min((m - 2) // 3 + 1, nbmax)
        sh = []
        for q in range((nb + 1) // 2):
            i = u - 2 * q
            a11, a12, a21, a22 = H[i - 1, i - 1], H[i - 1, i], H[i, i - 1], H[i, i]
            hd = (a11 - a22) / 2
            disc = np.sqrt(hd * hd + a12 * a21)
            mid = (a11 + a22) / 2
            e1, e2 = mid + disc, mid - disc
            if cabs1(e2 - a22) < cabs1(e1 - a22):
                e1, e2 = e2, e1
            if its % 10 == 0:
                ex = 0.75 * cabs1(H[u, u - 1])
                e1, e2 = e1 + ex, e2 + 1j * ex
            sh += [e1, e2]
        chase_batch(H, l, u, sh[:nb], s, wsz)
        batches += 1
    return eig, batches, max(wsz)


if __name__ == '__main__':
    import scipy.linalg as la
    rng = np.random.default_rng(0)
    for n in (40, 97, 200):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        Hh = la.hessenberg(A)
        e, nbatch, wmax = qr_eigs(Hh, 8, 24)
        ref = np.linalg.eigvals(A)
        d = np.abs(e[:, None] - ref[None, :]).min(axis=1).max()
        print(n, 'batches', nbatch, 'max window', wmax, 'err', d / np.abs(ref).max())
