"""Development prototype (NumPy) of the GPU eigenvalue algorithm: LAPACK-style
scaling-only balancing, Householder Hessenberg, explicit single-shift QR with
Wilkinson shifts on the active window.  Sequential semantics identical to
csrc/eig_small.cu; used to validate accuracy/sweep counts before writing CUDA.
Not imported by the product or the tests."""
import numpy as np
import scipy.linalg as la

EPS = np.finfo(float).eps
SMALL = np.finfo(float).tiny / EPS


def cabs1(z):
    return abs(z.real) + abs(z.imag)


def balance(A, maxit=50):
    """zgebal('S')-like: Gauss-Seidel power-of-2 scaling using 2-norms incl. diagonal."""
    A = A.copy()
    n = A.shape[0]
    d = np.ones(n)
    sweeps = 0
    for _ in range(maxit):
        noconv = False
        sweeps += 1
        for i in range(n):
            c = np.linalg.norm(A[:, i])
            r = np.linalg.norm(A[i, :])
            if c == 0 or r == 0:
                continue
            g = r / 2
            f = 1.0
            s = c + r
            while c < g:
                f *= 2; c *= 2; g /= 2; r /= 2
            g = c / 2
            while g >= r:
                f /= 2; c /= 2; g /= 2; r *= 2
            if c + r >= 0.95 * s:
                continue
            noconv = True
            d[i] *= f
            A[i, :] /= f
            A[:, i] *= f
        if not noconv:
            break
    return A, d, sweeps


def hessenberg(A):
    """Unblocked Householder reduction (zgehd2 semantics)."""
    H = A.copy()
    n = H.shape[0]
    for k in range(n - 2):
        x = H[k + 1:, k].copy()
        alpha = x[0]
        xnorm = np.linalg.norm(x[1:])
        if xnorm == 0 and alpha.imag == 0:
            continue
        beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xnorm ** 2), alpha.real)
        tau = (beta - alpha) / beta
        v = x / (alpha - beta)
        v[0] = 1.0
        # H = (I - tau^H v v^H) H (I - tau v v^H)
        H[k + 1:, k] = 0
        H[k + 1, k] = beta
        w = H[:, k + 1:] @ v
        H[:, k + 1:] -= tau * np.outer(w, v.conj())
        y = v.conj() @ H[k + 1:, k + 1:]
        H[k + 1:, k + 1:] -= np.conj(tau) * np.outer(v, y)
    return np.triu(H, -1)


def wilkinson(H, i):
    t = H[i, i]
    u2 = np.sqrt(H[i - 1, i] + 0j) * np.sqrt(H[i, i - 1] + 0j)
    s = cabs1(u2)
    if s != 0:
        x = 0.5 * (H[i - 1, i - 1] - t)
        sx = cabs1(x)
        s = max(s, sx)
        y = s * np.sqrt((x / s) ** 2 + (u2 / s) ** 2)
        if sx > 0:
            xs = x / sx
            if xs.real * y.real + xs.imag * y.imag < 0:
                y = -y
        t = t - u2 * (u2 / (x + y))
    return t


def qr_eigs(H, stats=None):
    H = H.copy()
    n = H.shape[0]
    eig = np.zeros(n, complex)
    u = n - 1
    its = 0
    total = 0
    rots = 0
    while u >= 0:
        l = 0
        for k in range(u, 0, -1):
            sub = cabs1(H[k, k - 1])
            if sub <= SMALL:
                l = k; break
            tst = cabs1(H[k - 1, k - 1]) + cabs1(H[k, k])
            if tst == 0:
                if k - 2 >= 0: tst += abs(H[k - 1, k - 2])
                if k + 1 <= n - 1: tst += abs(H[k + 1, k])
            if sub <= EPS * tst:
                ab = max(sub, cabs1(H[k - 1, k]))
                ba = min(sub, cabs1(H[k - 1, k]))
                aa = max(cabs1(H[k, k]), cabs1(H[k - 1, k - 1] - H[k, k]))
                bb = min(cabs1(H[k, k]), cabs1(H[k - 1, k - 1] - H[k, k]))
                s = aa + ab
                if ba * (ab / s) <= max(SMALL, EPS * (bb * (aa / s))):
                    l = k; break
        if l > 0:
            H[l, l - 1] = 0
        if l == u:
            eig[u] = H[u, u]; u -= 1; its = 0
            continue
        its += 1; total += 1
        if its > 60:
            raise RuntimeError('no convergence')
        if its % 20 == 10:
            sig = 0.75 * abs(H[l + 1, l].real) + H[l, l]
        elif its % 20 == 0:
            sig = 0.75 * abs(H[u, u - 1].real) + H[u, u]
        else:
            sig = wilkinson(H, u)
        idx = np.arange(l, u + 1)
        H[idx, idx] -= sig
        rot = []
        for k in range(l, u):
            f, g = H[k, k], H[k + 1, k]
            r2 = abs(f) ** 2 + abs(g) ** 2
            if r2 == 0:
                rot.append((1.0 + 0j, 0j)); continue
            ir = 1 / np.sqrt(r2)
            a, b = np.conj(f) * ir, np.conj(g) * ir       # G = [[a, b], [-conj(b), conj(a)]]
            rot.append((a, b))
            top = a * H[k, k:u + 1] + b * H[k + 1, k:u + 1]
            bot = -np.conj(b) * H[k, k:u + 1] + np.conj(a) * H[k + 1, k:u + 1]
            H[k, k:u + 1], H[k + 1, k:u + 1] = top, bot
            H[k + 1, k] = 0
            rots += 1
        for k in range(l, u):
            a, b = rot[k - l]
            # right-multiply by G^H = [[conj(a), -b], [conj(b), a]]
            c0 = H[l:k + 2, k] * np.conj(a) + H[l:k + 2, k + 1] * np.conj(b)
            c1 = -H[l:k + 2, k] * b + H[l:k + 2, k + 1] * a
            H[l:k + 2, k], H[l:k + 2, k + 1] = c0, c1
        H[idx, idx] += sig
    if stats is not None:
        stats['sweeps'] = total; stats['rots'] = rots
    return eig


def eigvals(S):
    Sb, d, sw = balance(S)
    H = hessenberg(Sb)
    st = {}
    ev = qr_eigs(H, st)
    st['bal_sweeps'] = sw
    return ev, st, (Sb, d, H)


def hessenberg_with_q(A):
    """As hessenberg() but also returns the reflectors (V columns, tau)."""
    H = A.copy()
    n = H.shape[0]
    V = np.zeros([n, n], complex)
    tau = np.zeros(n, complex)
    for k in range(n - 2):
        x = H[k + 1:, k].copy()
        alpha = x[0]
        xnorm = np.linalg.norm(x[1:])
        if xnorm == 0 and alpha.imag == 0:
            continue
        beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xnorm ** 2), alpha.real)
        tau[k] = (beta - alpha) / beta
        v = x / (alpha - beta)
        v[0] = 1.0
        V[k + 1:, k] = v
        H[k + 1:, k] = 0
        H[k + 1, k] = beta
        w = H[:, k + 1:] @ v
        H[:, k + 1:] -= tau[k] * np.outer(w, v.conj())
        y = v.conj() @ H[k + 1:, k + 1:]
        H[k + 1:, k + 1:] -= np.conj(tau[k]) * np.outer(v, y)
    return np.triu(H, -1), V, tau


def hess_inverse_iteration(H, lam, iters=2):
    """Eigenvectors of upper-Hessenberg H for all eigenvalues lam at once.

    One bottom-up pass per iteration: column rotations G_j (cols j, j+1) reduce
    H - lam I to upper triangular R row by row while the back-substitution R z = b is
    carried along; storage per eigenvalue is O(n) (rotations + z).  x = W z."""
    n = H.shape[0]
    ne = len(lam)
    hn = np.abs(H).sum(axis=0).max()
    tiny = EPS * hn
    b = np.ones([n, ne], complex)
    for it in range(iters):
        g11 = np.zeros([n, ne], complex); g12 = np.zeros([n, ne], complex)
        g21 = np.zeros([n, ne], complex); g22 = np.zeros([n, ne], complex)
        z = np.zeros([n, ne], complex)
        for k in range(n - 1, -1, -1):
            carry = H[k, n - 1] - (lam if k == n - 1 else 0)
            carry = carry * np.ones(ne, complex)
            acc = b[k].copy()
            for j in range(n - 2, k - 1, -1):
                xj = H[k, j] - (lam if j == k else 0)
                rkj1 = xj * g12[j] + carry * g22[j]
                carry = xj * g11[j] + carry * g21[j]
                acc -= rkj1 * z[j + 1]
            if k > 0:
                x = H[k, k - 1]
                r = np.sqrt(abs(x) ** 2 + abs(carry) ** 2)
                r = np.where(r == 0, tiny, r)
                g12[k - 1] = np.conj(x) / r; g22[k - 1] = np.conj(carry) / r
                g11[k - 1] = carry / r; g21[k - 1] = -x / r
                rkk = r
            else:
                rkk = carry
            rkk = np.where(abs(rkk) < tiny, tiny, rkk)
            z[k] = acc / rkk
        x = z.copy()
        for j in range(n - 1):
            xj, xj1 = x[j].copy(), x[j + 1].copy()
            x[j] = g11[j] * xj + g12[j] * xj1
            x[j + 1] = g21[j] * xj + g22[j] * xj1
        x /= np.linalg.norm(x, axis=0)
        b = x
    return x


def back_transform(x, V, tau, d):
    n = x.shape[0]
    y = x.copy()
    for k in range(n - 3, -1, -1):
        v = V[k + 1:, k]
        y[k + 1:] -= tau[k] * np.outer(v, v.conj() @ y[k + 1:])
    return d[:, None] * y


def hess_null_vectors(H, lam):
    """Single bottom-up RQ pass per eigenvalue, no back-substitution: x = W e_0, which
    satisfies (H - lam I) x = r_00 e_0.  Returns x and |r_00| / ||H||."""
    n = H.shape[0]
    ne = len(lam)
    hn = np.abs(H).sum(axis=0).max()
    tiny = EPS * hn
    p = np.zeros([n, ne], complex); q = np.zeros([n, ne], complex)
    r00 = None
    for k in range(n - 1, -1, -1):
        carry = (H[k, n - 1] - (lam if k == n - 1 else 0)) * np.ones(ne, complex)
        for j in range(n - 2, k - 1, -1):
            xj = H[k, j] - (lam if j == k else 0)
            carry = xj * p[j] - carry * np.conj(q[j])
        if k > 0:
            x = H[k, k - 1]
            r = np.sqrt(abs(x) ** 2 + abs(carry) ** 2)
            r = np.where(r == 0, tiny, r)
            p[k - 1] = carry / r; q[k - 1] = np.conj(x) / r
        else:
            r00 = carry
    x = np.zeros([n, ne], complex)
    prod = np.ones(ne, complex)
    for j in range(n - 1):
        x[j] = p[j] * prod
        prod = prod * (-np.conj(q[j]))
    x[n - 1] = prod
    return x, abs(r00) / hn


def hess_inverse_iteration_lu(H, lam, iters=1):
    """Like hess_inverse_iteration but with pivoted Gauss column operations (zlaein-style
    LU of a Hessenberg matrix, processed bottom-up): 16 B multiplier + 1 bit per step."""
    n = H.shape[0]
    ne = len(lam)
    hn = np.abs(H).sum(axis=0).max()
    tiny = EPS * hn
    b = np.ones([n, ne], complex)
    c1 = lambda v: abs(v.real) + abs(v.imag)
    for it in range(iters):
        mul = np.zeros([n, ne], complex); swp = np.zeros([n, ne], bool)
        z = np.zeros([n, ne], complex)
        for k in range(n - 1, -1, -1):
            carry = (H[k, n - 1] - (lam if k == n - 1 else 0)) * np.ones(ne, complex)
            acc = b[k].copy()
            for j in range(n - 2, k - 1, -1):
                a = (H[k, j] - (lam if j == k else 0)) * np.ones(ne, complex)
                a2 = np.where(swp[j], carry, a); c2 = np.where(swp[j], a, carry)
                a2 = a2 - mul[j] * c2
                acc -= c2 * z[j + 1]
                carry = a2
            if k > 0:
                a = H[k, k - 1] * np.ones(ne, complex)
                s = c1(a) > c1(carry)
                a2 = np.where(s, carry, a); c2 = np.where(s, a, carry)
                c2 = np.where(c1(c2) < tiny, tiny, c2)
                mul[k - 1] = a2 / c2; swp[k - 1] = s
                rkk = c2
            else:
                rkk = np.where(c1(carry) < tiny, tiny, carry)
            z[k] = acc / rkk
        x = z.copy()
        for j in range(n - 1):
            x[j + 1] = x[j + 1] - mul[j] * x[j]
            xj, xj1 = x[j].copy(), x[j + 1].copy()
            x[j] = np.where(swp[j], xj1, xj); x[j + 1] = np.where(swp[j], xj, xj1)
        x /= np.linalg.norm(x, axis=0)
        b = x
    return x
