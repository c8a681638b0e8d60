"""Development prototype (NumPy): does aggressive early deflation (AED, LAPACK zlaqr2/zlaqr3)
shorten the multi-bulge QR cascade of csrc/eig_small.cu on the SAFE matrices?

Model of k_hqr_mb: a batch takes up to NB = 8 shifts from the trailing 2 x 2 diagonal blocks of the
active window [l, u] and runs them as NB successive implicit single-shift sweeps (the kernel chases
them concurrently, 3 positions apart, with the same result), which costs (u - l) + 3 (NB - 1)
macro steps; deflation tests (standard + Ahues-Tisseur) run between batches.  With AED the
trailing nw x nw window is Schur-decomposed before a batch, the spike is tested from the bottom,
converged eigenvalues are deflated (undeflatable ones are moved up, ztrexc), the rest is returned to
Hessenberg form, and the undeflated window eigenvalues are the next shifts.

Counts per matrix: batches, macro steps (split by the top index u = kernel stage), AED calls.
Not imported by the product or the tests.   usage: python dev/proto_aed.py [nw ...]
"""
import os
import sys

import numpy as np
import scipy.linalg as la

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'axisym-safe-python_b200'))
from dev.proto_eig import balance, hessenberg, cabs1, EPS, SMALL   # noqa: E402

NB = 8


def find_split(H, l0, u):
    """Lowest l >= l0 such that H[l, l-1] is negligible (zlahqr tests), scanning up from u."""
    n = H.shape[0]
    for k in range(u, l0, -1):
        sub = cabs1(H[k, k - 1])
        if sub <= SMALL:
            return k
        tst = cabs1(H[k - 1, k - 1]) + cabs1(H[k, k])
        if tst == 0:
            if k - 2 >= 0:
                tst += abs(H[k - 1, k - 2])
            if k + 1 <= n - 1:
                tst += abs(H[k + 1, k])
        if sub <= EPS * tst:
            ab = max(sub, cabs1(H[k - 1, k]))
            ba = min(sub, cabs1(H[k - 1, k]))
            aa = max(cabs1(H[k, k]), cabs1(H[k - 1, k - 1] - H[k, k]))
            bb = min(cabs1(H[k, k]), cabs1(H[k - 1, k - 1] - H[k, k]))
            s = aa + ab
            if ba * (ab / s) <= max(SMALL, EPS * (bb * (aa / s))):
                return k
    return l0


def sweep(H, l, u, sig):
    """One implicit-equivalent (explicit) single-shift QR sweep on the active block [l, u]."""
    idx = np.arange(l, u + 1)
    H[idx, idx] -= sig
    rot = []
    for k in range(l, u):
        f, g = H[k, k], H[k + 1, k]
        r2 = abs(f) ** 2 + abs(g) ** 2
        if r2 == 0:
            rot.append((1.0 + 0j, 0j))
            continue
        ir = 1 / np.sqrt(r2)
        a, b = np.conj(f) * ir, np.conj(g) * ir
        rot.append((a, b))
        top = a * H[k, k:u + 1] + b * H[k + 1, k:u + 1]
        bot = -np.conj(b) * H[k, k:u + 1] + np.conj(a) * H[k + 1, k:u + 1]
        H[k, k:u + 1], H[k + 1, k:u + 1] = top, bot
        H[k + 1, k] = 0
    for k in range(l, u):
        a, b = rot[k - l]
        c0 = H[l:k + 2, k] * np.conj(a) + H[l:k + 2, k + 1] * np.conj(b)
        c1 = -H[l:k + 2, k] * b + H[l:k + 2, k + 1] * a
        H[l:k + 2, k], H[l:k + 2, k + 1] = c0, c1
    H[idx, idx] += sig


def shifts_2x2(H, l, u, nb):
    """Eigenvalues of the trailing 2 x 2 diagonal blocks, per block the one nearer to its lower
    diagonal entry first (what k_hqr_mb takes, eig_small.cu 'shifts')."""
    out = []
    k = u
    while len(out) < nb and k - 1 >= l:
        e = list(np.linalg.eigvals(H[k - 1:k + 1, k - 1:k + 1]))
        e.sort(key=lambda z: cabs1(z - H[k, k]))
        out.extend(e)
        k -= 2
    if len(out) < nb and k >= l:
        out.append(H[k, k])
    return out[:nb]


def lartg(f, g):
    """[c s; -conj(s) c] [f; g] = [r; 0], c real."""
    if g == 0:
        return 1.0, 0j
    if f == 0:
        return 0.0, np.conj(g) / abs(g)
    d = np.hypot(abs(f), abs(g))
    c = abs(f) / d
    s = (f / abs(f)) * np.conj(g) / d
    return c, s


def trexc_swap(T, Z, k):
    """Swap the adjacent diagonal entries k, k+1 of the upper-triangular T (ztrexc step)."""
    n = T.shape[0]
    t11, t22 = T[k, k], T[k + 1, k + 1]
    c, s = lartg(T[k, k + 1], t22 - t11)
    if k + 2 < n:
        x, y = T[k, k + 2:].copy(), T[k + 1, k + 2:].copy()
        T[k, k + 2:], T[k + 1, k + 2:] = c * x + s * y, c * y - np.conj(s) * x
    x, y = T[:k, k].copy(), T[:k, k + 1].copy()
    T[:k, k], T[:k, k + 1] = c * x + np.conj(s) * y, c * y - s * x
    T[k, k], T[k + 1, k + 1] = t22, t11
    x, y = Z[:, k].copy(), Z[:, k + 1].copy()
    Z[:, k], Z[:, k + 1] = c * x + np.conj(s) * y, c * y - s * x


def aed(H, l, u, nw):
    """Aggressive early deflation on the trailing window of [l, u].  Returns (new u, shifts)."""
    nw = min(nw, u - l + 1)
    kw = u - nw + 1                                  # window = rows / columns kw .. u
    spike = H[kw, kw - 1] if kw > l else 0.0
    T, Z = la.schur(H[kw:u + 1, kw:u + 1], output='complex')
    if kw == l:                                      # the window is the whole active block: all converge
        ev = np.diag(T).copy()
        H[kw:u + 1, kw:u + 1] = T
        return kw - 1, list(ev), ev
    ns = nw                                          # T[:ns, :ns] not (yet) deflated
    ilst = 0
    for _ in range(nw):
        foo = cabs1(T[ns - 1, ns - 1])
        if foo == 0:
            foo = cabs1(spike)
        if cabs1(spike) * cabs1(Z[0, ns - 1]) <= max(SMALL, EPS * foo):
            ns -= 1                                  # deflatable
        else:
            for k in range(ns - 2, ilst - 1, -1):    # move it to the top of the undeflated part
                trexc_swap(T, Z, k)
            ilst += 1
        if ilst >= ns:
            break
    deflated = np.diag(T)[ns:].copy()
    if ns == nw:                                     # nothing deflated: leave H alone
        return u, list(np.diag(T)), deflated
    # window <- T with the spike in column kw-1; the undeflated part back to Hessenberg form
    s = spike * np.conj(Z[0, :])                     # Z^H (spike e_1)
    Q = np.eye(nw, dtype=complex)
    if ns > 1:
        v = s[:ns].copy()
        alpha = v[0]
        xn = np.linalg.norm(v[1:])
        if xn != 0:
            beta = -np.copysign(np.hypot(abs(alpha), xn), alpha.real if alpha.real != 0 else 1.0)
            tau = (beta - alpha) / beta
            v = v / (alpha - beta)
            v[0] = 1.0
            P = np.eye(ns, dtype=complex) - tau * np.outer(v, v.conj())              # P^H s = beta e_1
            Hs, Q2 = la.hessenberg(P.conj().T @ T[:ns, :ns] @ P, calc_q=True)
            Qs = P @ Q2
            Q[:ns, :ns] = Qs
            T[:ns, ns:] = Qs.conj().T @ T[:ns, ns:]
            T[:ns, :ns] = np.triu(Hs, -1)
            s = np.concatenate([Qs.conj().T @ s[:ns], np.zeros(nw - ns)])
    s[ns:] = 0
    ZZ = Z @ Q
    H[l:kw, kw:u + 1] = H[l:kw, kw:u + 1] @ ZZ       # rows of the active block above the window
    H[kw:u + 1, kw:u + 1] = T
    H[kw:u + 1, kw - 1] = 0
    if ns > 0:
        H[kw, kw - 1] = s[0]
    shifts = list(np.linalg.eigvals(H[kw:kw + ns, kw:kw + ns])) if ns > 0 else []
    return kw + ns - 1, shifts, deflated


def run(H, nw=0, nb=NB, stages=(117, 95, 82), aed_above=0):
    H = H.copy()
    n = H.shape[0]
    eig = []
    st = dict(batches=0, macro=0, aed=0, aed_defl=0, by_stage=[0, 0, 0, 0], fail=0)
    u = n - 1
    stall = 0
    while u >= 0:
        l = find_split(H, 0, u)
        if l > 0:
            H[l, l - 1] = 0
        if l == u:
            eig.append(H[u, u])
            u -= 1
            stall = 0
            continue
        shifts = None
        if nw and u >= aed_above:                        # aed_above = 117: AED in the one-CTA-per-SM stage only
            st['aed'] += 1
            u2, shifts, defl = aed(H, l, u, nw)
            eig.extend(defl)
            st['aed_defl'] += len(defl)
            if u2 < u:
                u = u2
                stall = 0
                if u < l + 1 or len(defl) * 100 >= 14 * min(nw, u - l + 1 + len(defl)):   # nibble: AED again
                    continue
        w = u - l + 1
        nbw = min(nb, (w - 2) // 3 + 1)                  # the kernel's bulge count for this window
        if not shifts:
            shifts = shifts_2x2(H, l, u, nbw)
        else:
            shifts = sorted(shifts, key=lambda z: abs(z - H[u, u]))[:nbw]
        stall += 1
        if stall % 10 == 0:
            shifts = [0.75 * abs(H[u, u - 1].real) + H[u, u]] * len(shifts)
        if stall > 80:
            st['fail'] += 1
            break
        for sg in shifts:
            sweep(H, l, u, sg)
        steps = (w - 1) + 3 * (len(shifts) - 1)
        st['batches'] += 1
        st['macro'] += steps
        st['by_stage'][sum(u < s for s in stages)] += steps      # the kernel's stages end at u < m_stop
    return np.array(eig), st


def main():
    from oracle import safe_oracle as so
    from axisafe_b200 import cases
    nws = [int(a) for a in sys.argv[1:]] or [0, 8, 12, 16, 24]
    c = cases.pipe_soil_pml(nf=2000)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    pick = [3, 400, 900, 1400, 1999]
    mats = []
    for i in pick:
        S = so.S_matrix(op, 2 * np.pi * c['f'][i])
        Sb, d, _ = balance(S)
        mats.append((hessenberg(Sb), np.linalg.eigvals(S)))
    print('cfg1, 2N = %d, %d matrices; stages by top index u: >=117 | 95-116 | 82-94 | <82' % (mats[0][0].shape[0], len(mats)))
    for nw in nws:
        tot = dict(batches=0, macro=0, aed=0, aed_defl=0, by_stage=[0, 0, 0, 0], fail=0)
        worst = 0.0
        for H, ref in mats:
            ev, st = run(H, nw)
            for k_ in ('batches', 'macro', 'aed', 'aed_defl', 'fail'):
                tot[k_] += st[k_]
            tot['by_stage'] = [a + b for a, b in zip(tot['by_stage'], st['by_stage'])]
            if len(ev) == len(ref):
                worst = max(worst, so.match_eigenvalues(ref, ev)[1])
            else:
                worst = np.inf
        m = len(mats)
        print('nw = %2d: batches %5.1f  macro steps %6.0f  by stage %s  AED calls %5.1f (deflated %5.1f)  '
              'max rel err vs eigvals %.1e  fails %d' % (nw, tot['batches'] / m, tot['macro'] / m,
                                                        [int(x / m) for x in tot['by_stage']], tot['aed'] / m,
                                                        tot['aed_defl'] / m, worst, tot['fail']))


if __name__ == '__main__':
    main()
