"""Development prototype (NumPy): the blocked compact-WY Hessenberg reduction exactly as csrc/eig_bigred.cu
stages it (panel column update, reflector, GEMV over the rows below the panel top, Y / T recurrences,
top part of Y by a GEMM, right and left trailing updates).  Checks H = Q^H A Q against scipy."""
import numpy as np, scipy.linalg as la

def larfg(alpha, x):
    xn2 = np.vdot(x, x).real
    if xn2 == 0 and alpha.imag == 0:
        return alpha, x * 0, 0.0
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
    tau = (beta - alpha.real) / beta - 1j * alpha.imag / beta
    return beta, x / (alpha - beta), tau

def blocked(A, nb):
    A = A.copy(); n = A.shape[0]; tau_all = np.zeros(n, complex)
    k0 = 0
    while k0 < n - 2:
        nbc = min(nb, n - 2 - k0)
        V = np.zeros([n, nbc], complex); Y = np.zeros([n, nbc], complex); T = np.zeros([nbc, nbc], complex)
        A0 = A.copy()                      # the matrix at panel start (the kernels never need a copy: see below)
        for j in range(nbc):
            c = k0 + j
            b = A[k0 + 1:, c].copy()
            if j > 0:
                b -= Y[k0 + 1:, :j] @ V[c, :j].conj()
                w = V[k0 + 1:, :j].conj().T @ b
                w = T[:j, :j].conj().T @ w
                b -= V[k0 + 1:, :j] @ w
            A[k0 + 1:, c] = b
            beta, v, tau = larfg(A[c + 1, c], A[c + 2:, c])
            A[c + 1, c] = beta; A[c + 2:, c] = v
            V[c + 1, j] = 1; V[c + 2:, j] = v
            tau_all[c] = tau
            # GEMV with the ORIGINAL columns right of c (panel columns right of c are still original in A,
            # trailing columns too), rows below the panel top only
            yv = A[k0 + 1:, c + 1:] @ V[c + 1:, j]
            assert np.allclose(yv, A0[k0 + 1:, c + 1:] @ V[c + 1:, j])
            z = V[:, :j].conj().T @ V[:, j]
            Y[k0 + 1:, j] = tau * (yv - Y[k0 + 1:, :j] @ z)
            T[:j, j] = -tau * (T[:j, :j] @ z); T[j, j] = tau
        # top part of Y: rows 0..k0 (panel columns right of the first are untouched there)
        Y[:k0 + 1, :] = (A[:k0 + 1, k0 + 1:] @ V[k0 + 1:, :]) @ T
        # top rows of the panel columns: right update
        for j in range(1, nbc):
            A[:k0 + 1, k0 + j] -= Y[:k0 + 1, :j] @ V[k0 + j, :j].conj()
        # trailing columns
        t0 = k0 + nbc
        A[:, t0:] -= Y @ V[t0:, :].conj().T
        W = V[k0 + 1:, :].conj().T @ A[k0 + 1:, t0:]
        W = T.conj().T @ W
        A[k0 + 1:, t0:] -= V[k0 + 1:, :] @ W
        k0 += nbc
    return A, tau_all

rng = np.random.default_rng(1)
for n, nb in [(37, 8), (64, 16), (150, 32)]:
    A = rng.standard_normal([n, n]) + 1j * rng.standard_normal([n, n])
    R, tau = blocked(A, nb)
    H = np.triu(R, -1)
    ev0, ev1 = np.sort_complex(np.linalg.eigvals(A)), np.sort_complex(np.linalg.eigvals(H))
    # rebuild Q from the stored reflectors
    Q = np.eye(n, dtype=complex)
    for c in range(n - 2):
        v = np.zeros(n, complex); v[c + 1] = 1; v[c + 2:] = R[c + 2:, c]
        Q = Q @ (np.eye(n) - tau[c] * np.outer(v, v.conj()))
    print(n, nb, 'eig err %.1e' % np.abs(ev0 - ev1).max(), 'Q^H A Q - H %.1e' % np.abs(Q.conj().T @ A @ Q - H).max(),
          'unitary %.1e' % np.abs(Q.conj().T @ Q - np.eye(n)).max())
