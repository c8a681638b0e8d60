"""Development prototype (NumPy): can the eigenpairs of frequency i-1 be continued to frequency i
by a few inverse-iteration steps with shift updates, instead of a fresh QR solve?  Counts, on the cfg1
sweep at the north-star resolution, how many of the 2N wavenumbers are recovered (bijectively, 1e-8)
and how often two continued modes collapse onto the same eigenvalue.  Not imported by the product."""
import os
import sys

import numpy as np
import scipy.linalg as la

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'axisym-safe-python_b200'))
from oracle import safe_oracle as so          # noqa: E402
from axisafe_b200 import cases                 # noqa: E402


def continue_modes(S, k0, X0, iters=3):
    """Inverse iteration with Rayleigh-type shift updates, all modes independently."""
    n = S.shape[0]
    k, X = k0.copy(), X0.copy()
    for j in range(n):
        mu, x = k[j], X[:, j] / np.linalg.norm(X[:, j])
        for _ in range(iters):
            try:
                y = la.solve(S - mu * np.eye(n), x)
            except la.LinAlgError:
                break
            # shift update from the generalised Rayleigh quotient x^H y: mu_new = mu + 1 / (x^H y)
            d = np.vdot(x, y)
            if d != 0:
                mu = mu + 1.0 / d
            x = y / np.linalg.norm(y)
        k[j], X[:, j] = mu, x
    return k, X


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    c = cases.pipe_soil_pml(nf=nf)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    w = 2 * np.pi * c['f']
    for start in (5, nf // 4, nf // 2, nf - 12):
        k, X = la.eig(so.S_matrix(op, w[start]))
        lost_total = 0
        for i in range(start + 1, start + 9):
            S = so.S_matrix(op, w[i])
            ref = la.eigvals(S)
            k, X = continue_modes(S, k, X)
            # bijective match: how many reference eigenvalues have their own continued mode
            d = abs(ref[:, None] - k[None, :]) / np.maximum(abs(ref[:, None]), 1e-300)
            hit = (d.min(axis=1) <= 1e-8)
            owners = d.argmin(axis=1)[hit]
            lost = len(ref) - len(set(owners))
            lost_total += lost
            print('start %4d step %d: %3d of %d eigenvalues recovered, %d lost (collapsed or drifted)'
                  % (start, i - start, len(set(owners)), len(ref), lost))
        print()


def grid(nf=2000, starts=(2, 40, 300, 700, 1100, 1500, 1900)):
    """Iterations x jump size: eigenpairs of frequency s continued DIRECTLY to frequency s + jump.
    A case fails when fewer than 2N eigenvalues are recovered one-to-one at 1e-8.  Run-time check a
    kernel could afford: the power sums sum(k) = tr S and sum(k^2) = tr S^2, normalised by sum |k| and
    sum |k|^2 (the SAFE spectrum is +-k symmetric, tr S ~ 0, so a check relative to |tr S| is useless)."""
    import warnings
    warnings.filterwarnings('ignore')
    c = cases.pipe_soil_pml(nf=nf)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    w = 2 * np.pi * c['f']
    for iters in (1, 2, 3, 4):
        for jump in (1, 2, 4, 8, 16, 32):
            worst, fails, caught = 10 ** 9, 0, 0
            for s0 in starts:
                k, X = la.eig(so.S_matrix(op, w[s0]))
                S = so.S_matrix(op, w[s0 + jump])
                ref = la.eigvals(S)
                k2, _ = continue_modes(S, k, X, iters)
                d = abs(ref[:, None] - k2[None, :]) / np.maximum(abs(ref[:, None]), 1e-300)
                hit = d.min(axis=1) <= 1e-8
                rec = len(set(d.argmin(axis=1)[hit]))
                worst = min(worst, rec)
                if rec < len(ref):
                    fails += 1
                    chk = max(abs(k2.sum() - np.trace(S)) / abs(k2).sum(),
                              abs((k2 ** 2).sum() - np.trace(S @ S)) / (abs(k2) ** 2).sum())
                    caught += chk > 1e-10
            print('iters %d jump %2d: worst recovered %3d/%d, failing cases %d/%d, flagged by the power sums (> 1e-10) %d'
                  % (iters, jump, worst, len(ref), fails, len(starts), caught))


if __name__ == '__main__':
    if 'grid' in sys.argv[1:]:
        grid()
    else:
        main()
