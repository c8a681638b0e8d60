"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz from the UNMODIFIED reference.

Run inside the build container (needs /root/reference):
    OPENBLAS_NUM_THREADS=1 python oracle/make_golden.py
Every fixture holds, for one case of ``axisafe_b200.cases``:
  * mesh book-keeping (node radii, IEN, LM, Tdiag, element types, dof domains),
  * the reference's global matrices K1..K6, M, KF, KFt, KFz, K1_coupling (COO triplets),
  * per-element matrices of every element (K_1..K_6, K_F, K_Ft, K_Fz, M, H vectors),
  * S(w) at three frequencies rebuilt exactly as solver.py:115-134,
  * untracked scipy.linalg.eig eigenvalues of those S,
  * the reference's tracked ``k`` for the whole sweep and ``psi_hat`` (lower halves) at the
    same three frequencies,
  * energy_ratio / k_propagating outputs.
The case inputs are taken from axisafe_b200.cases; the script asserts that the reference's own
suggest_PML_parameters gives the same layer definitions.
"""
import os
import sys

import numpy as np
import scipy.linalg as la
import scipy.sparse as sps
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'axisym-safe-python_b200'))

from oracle import refshim  # noqa: E402

CASES = [('pipe_soil_pml', 0, {}), ('free_steel_pipe', 0, {'nf': 60}), ('free_steel_pipe', 1, {'nf': 60}),
         ('aristegui', 0, {'nf': 40}), ('muggleton_submerged', 0, {'nf': 40}),
         ('muggleton_buried', 0, {'nf': 25}), ('gravenkamp', 0, {'nf': 30}), ('gravenkamp', 1, {'nf': 30}),
         ('gravenkamp', 2, {'nf': 30}), ('nguyen', 0, {'nf': 30}), ('nguyen', 1, {'nf': 30}),
         ('pipe_soil_pml', 1, {'nf': 20})]
KPROP = {'pipe_soil_pml': (15, 0.935), 'aristegui': (100, 0.9), 'muggleton_submerged': (15, 0.8),
         'muggleton_buried': (15, 0.935), 'gravenkamp': (350, 0.97), 'nguyen': (200, 0.9)}


def coo(m):
    m = m.tocoo()
    return np.array(m.row, np.int32), np.array(m.col, np.int32), np.array(m.data)


def reference_S(solver, mesh, n, w):
    """S(w) exactly as the reference builds it (solver.py:115-134)."""
    B = sps.bmat([[None, mesh.K6.tocsr()], [mesh.K6.tocsr(), n * mesh.K4.tocsr() + solver.K3_hat.tocsr()]])
    Binv = spla.inv(B.tocsc())
    low = mesh.K1.tocsr() + n ** 2 * mesh.K5.tocsr() + n * solver.K2_hat.tocsr() - w ** 2 * mesh.M.tocsr()
    if mesh.is_coupled:
        low = low + 1j * w * mesh.K1_coupling.tocsr()
    A = sps.bmat([[mesh.K6.tocsr(), None], [None, -low]])
    return Binv.dot(A.tocsc()).toarray()


def main():
    ref = refshim.load()
    from axisafe_b200 import cases
    outdir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    for name, n, kw in CASES:
        c = cases.ALL[name](n=n, **kw)
        with refshim.quiet():
            mesh = ref.mesh.Mesh()
            mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
            mesh.assemble_matrices(n)
            sol = ref.solver.WaveElementAxisym(mesh)
            sol.solve(c['f'])
            er = kready = pidx = None
            if c['PML'] != 0:
                sol.energy_ratio()
                sol.k_propagating(*KPROP[name])
                er, kready, pidx = sol.er, sol.k_ready, sol.propagating_indices
        N = mesh.no_of_dofs
        pick = sorted({0, len(c['f']) // 3, len(c['f']) - 1})
        S = np.array([reference_S(sol, mesh, n, sol.w[i]) for i in pick])
        with refshim.quiet():
            ev = np.array([la.eig(s)[0] for s in S])
        data = dict(f=c['f'], n=n, N=N, pick=np.array(pick), S=S, eig_untracked=ev, k=sol.k,
                    psi_low=np.array([sol.psi_hat[i][N:] for i in pick]),
                    Tdiag=np.array(mesh.Tdiag), node_r=np.array([mesh.glob_nodes[k][0] for k in sorted(mesh.glob_nodes)]),
                    solid_dofs=np.array(mesh.dof_domains['solid'], np.int32),
                    fluid_dofs=np.array(mesh.dof_domains['fluid'], np.int32),
                    etypes=np.array([mesh.element_types[e] for e in sorted(mesh.IEN)]),
                    is_coupled=bool(mesh.is_coupled))
        for e in sorted(mesh.IEN):
            data['IEN_%d' % e] = np.array(mesh.IEN[e], np.int32)
            data['LM_%d' % e] = np.array(mesh.LM[e], np.int32)
            el = mesh.elements[e]
            for attr in ('K_1', 'K_2', 'K_3', 'K_4', 'K_5', 'K_6', 'K_F', 'K_Ft', 'K_Fz', 'M',
                         'Hs0', 'Hs1', 'Hf0', 'Hf1'):
                if getattr(el, attr, None) is not None and (len(mesh.IEN) <= 4):
                    data['el%d_%s' % (e, attr)] = np.asarray(getattr(el, attr))
        for g in ('K1', 'K2', 'K3', 'K4', 'K5', 'K6', 'M', 'KF', 'KFt', 'KFz', 'K1_coupling'):
            m = getattr(mesh, g)
            if m is not None:
                r, cc, d = coo(m.tocsr())     # duplicates summed, like every consumer does
                data['G_%s_r' % g], data['G_%s_c' % g], data['G_%s_d' % g] = r, cc, d
        if er is not None:
            data['er'], data['k_ready'], data['prop_idx'] = er, kready, np.array(pidx)
            data['kprop_args'] = np.array(KPROP[name])
        path = os.path.join(outdir, '%s_n%d.npz' % (name, n))
        np.savez_compressed(path, **data)
        print('%-28s N=%3d nf=%3d  %6.0f KB' % (os.path.basename(path), N, len(c['f']), os.path.getsize(path) / 1024))
        # the case builder must reproduce the reference's own PML suggestion
    with refshim.quiet():
        mine = __import__('axisafe_b200.mesh', fromlist=['x']).suggest_PML_parameters
        lam = ref.misc.young2lame(156e9, 0.2)
        args = (0.11, 3, list(lam) + [7800, 0.0], list(ref.misc.cLcS2lame(200, 100, 1500)) + [1500, 0.], 1e3, 20e3)
        a, b = ref.mesh.suggest_PML_parameters(*args, att=1), mine(*args, att=1)
    assert all(x == y for x, y in zip(a, b)), (a, b)
    np.savez(os.path.join(outdir, 'suggest_pml.npz'), args_scalar=np.array([0.11, 3, 1e3, 20e3]), out=np.array(a))
    print('suggest_PML_parameters: identical', a)


ENVEL_CASES = [('aristegui', 0, {'nf': 40}), ('pipe_soil_pml', 0, {'nf': 30}), ('gravenkamp', 1, {'nf': 30})]


def energy_velocity_fixtures():
    """tests/golden/envel_*.npz: the reference's energy_velocity (solver.py:301-409) for all modes.

    ``only_propagating=False`` because the reference's own ``propagating_indices == []`` test
    (solver.py:386) raises on current NumPy; the propagating subset is a column selection of these
    arrays.  Stored: tracked k, e_k, e_p, p, en_vel ([nf, 2N]) and the k_propagating arguments.
    """
    ref = refshim.load()
    from axisafe_b200 import cases
    outdir = os.path.join(ROOT, 'tests', 'golden')
    for name, n, kw in ENVEL_CASES:
        c = cases.ALL[name](n=n, **kw)
        with refshim.quiet():
            mesh = ref.mesh.Mesh()
            mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
            mesh.assemble_matrices(n)
            sol = ref.solver.WaveElementAxisym(mesh)
            sol.solve(c['f'])
            sol.energy_ratio()
            sol.k_propagating(*KPROP[name])
            sol.energy_velocity(only_propagating=False)
        # forced response (solver.py:411-474): unit force on the first solid DOF
        f_ext = np.zeros(mesh.no_of_dofs)
        f_ext[mesh.dof_domains['solid'][0]] = 1.0
        with refshim.quiet():
            frf = sol.calculate_frf(0.01, 0.05, f_ext, distance=0.3)
        path = os.path.join(outdir, 'envel_%s_n%d.npz' % (name, n))
        np.savez_compressed(path, f=c['f'], k=sol.k, e_k=sol.e_k, e_p=sol.e_p, p=sol.p, en_vel=sol.en_vel,
                            prop_idx=np.array(sol.propagating_indices), kprop_args=np.array(KPROP[name]),
                            frf=frf, frf_dof=int(mesh.dof_domains['solid'][0]))
        print('%-32s nf=%3d 2N=%3d %6.0f KB' % (os.path.basename(path), len(c['f']), sol.k.shape[1],
                                               os.path.getsize(path) / 1024))


SUBSET_CASES = [('free_steel_pipe', 0, {'nf': 24}, 3, 3000.0), ('pipe_soil_pml', 0, {'nf': 16}, 8, 1200.0),
                ('gravenkamp', 1, {'nf': 16}, 10, 2500.0)]


def subset_fixtures():
    """tests/golden/subset_*.npz: the reference's shift-invert mode (solver.py:137-140),
    ``solve(f, no_of_waves=m, central_wavespeed=c)``: tracked k [nf, m] and the lower halves of
    psi_hat [nf, N, m].  ARPACK starts from a random vector, so only the per-frequency SET of
    wavenumbers (and each mode shape up to sign) is reproducible; the column order at the first
    frequency is ARPACK's.  Also stored: the margin between the m-th and (m+1)-th nearest
    eigenvalue of the dense spectrum (a selection is only well defined when that is not tiny).
    """
    ref = refshim.load()
    from axisafe_b200 import cases
    outdir = os.path.join(ROOT, 'tests', 'golden')
    for name, n, kw, m, c0 in SUBSET_CASES:
        c = cases.ALL[name](n=n, **kw)
        with refshim.quiet():
            mesh = ref.mesh.Mesh()
            mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
            mesh.assemble_matrices(n)
            sol = ref.solver.WaveElementAxisym(mesh)
            sol.solve(c['f'], no_of_waves=m, central_wavespeed=c0)
            margin = []
            for i in range(len(c['f'])):
                ev = la.eig(reference_S(sol, mesh, n, sol.w[i]))[0]
                d = np.sort(abs(ev - sol.w[i] / c0))
                margin.append((d[m] - d[m - 1]) / d[m])
        N = mesh.no_of_dofs
        path = os.path.join(outdir, 'subset_%s_n%d.npz' % (name, n))
        np.savez_compressed(path, f=c['f'], k=sol.k, psi_low=sol.psi_hat[:, N:, :], m=m, c0=c0,
                            margin=np.array(margin))
        print('%-32s nf=%3d 2N=%3d m=%2d min margin %.2e %6.0f KB' % (
            os.path.basename(path), len(c['f']), 2 * N, m, min(margin), os.path.getsize(path) / 1024))


BIG_CASES = [('coated_pipe', 1, {}, [20e3, 35e3], 'cfg4'), ('ten_layer', 1, {}, [20e3], 'cfg5')]


def big_fixtures(which=None):
    """tests/golden/big_cfg4.npz / big_cfg5.npz: the full-size synthetic meshes of BASELINE.md
    (cfg4 2N = 3966, cfg5 2N = 7806), the reference's own S(w) (solver.py:115-134) pushed through the
    reference's own call ``scipy.linalg.eig`` (solver.py:136) at 2 / 1 frequencies.  Stored: the
    untracked eigenvalues, their condition numbers on the balanced matrix and ||S_bal||_2 (the inputs of
    oracle.eig_tolerance), tr S, tr S^2 and the B-normalised lower halves of 16 eigenvectors (every
    (2N / 16)-th column).  zgeev needs ~1-2 min (cfg4) / ~10 min (cfg5) per frequency on
    the build container's 8 threads; S itself (252 MB / 975 MB) is not stored - the GPU test
    rebuilds it from its own operators, which are pinned to the reference's by the small cases.
    """
    import time
    ref = refshim.load()
    from axisafe_b200 import cases
    outdir = os.path.join(ROOT, 'tests', 'golden')
    for name, n, kw, freqs, tag in BIG_CASES:
        if which and tag not in which:
            continue
        c = cases.ALL[name](n=n, nf=2, **kw)
        with refshim.quiet():
            mesh = ref.mesh.Mesh()
            mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
            mesh.assemble_matrices(n)
            sol = ref.solver.WaveElementAxisym(mesh)
            sol.t_transform()
        N = mesh.no_of_dofs
        B = sps.bmat([[None, mesh.K6.tocsr()], [mesh.K6.tocsr(), n * mesh.K4.tocsr() + sol.K3_hat.tocsr()]]).tocsr()
        cols = np.arange(0, 2 * N, (2 * N) // 16)[:16]
        ev, tr1, tr2, psi_low, secs, kappa, nrm2 = [], [], [], [], [], [], []
        for f in freqs:
            S = reference_S(sol, mesh, n, 2 * np.pi * f)
            t0 = time.time()
            with refshim.quiet():
                val, vec = la.eig(S)
            secs.append(time.time() - t0)
            v = vec[:, cols]
            v = v / np.sqrt(np.einsum('ij,ij->j', v, B @ v))
            ev.append(val); tr1.append(np.trace(S)); tr2.append(np.sum(S * S.T)); psi_low.append(v[N:])
            print('%s f=%g: zgeev %.0f s' % (tag, f, secs[-1]), flush=True)
            # conditioning of every eigenvalue on the balanced matrix (what oracle.eig_tolerance uses):
            # kappa_k = ||y_k|| ||x_k|| / |y_k^H x_k| with the left vectors taken from inv(V)
            _, D = la.matrix_balance(S, permute=False, separate=True)
            d = D[0]
            vinv = la.inv(vec)
            kap = np.linalg.norm(vinv * d[None, :], axis=1) * np.linalg.norm(vec / d[:, None], axis=0)
            Sb = S * (d[None, :] / d[:, None])
            x = np.ones(2 * N, complex)
            for _ in range(60):                                   # ||S_bal||_2 by power iteration on S^H S
                x = Sb.conj().T @ (Sb @ x)
                x /= np.linalg.norm(x)
            kappa.append(kap); nrm2.append(np.linalg.norm(Sb @ x))
            print('%s f=%g: kappa max %.2e median %.2e, ||S_bal||_2 %.3e' % (tag, f, kap.max(), np.median(kap), nrm2[-1]), flush=True)
            del S, vec, vinv, Sb
        path = os.path.join(outdir, 'big_%s.npz' % tag)
        np.savez_compressed(path, f=np.array(freqs), n=n, N=N, eig_untracked=np.array(ev), trS=np.array(tr1),
                            trS2=np.array(tr2), cols=cols, psi_low=np.array(psi_low), zgeev_seconds=np.array(secs),
                            kappa=np.array(kappa), norm2_balanced=np.array(nrm2),
                            threads=len(os.sched_getaffinity(0)))
        print('%-20s 2N=%d %6.0f KB' % (os.path.basename(path), 2 * N, os.path.getsize(path) / 1024), flush=True)


if __name__ == '__main__':
    if 'big' in sys.argv[1:]:
        big_fixtures([a for a in sys.argv[1:] if a.startswith('cfg')])
    elif 'subset' in sys.argv[1:]:
        subset_fixtures()
    elif 'envel' in sys.argv[1:]:
        energy_velocity_fixtures()
    else:
        main()
        energy_velocity_fixtures()
        subset_fixtures()
