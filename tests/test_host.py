"""CPU tests: tables, host mesh logic, C-ABI surface, tracking scan (no GPU needed)."""
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_CASES, CASE_IDS, ROOT, load_golden, golden_dense, make_case


def test_quadrature_invariants():
    from axisafe_b200 import shape_functions as sf
    for m in (4, 5, 9, 11, 14, 23, 30):
        for build, lag in ((sf.build_GLL_quadrature, sf.lagrange_poly),
                           (sf.build_GLJ_quadrature, sf.lagrange_GLJ)):
            ksi, w = build(m)
            assert abs(w.sum() - 2.0) < 1e-13            # SURVEY.md section 4 invariant
            assert ksi[0] == -1.0 and abs(ksi[-1] - 1.0) < 1e-15 and np.all(np.diff(ksi) > 0)
            N, D = lag(ksi)
            assert np.array_equal(N, np.eye(m))
            assert np.abs(D.sum(axis=0)).max() < 1e-11   # derivative of the constant function
            assert np.abs(D.T @ ksi - 1).max() < 1e-11   # derivative of x
    # GLL integrates polynomials of degree 2m-3 exactly, GLJ integrates (1+x) x^k for k <= 2Q-4
    ksi, w = sf.build_GLL_quadrature(7)
    assert abs(w @ ksi ** 10 - 2 / 11) < 1e-14
    ksi, w = sf.build_GLJ_quadrature(7)
    assert abs(w @ ksi ** 10 - (2 / 11)) < 1e-14          # int (1+x) x^10 = 2/11


def test_tables_match_oracle():
    from axisafe_b200 import shape_functions as sf
    from oracle import safe_oracle as so
    for m in (5, 9, 14, 30):
        k1, w1 = sf.build_GLL_quadrature(m)
        k2, w2 = so.gll(m)
        assert np.abs(k1 - k2).max() < 1e-15 and np.abs(w1 - w2).max() < 1e-15
        assert np.abs(sf.lagrange_poly(k1)[1] - so.dmatrix(k2)).max() < 1e-12 * m * m
        k1, w1 = sf.build_GLJ_quadrature(m)
        k2, w2 = so.glj(m)
        assert np.abs(k1 - k2).max() < 1e-15 and np.abs(w1 - w2).max() < 1e-14


def test_misc_and_pml_suggestion():
    import axisafe_b200 as ax
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'suggest_pml.npz'))
    lam = ax.misc.young2lame(156e9, 0.2)
    soil = list(ax.misc.cLcS2lame(200, 100, 1500)) + [1500, 0.]
    got = ax.mesh.suggest_PML_parameters(0.11, 3, list(lam) + [7800, 0.0], soil, 1e3, 20e3, att=1)
    assert np.array_equal(np.array(got, float), g['out'])     # bit-identical to the reference
    assert ax.misc.G_nu2lame(2.0, 0.25) == (2.0, 2.0)
    E, nu = ax.misc.bulkvel2E(5960., 3260., 7932.)
    l1, m1 = ax.misc.young2lame(E, nu)
    l2, m2 = ax.misc.cLcS2lame(5960., 3260., 7932.)
    assert abs(l1 - l2) < 1e-6 * l2 and abs(m1 - m2) < 1e-6 * m2
    assert np.allclose(ax.misc.young2C(E, nu), ax.misc.lame2C(l1, m1))


@pytest.mark.parametrize('name,n,kw', GOLDEN_CASES, ids=CASE_IDS)
def test_mesh_numbering_matches_reference(name, n, kw):
    """Mesh.create + DOF numbering (mesh.py:153-288, :337-464) against the reference's own output."""
    import axisafe_b200 as ax
    g = load_golden(name, n)
    c = make_case(name, n, kw)
    m = ax.mesh.Mesh()
    m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    m.n = n
    m._instantiate_elements(n)
    m._number_dofs()
    assert m.no_of_dofs == int(g['N'])
    assert np.array_equal(np.array([m.glob_nodes[k][0] for k in sorted(m.glob_nodes)]), g['node_r'])
    for e in sorted(m.IEN):
        assert np.array_equal(np.array(m.IEN[e]), g['IEN_%d' % e])
        assert np.array_equal(np.array(m.LM[e]), g['LM_%d' % e])
    assert np.array_equal(np.array(m.Tdiag), g['Tdiag'])
    assert list(g['etypes']) == [m.element_types[e] for e in sorted(m.IEN)]
    assert np.array_equal(np.array(m.dof_domains['solid'], np.int32), g['solid_dofs'])
    assert np.array_equal(np.array(m.dof_domains['fluid'], np.int32), g['fluid_dofs'])
    assert m.is_coupled == bool(g['is_coupled'])


def test_lubricated_coupling_numbering():
    """Radial-only contact duplicates the interface node and shares u_r (mesh.py:219-225, :377-404)."""
    import axisafe_b200 as ax
    steel = list(ax.misc.cLcS2lame(5960, 3260, 7932)) + [7932, 0.0]
    sets = [['a', 0.1, 0.01, steel, 'SLAX6', 3], ['b', 0.11, 0.01, steel, 'SLAX6', 3]]
    m = ax.mesh.Mesh()
    m.create(sets, 1e4, lubricated_coupling=('a', 'b'))
    assert m.uncoupled_nodes == [[4, 5]] and m.IEN[2][0] == 5
    m.n = 0
    m._instantiate_elements(0)
    m._number_dofs()
    assert m.ID[5] == [m.ID[4][0], 12, 13] and m.no_of_dofs == 23
    assert m.Tdiag[12:14] == [1j, 1j]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import axisafe_b200 as ax
    c = make_case('free_steel_pipe', 0, {'nf': 4})
    m = ax.mesh.Mesh()
    m.create(c['sets'], c['fmax'])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.assemble_matrices(0)


def test_cabi_exports_every_declared_symbol(built_lib):
    """The shared library loads and exports exactly what include/axisafe_b200.h declares."""
    import ctypes
    from axisafe_b200 import _native
    header = open(os.path.join(ROOT, 'include', 'axisafe_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(axs_[a-z_0-9S]+)\s*\(', header)))
    assert declared, 'no declarations parsed'
    handle = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(handle, name), name
    assert sorted(_native.EXPORTED) == declared
    L = _native.lib()
    assert L.axs_version() >= 100
    assert L.axs_strerror(0) == b'success' and b'invalid argument 3' in L.axs_strerror(-3)
    assert L.axs_max_small_n2() >= 166
    assert L.axs_eig_worksize(63, 10) > 10 * 126 * 126 * 16
    # argument checking happens before any CUDA call
    assert L.axs_element_matrices(0, None, None, None, None, None, None) == -1
    assert L.axs_hqr(63, 4, None, None, None, 0, None) == -3
    assert L.axs_max_n2() >= 7806                    # cfg5 of BASELINE.json
    assert L.axs_eig_worksize(200, 4) > 4 * 400 * 400 * 32
    assert L.axs_reduce(5000, 1, None, None, None, None, None, None, 1, None, None, 0, None) == 100001


def _python_scan(arg, amax, n2):
    """Straight restatement of solver.py:159-175 on the arg-max tables."""
    nf = len(arg) + 1
    order = np.zeros([nf, n2], int)
    order[0] = np.arange(n2)
    for i in range(1, nf):
        prev = order[i - 1]
        new_order = arg[i - 1][prev].copy()
        valid = np.round(amax[i - 1][prev], 2) > 0.9
        valid_entries = np.delete(new_order, np.where(valid == False))  # noqa: E712
        if not valid.all():
            unassigned = list(set(range(n2)) - set(valid_entries))
            j = 0
            for ii in range(n2):
                if not valid[ii]:
                    new_order[ii] = unassigned[j]
                    j += 1
        order[i] = new_order
    return order


def test_track_scan_matches_python_semantics(built_lib):
    from axisafe_b200 import _native
    rng = np.random.default_rng(1)
    n2, nf = 30, 200
    arg = np.array([rng.permutation(n2) for _ in range(nf - 1)], dtype=np.int32)
    amax = 0.95 + 0.05 * rng.random([nf - 1, n2])
    for i in rng.choice(nf - 1, 12, replace=False):          # some weak / duplicated matches
        rows = rng.choice(n2, 3, replace=False)
        amax[i, rows] = rng.choice([0.2, 0.9049, 0.905, 0.9051, 0.89])
        arg[i, rows[0]] = arg[i, rows[1]]
    got = _native.track_scan(arg, amax, nf, n2)
    assert np.array_equal(got, _python_scan(arg, amax, n2))


def test_subset_scan_matches_reference_semantics(built_lib):
    """no_of_waves = m: the position-based tracking of solver.py:151-182 on the m selected
    columns, restated literally from |biorth| matrices, against _subset_scan on native tables."""
    from axisafe_b200 import solver
    rng = np.random.default_rng(5)
    n2, m, nf = 14, 5, 40
    rank = np.array([rng.permutation(n2) for _ in range(nf)], dtype=np.int32)
    sel = np.argsort(rank, axis=1)[:, :m]                       # native column of position p
    P = rng.random([nf - 1, n2, n2]) * 0.85                     # |biorth| in native indices
    for i in range(nf - 1):                                     # one dominant entry per row
        for a, b in zip(rng.permutation(n2), rng.permutation(n2)):
            P[i, a, b] = 0.97 if rng.random() > 0.08 else 0.5
    # device tables: arg-max over the SELECTED columns only (the others are zeroed)
    arg = np.zeros([nf - 1, n2], np.int32)
    amax = np.zeros([nf - 1, n2])
    for i in range(nf - 1):
        masked = np.zeros_like(P[i])
        masked[:, sel[i + 1]] = P[i][:, sel[i + 1]]
        arg[i], amax[i] = masked.argmax(1), masked.max(1)
    got = solver._subset_scan(arg, amax, rank, m)
    cols = np.empty([nf, m], np.int64)
    cols[0] = sel[0]
    fixups = 0
    for i in range(1, nf):
        biorth = P[i - 1][np.ix_(cols[i - 1], sel[i])]
        new_order = np.argmax(abs(biorth), axis=1)
        valid = np.round(abs(biorth[range(len(biorth)), list(new_order)]), 2) > 0.9
        valid_entries = np.delete(new_order, np.where(valid == False))   # noqa: E712
        if not valid.all():
            fixups += 1
            unassigned = list(set(range(len(biorth))) - set(valid_entries))
            j = 0
            for ii in range(len(valid)):
                if not valid[ii]:
                    new_order[ii] = unassigned[j]
                    j += 1
        cols[i] = sel[i][new_order]
    assert fixups > 3
    assert np.array_equal(got, cols)


def test_solve_rejects_non_finite_frequencies():
    """Reference error behaviour: scipy.linalg.eig(check_finite=True) -> ValueError."""
    import contextlib
    import io
    import axisafe_b200 as ax
    sol = ax.solver.WaveElementAxisym(ax.mesh.Mesh())
    with contextlib.redirect_stdout(io.StringIO()):
        for bad in ([1e3, np.nan], [np.inf]):
            with pytest.raises(ValueError, match='infs or NaNs'):
                sol.solve(np.array(bad))


# ------------------------------------------------------------- aggressive early deflation (host build)
def _aed_host_exe():
    """tests/host/aed_host: csrc/aed.cuh compiled as sequential host code (the same source the kernel
    k_hqr_mb runs per warp) + a model of the kernel's batch loop."""
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    src = os.path.join(ROOT, 'tests', 'host', 'aed_host.cu')
    exe = os.path.join(ROOT, 'tests', 'host', 'aed_host')
    deps = [src, os.path.join(ROOT, 'tests', 'host', 'warp_emul.h'),
            os.path.join(ROOT, 'axisym-safe-python_b200', 'csrc', 'aed.cuh'),
            os.path.join(ROOT, 'axisym-safe-python_b200', 'csrc', 'aed_variants.cuh'),
            os.path.join(ROOT, 'axisym-safe-python_b200', 'csrc', 'common.cuh')]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.run([nvcc, '-x', 'cu', '-DAED_HOST', '-O2', '-Wno-deprecated-gpu-targets', '-o', exe, src], check=True)
    return exe


def _aed_run(exe, mode, H, l, u, nw, nb=8, reg=0, seed=None):
    """reg = 0: the product routine (aed_window, shared memory) as sequential code; reg = 1 / 2: the A/B
    formulations of csrc/aed_variants.cuh (aed_window_reg / aed_window_flat) on the emulated warp of
    tests/host/warp_emul.h; seed: the emulator visits the lanes in a pseudo-random order (race check)."""
    import struct
    import subprocess
    import tempfile
    n = H.shape[0]
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', n, l, u, nw, nb))
            f.write(np.asarray(H, dtype=complex).tobytes(order='F'))
        env = dict(os.environ, AED_REG=str(int(reg)))
        env.pop('WEMU_SEED', None)
        if seed is not None:
            env['WEMU_SEED'] = str(seed)
        subprocess.run([exe, mode, fi, fo], check=True, env=env)
        raw = open(fo, 'rb').read()
    if mode == 'solve':
        return struct.unpack('5i', raw[:20]), np.frombuffer(raw[20:], dtype=complex)
    ns, ok, iters = struct.unpack('3i', raw[:12])
    a = np.frombuffer(raw[12:], dtype=complex)
    return ns, ok, iters, a[:64].reshape(8, 8), a[64:128].reshape(8, 8), a[128:136], a[136]


def _safe_hessenberg(i, nf=2000):
    import scipy.linalg as la
    from oracle import safe_oracle as so
    from axisafe_b200 import cases
    c = cases.pipe_soil_pml(nf=nf)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    S = so.S_matrix(so.operators(G, c['n']), 2 * np.pi * c['f'][i])
    Sb, _ = la.matrix_balance(S, permute=False)
    return np.triu(la.hessenberg(Sb), -1), np.linalg.eigvals(S)


@pytest.mark.parametrize('reg', [0, 1, 2], ids=['shared', 'registers', 'one-loop'])
def test_aed_window_is_a_similarity(reg):
    """One AED call (zlaqr2 semantics): Z unitary, Z^H W Z = T with T Hessenberg in its leading ns x ns
    part and triangular below, the spike vector reduced to one entry, every deflated eigenvalue
    meeting the spike criterion, spectrum of the window preserved."""
    exe = _aed_host_exe()
    eps = np.finfo(float).eps
    rng = np.random.default_rng(5)
    H, _ = _safe_hessenberg(700)
    n = H.shape[0]
    # a few QR-like iterations so that the trailing window carries converged eigenvalues
    cases_ = []
    for nw in (8, 5, 3):
        for u in (n - 1, 60):
            cases_.append((H, 0, u, nw))
    R = np.triu(rng.standard_normal([20, 20]) + 1j * rng.standard_normal([20, 20]), -1)
    R[np.arange(13, 20), np.arange(12, 19)] *= 1e-9          # nearly decoupled tail: deflation must trigger
    cases_ += [(R, 0, 19, 8), (R, 4, 19, 6), (R, 12, 19, 8)]
    deflated_total = 0
    for Hm, l, u, nw in cases_:
        nw = min(nw, u - l + 1)
        kw = u - nw + 1
        ns, ok, iters, T, Z, shift, spike_new = _aed_run(exe, 'window', Hm, l, u, nw, reg=reg)
        assert ok == 1 and 0 <= ns <= nw
        T, Z = T[:nw, :nw], Z[:nw, :nw]
        W = Hm[kw:u + 1, kw:u + 1]
        scale = np.abs(W).max()
        assert np.abs(Z.conj().T @ Z - np.eye(nw)).max() < 50 * eps
        if ns == nw:
            # nothing deflated: T / Z are the Schur form (the kernel discards them), shifts = its diagonal
            assert np.abs(np.tril(T, -1)).max() == 0.0
            assert np.abs(Z.conj().T @ W @ Z - T).max() < 200 * eps * scale
            assert np.allclose(np.sort_complex(shift[:ns]), np.sort_complex(np.diag(T)))
            continue
        deflated_total += nw - ns
        assert np.abs(Z.conj().T @ W @ Z - T).max() < 200 * eps * scale
        assert np.abs(np.tril(T[:ns, :ns], -2)).max() == 0.0 if ns else True
        assert np.abs(np.tril(T[ns:, ns:], -1)).max() == 0.0
        assert np.abs(T[ns:, :ns]).max() == 0.0 if ns else True
        spike = Hm[kw, kw - 1] if kw > l else 0.0
        sv = np.conj(Z[0, :]) * spike                         # Z^H (spike e_1)
        for j in range(ns, nw):                               # deflated: negligible spike entries
            foo = abs(T[j, j].real) + abs(T[j, j].imag) or abs(spike)
            assert abs(sv[j].real) + abs(sv[j].imag) <= 2.1 * max(np.finfo(float).tiny / eps, eps * foo)
        if ns:
            assert abs(sv[0] - spike_new) <= 200 * eps * abs(spike)
            assert np.abs(sv[1:ns]).max() <= 200 * eps * abs(spike) if ns > 1 else True
        ev = np.sort_complex(np.linalg.eigvals(W))
        got = np.sort_complex(np.concatenate([np.linalg.eigvals(T[:ns, :ns]) if ns else [], np.diag(T)[ns:]]))
        assert np.abs(ev - got).max() < 1e-9 * max(np.abs(ev).max(), 1e-300)
    assert deflated_total > 0


def test_aed_variants_match_the_product_window():
    """The A/B formulations against the product routine on 60 windows (sizes 2..8, SAFE Hessenberg blocks
    and random nearly decoupled ones).  The register-resident routine (lane 4 i + p owns T(i, 2p..2p+1), Z
    likewise; rotations through shuffles): same number of deflations and QR sweeps, same shifts; T and Z are
    both valid Schur data of the same window (test_aed_window_is_a_similarity), compared loosely because
    Schur vectors of close eigenvalues are ill-conditioned.  The one-loop routine: bit-identical.  Both give
    bit-identical results whatever order the emulated warp runs its lanes in (no missing warp barrier)."""
    exe = _aed_host_exe()
    rng = np.random.default_rng(1)
    H, _ = _safe_hessenberg(700)
    n = H.shape[0]
    tests = [(H, 0, u, nw) for u in (n - 1, 100, 60, 20, 9) for nw in (8, 7, 6, 5, 4, 3, 2)]
    for t in range(25):
        m = int(rng.integers(9, 24))
        R = np.triu(rng.standard_normal([m, m]) + 1j * rng.standard_normal([m, m]), -1)
        if t % 2:
            q = int(rng.integers(2, 7))
            R[np.arange(m - q, m), np.arange(m - q - 1, m - 1)] *= 10.0 ** (-int(rng.integers(6, 16)))
        tests.append((R, int(rng.integers(0, m - 8)), m - 1, int(rng.integers(2, 9))))
    deflated = 0
    for Hm, l, u, nw in tests:
        nw = min(nw, u - l + 1)
        a = _aed_run(exe, 'window', Hm, l, u, nw, reg=0)
        b = _aed_run(exe, 'window', Hm, l, u, nw, reg=1)
        f = _aed_run(exe, 'window', Hm, l, u, nw, reg=2)
        raw = lambda o: [v if isinstance(v, int) else np.asarray(v).tobytes() for v in o]
        assert raw(f) == raw(a), (l, u, nw)
        for variant, first in ((1, b), (2, f)):
            assert raw(_aed_run(exe, 'window', Hm, l, u, nw, reg=variant, seed=len(tests) + l + u)) == raw(first)
        assert a[:3] == b[:3], (a[:3], b[:3], l, u, nw)
        ns = a[0]
        deflated += nw - ns
        scale = np.abs(Hm[u - nw + 1:u + 1, u - nw + 1:u + 1]).max()
        if ns:
            assert np.abs(a[5][:ns] - b[5][:ns]).max() <= 1e-13 * scale
        assert np.abs(a[3][:nw, :nw] - b[3][:nw, :nw]).max() <= 1e-8 * scale
        assert np.abs(a[4][:nw, :nw] - b[4][:nw, :nw]).max() <= 1e-8
        if ns < nw:
            assert abs(a[6] - b[6]) <= 1e-13 * scale
    assert deflated > 20


@pytest.mark.parametrize('reg', [0, 1, 2], ids=['shared', 'registers', 'one-loop'])
@pytest.mark.parametrize('i', [3, 1400])
def test_aed_solve_matches_zgeev_and_cuts_the_sweeps(i, reg):
    """The kernel's batch loop with the AED window in front of every batch (host model): same
    eigenvalues as LAPACK, no failures, less than 70 % of the macro steps of the plain cascade."""
    from oracle import safe_oracle as so
    exe = _aed_host_exe()
    H, ref = _safe_hessenberg(i)
    n = H.shape[0]
    plain, ev0 = _aed_run(exe, 'solve', H, 0, n - 1, 0)
    aed, ev8 = _aed_run(exe, 'solve', H, 0, n - 1, 8, reg=reg)
    assert plain[4] == 0 and aed[4] == 0
    assert so.match_eigenvalues(ref, ev0)[1] < 1e-9
    assert so.match_eigenvalues(ref, ev8)[1] < 1e-9
    assert aed[2] > 0 and aed[3] > 0.8 * n               # most eigenvalues deflate inside the window
    assert aed[0] < 0.7 * plain[0], (aed, plain)


# --------------------------------------------------- the QR kernel itself on an emulated thread block
def _emul_exe(name, headers):
    """tests/host/<name>: a kernel header of csrc/ compiled as plain C++ against tests/host/shim/ and run
    on the emulated CTA of tests/host/cta_emul.h."""
    import shutil
    import subprocess
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('g++ not available')
    host = os.path.join(ROOT, 'tests', 'host')
    csrc = os.path.join(ROOT, 'axisym-safe-python_b200', 'csrc')
    exe = os.path.join(host, name)
    deps = [os.path.join(host, name + '.cpp'), os.path.join(host, 'cta_emul.h'),
            os.path.join(host, 'shim', 'cuda_runtime.h'), os.path.join(host, 'shim', 'cuda_pipeline.h'),
            os.path.join(host, 'shim', 'dmma.h'), os.path.join(host, 'shim', 'cooperative_groups.h')] + \
           [os.path.join(csrc, h) for h in headers]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.run([gxx, '-std=c++17', '-O2', '-I', os.path.join(host, 'shim'), '-o', exe,
                        os.path.join(host, name + '.cpp')], check=True)
    return exe


def _hqr_emul_exe():
    """csrc/hqr_mb.cuh: the product's multi-bulge QR kernel."""
    return _emul_exe('hqr_emul', ('hqr_mb.cuh', 'aed.cuh', 'common.cuh'))


def _hqr_emul_run(exe, H, aed=0, mask=15, bulges=8, seed=None):
    import struct
    import subprocess
    import tempfile
    n = H.shape[0]
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', n, aed, mask, bulges, 0))
            f.write(np.asarray(H, dtype=complex).tobytes(order='F'))
        env = dict(os.environ)
        env.pop('CTA_SEED', None)
        if seed is not None:
            env['CTA_SEED'] = str(seed)
        subprocess.run([exe, fi, fo], check=True, env=env)
        raw = open(fo, 'rb').read()
    info, macro, batches, aed_calls, _, launches = struct.unpack('6i', raw[:24])
    return dict(info=info, macro=macro, batches=batches, aed_calls=aed_calls, launches=launches), \
        np.frombuffer(raw[24:], dtype=complex)


def _case_hessenberg(case, i):
    import scipy.linalg as la
    from oracle import safe_oracle as so
    G = so.assemble(so.build_mesh(case['sets'], case['fmax'], case['PML'], case['PML_props']), case['n'])
    S = so.S_matrix(so.operators(G, case['n']), 2 * np.pi * case['f'][i])
    Sb, _ = la.matrix_balance(S, permute=False)
    return np.triu(la.hessenberg(Sb), -1), np.linalg.eigvals(S)


HQR_EMUL_CASES = [('pipe_soil_pml', {'nf': 2000}, 700, 4), ('free_steel_pipe', {}, 100, 1), ('muggleton_submerged', {}, 50, 1),
                  ('aristegui', {}, 20, 4)]


@pytest.mark.parametrize('name,kw,i,launches', HQR_EMUL_CASES, ids=[c[0] for c in HQR_EMUL_CASES])
def test_hqr_kernel_on_emulated_cta(name, kw, i, launches):
    """k_hqr_mb - the product kernel source, every instance of the shrinking-window cascade the problem
    size selects (2N = 126: four launches at 1 / 2 / 3 / 4 CTAs per SM; 2N = 110: four launches from the
    2-per-SM instance down; 2N = 68: the 5-per-SM instance alone; 2N = 30: the 128-thread instance) -
    on the emulated thread block: LAPACK's eigenvalues (1e-9 after the unitary Hessenberg reduction of
    SciPy), no unconverged eigenvalue, the expected number of launches; with the early-deflation window
    (8 and 6) the same spectrum in fewer macro steps."""
    from axisafe_b200 import cases
    from oracle import safe_oracle as so
    exe = _hqr_emul_exe()
    H, ref = _case_hessenberg(getattr(cases, name)(**kw), i)
    st, ev = _hqr_emul_run(exe, H)
    assert st['info'] == 0 and st['launches'] == launches and st['aed_calls'] == 0
    assert so.match_eigenvalues(ref, ev)[1] < 1e-9
    for aed in (8, 6):
        st_a, ev_a = _hqr_emul_run(exe, H, aed=aed)
        assert st_a['info'] == 0 and st_a['aed_calls'] > 0
        assert so.match_eigenvalues(ref, ev_a)[1] < 1e-9
        assert st_a['macro'] < st['macro']


def test_hqr_kernel_has_no_shared_memory_race():
    """Between two barriers the emulated threads run one after the other; CTA_SEED shuffles that order
    in every scheduling round.  A kernel without shared-memory races gives bit-identical eigenvalues and
    counters for every order (512 threads, 2N = 126, ~7800 block barriers; also with the one-warp
    early-deflation window between the sweeps and with fewer bulges in flight)."""
    exe = _hqr_emul_exe()
    H, _ = _safe_hessenberg(1400)
    for aed, bulges in ((0, 8), (8, 8), (0, 5)):
        st0, ev0 = _hqr_emul_run(exe, H, aed=aed, bulges=bulges)
        for seed in (1, 2):
            st, ev = _hqr_emul_run(exe, H, aed=aed, bulges=bulges, seed=seed)
            assert st == st0 and np.array_equal(ev, ev0), (aed, bulges, seed)


def _modes_emul_run(exe, H, ev, nw=4, seed=None):
    import struct
    import subprocess
    import tempfile
    n = H.shape[0]
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', n, nw, 0, 0, 0))
            f.write(np.asarray(H, dtype=complex).tobytes(order='F'))
            f.write(np.asarray(ev, dtype=complex).tobytes())
        env = dict(os.environ)
        env.pop('CTA_SEED', None)
        if seed is not None:
            env['CTA_SEED'] = str(seed)
        subprocess.run([exe, fi, fo], check=True, env=env)
        return np.frombuffer(open(fo, 'rb').read(), dtype=complex).reshape(n, n)


@pytest.mark.parametrize('name,kw,i', [c[:3] for c in HQR_EMUL_CASES], ids=[c[0] for c in HQR_EMUL_CASES])
def test_modes_kernel_on_emulated_cta(name, kw, i):
    """k_modes (csrc/modes.cuh, the product source: bottom-up pivoted LU of H - k I fused with the back
    substitution, four lane groups in a skewed pipeline, rows staged by cp.async double buffers - which the
    shim DEFERS until the matching wait) on the emulated thread block: every column is an eigenvector of
    the Hessenberg matrix (residual 1e-12 of ||H|| ||x||, measured 3e-15), scaled to unit max-norm, and
    the result does not depend on the order in which the lanes run between two warp barriers."""
    from axisafe_b200 import cases
    exe = _emul_exe('modes_emul', ('modes.cuh', 'common.cuh'))
    H, _ = _case_hessenberg(getattr(cases, name)(**kw), i)
    ev = np.linalg.eigvals(H)
    X = _modes_emul_run(exe, H, ev)
    scale = (np.abs(X.real) + np.abs(X.imag)).max(axis=0)
    assert np.abs(scale - 1).max() < 1e-12
    res = np.abs(H @ X - X * ev[None, :]).max(axis=0) / (np.abs(H).max() * np.abs(X).max(axis=0))
    assert res.max() < 1e-12, res.max()
    assert np.array_equal(_modes_emul_run(exe, H, ev, seed=5), X)
    if H.shape[0] <= 68:                                 # two-warp CTAs (16 modes per CTA) give the same vectors
        assert np.array_equal(_modes_emul_run(exe, H, ev, nw=2), X)


def _sweep_emul_point(name, kw, i, seed=None):
    """Input preparation + one run of tests/host/sweep_emul for frequency i of a case; returns (op, w, S, raw output)."""
    import struct
    import subprocess
    import tempfile
    from axisafe_b200 import cases
    from oracle import safe_oracle as so
    exe = _emul_exe('sweep_emul', ('reduce_oc.cuh', 'reduce_l2.cuh', 'hqr_mb.cuh', 'modes.cuh', 'backnorm.cuh', 'backx_kernels.cuh', 'aed.cuh', 'common.cuh'))
    c = getattr(cases, name)(**kw)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    w = 2 * np.pi * c['f'][i]
    N = op['N']
    S = so.S_matrix(op, w)
    C = op['C']
    r, cc = np.nonzero(C)
    hb = int(np.abs(r - cc).max())
    Cb = np.zeros([N, 2 * hb + 1], dtype=complex)                  # banded rows: Cb[r, c - r + hb] = C[r, c]
    for d in range(-hb, hb + 1):
        idx = np.arange(max(0, -d), min(N, N - d))
        Cb[idx, d + hb] = C[idx, idx + d]
    env = dict(os.environ)
    env.pop('CTA_SEED', None)
    if seed is not None:
        env['CTA_SEED'] = str(seed)
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', N, hb, 0, 0, 0))
            f.write(np.asarray(S, dtype=complex).tobytes(order='F'))
            f.write(Cb.tobytes())
            f.write(np.diag(op['K6']).astype(complex).tobytes())
        subprocess.run([exe, fi, fo], check=True, env=env)
        return op, w, S, open(fo, 'rb').read()


SWEEP_EMUL_CASES = [('pipe_soil_pml', dict(nf=2000), 700), ('aristegui', {}, 50), ('nguyen', dict(n=1), 40),
                    ('free_steel_pipe', dict(n=1), 250), ('muggleton_submerged', {}, 60), ('gravenkamp', dict(n=1), 30)]


@pytest.mark.parametrize('name,kw,i', SWEEP_EMUL_CASES, ids=[c[0] for c in SWEEP_EMUL_CASES])
def test_frequency_point_on_emulated_gpu(name, kw, i):
    """One frequency point through the product's own kernel sources on the emulated thread block
    (tests/host/sweep_emul.cpp): k_balance -> k_reduce_oc -> k_hqr_mb cascade -> k_modes -> k_backnorm, i.e.
    scipy.linalg.eig + B-normalisation of reference solver.py:136-142, without a GPU.  2N = 126 and 110 take
    the 12-warp padded reduction and the leader-isolated QR stage 1, 2N = 166 (the largest size of the
    shared-memory path) the 8-warp unpadded reduction, the plain stage 1 and three-warp k_modes CTAs; 2N = 30,
    68 and 68 (BASELINE.json configs[1] and [2]: free pipe, submerged pipe) the L2-resident reduction k_reduce
    behind k_balance and the QR instance that fits the matrix.
    Bars as on the GPU: balancing factors identical to LAPACK's zgebal, eigenvalues within the 1e-8
    bar (conditioning allowance as in the GPU tests), psi^T B psi = I, upper half = k x lower half, mode
    shapes within 1e-6 of the oracle's B-normalised zgeev vectors."""
    import struct
    import scipy.linalg as la
    from oracle import safe_oracle as so
    op, w, S, raw = _sweep_emul_point(name, kw, i)
    N = op['N']
    n2 = 2 * N
    info, launches = struct.unpack('2i', raw[:8])
    wy_diff, = struct.unpack('d', raw[8:16])
    scale = np.frombuffer(raw[16:16 + 8 * n2], dtype=float)
    k = np.frombuffer(raw[16 + 8 * n2:16 + 24 * n2], dtype=complex)
    psi = np.frombuffer(raw[16 + 24 * n2:], dtype=complex).reshape(n2, n2)
    assert info == 0
    # 16 <= 2N <= 128: psi comes from the blocked DMMA back-transformation (the library's default there) and
    # agrees with the per-reflector kernel; above that only the per-reflector kernel exists
    assert (0 <= wy_diff < 1e-11) if n2 <= 128 else wy_diff == -1.0
    _, (ref_scale, _) = la.matrix_balance(S, permute=False, separate=True)
    assert np.array_equal(scale, ref_scale)
    val, vec = so.eig_normalised(op, w)
    perm, _ = so.match_eigenvalues(val, k)
    assert np.all(np.abs(val - k[perm]) <= so.eig_tolerance(S, val))
    if name == 'pipe_soil_pml':
        assert np.all(np.abs(val - k[perm]) <= 1e-8 * np.abs(val))       # flagship: plain bar
    assert np.abs(psi[:N] - k * psi[N:]).max() <= 1e-13 * np.abs(psi[:N]).max()
    assert np.abs(np.diag(psi.T @ op['B'] @ psi) - 1).max() < 1e-9
    errs = [so.mode_shape_error(vec[N:, [j]], psi[N:, [perm[j]]])[0] for j in range(n2)]
    assert max(errs) <= 1e-6, max(errs)


# ----------------------------------------------------------- kernel 1 on the emulated thread block
def _element_arrays(elements):
    """Descriptor / table / node arrays of a list of element objects, as _native.element_matrices builds them."""
    from axisafe_b200 import shape_functions as sf
    tables, table_off, cache = [], {}, 0
    desc_i, desc_d, nodes, offsets, sizes = [], [], [], [], []
    node_off = out_off = 0
    for el in elements:
        (kind, family, pml, m), dbl = el._descriptor()
        key = (family, m)
        if key not in table_off:
            ksi, w = (sf.build_GLJ_quadrature if family == 1 else sf.build_GLL_quadrature)(m)
            D = (sf.lagrange_GLJ if family == 1 else sf.lagrange_poly)(ksi)[1]
            table_off[key] = cache
            block = np.concatenate([ksi, w, D.ravel()])
            tables.append(block)
            cache += len(block)
        nd = (3 if kind == 0 else 1) * m
        desc_i.append([kind, family, pml, m, table_off[key], node_off, out_off, 0])
        desc_d.append(list(dbl) + [0.0] * (12 - len(dbl)))
        nodes.append(np.asarray(el.nodes_at, dtype=np.float64))
        offsets.append(out_off)
        sizes.append(nd)
        node_off += m
        out_off += 10 * nd * nd
    return desc_i, desc_d, tables, nodes, offsets, sizes, out_off


def _elements_emulated(elements):
    """k_element_matrices on the emulated thread block for stand-alone element objects; installs the results."""
    import struct
    import subprocess
    import tempfile
    exe = _emul_exe('assemble_emul', ('assemble_kernels.cuh', 'common.cuh'))
    desc_i, desc_d, tables, nodes, offsets, sizes, out_off = _element_arrays(elements)
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            tab, nod = np.concatenate(tables), np.concatenate(nodes)
            # no scatter targets: one DOF, every local DOF unmapped (-1)
            full = sum(el._full_dofs for el in elements)
            lm_off = np.cumsum([0] + [el._full_dofs for el in elements])[:-1]
            f.write(struct.pack('10i', len(elements), len(tab), len(nod), out_off, full, 1, 0, 0, 0, 0))
            for a, t in ((desc_i, np.int32), (desc_d, np.float64), (tab, np.float64), (nod, np.float64),
                         (-np.ones(full), np.int32), (lm_off, np.int32), (np.zeros(1), np.int32)):
                f.write(np.ascontiguousarray(np.array(a), dtype=t).tobytes())
        subprocess.run([exe, fi, fo], check=True)
        host = np.frombuffer(open(fo, 'rb').read(), dtype=complex)[:out_off]
    for el, off, nd in zip(elements, offsets, sizes):
        el._load(host[off:off + 10 * nd * nd].reshape(10, nd, nd).copy())


def _assemble_emulated(mesh, n, omega):
    """Mesh.assemble_matrices(n) with the two library calls (axs_element_matrices, axs_assemble_operators) and
    axs_build_S replaced by the SAME kernels on the emulated thread block (tests/host/assemble_emul.cpp).  The
    array preparation mirrors _native.element_matrices / _native.assemble_operators; everything else (element
    objects, DOF numbering, coupling blocks, COO scatter of the global matrices) is the product's host code.
    Returns (banded K0s, Cb, Mb [N, W], K6d [N], K1c or None, hb, S [nf, 2N, 2N])."""
    import struct
    import subprocess
    import tempfile
    from axisafe_b200.mesh import _GLOBAL_NAME
    exe = _emul_exe('assemble_emul', ('assemble_kernels.cuh', 'common.cuh'))
    mesh.n = n
    mesh._instantiate_elements(n)
    mesh._number_dofs()
    order = sorted(mesh.IEN)
    elements = [mesh.elements[el] for el in order]
    desc_i, desc_d, tables, nodes, offsets, sizes, out_off = _element_arrays(elements)
    N = mesh.no_of_dofs
    lm_all, lm_off, hb = [], [], 0
    for el in order:                                      # as _native.assemble_operators
        obj = mesh.elements[el]
        reduced = iter(mesh.LM[el])
        full = [-1 if i in obj._dropped else next(reduced) for i in range(obj._full_dofs)]
        live = [g for g in full if g >= 0]
        if live:
            hb = max(hb, max(live) - min(live))
        lm_off.append(len(lm_all))
        lm_all.extend(full)
    tclass = np.array([0 if t == 1 else 1 for t in mesh.Tdiag], dtype=np.int32)

    def run(hb, K1c, omega):
        W = 2 * hb + 1
        with tempfile.TemporaryDirectory() as d:
            fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
            with open(fi, 'wb') as f:
                tab, nod = np.concatenate(tables), np.concatenate(nodes)
                f.write(struct.pack('10i', len(elements), len(tab), len(nod), out_off, len(lm_all), N, hb, int(n),
                                    len(omega), 0 if K1c is None else 1))
                for a, t in ((desc_i, np.int32), (desc_d, np.float64), (tab, np.float64), (nod, np.float64),
                             (lm_all, np.int32), (lm_off, np.int32), (tclass, np.int32)):
                    f.write(np.ascontiguousarray(np.array(a), dtype=t).tobytes())
                if K1c is not None:
                    f.write(np.ascontiguousarray(K1c, dtype=complex).tobytes())
                f.write(np.ascontiguousarray(omega, dtype=np.float64).tobytes())
            subprocess.run([exe, fi, fo], check=True)
            raw = np.frombuffer(open(fo, 'rb').read(), dtype=complex)
        o = [0]

        def take(cnt, shape):
            a = raw[o[0]:o[0] + cnt].reshape(shape)
            o[0] += cnt
            return a
        return (take(out_off, -1), take(N * W, (N, W)), take(N * W, (N, W)), take(N * W, (N, W)), take(N, -1),
                take(len(omega) * 4 * N * N, (len(omega), 2 * N, 2 * N)))
    host = run(hb, None, [])[0]                           # element matrices first: the coupling blocks need them
    for el, off, nd in zip(elements, offsets, sizes):
        el._load(host[off:off + 10 * nd * nd].reshape(10, nd, nd).copy())
    mesh._coupling_blocks()
    for name, matrix in mesh._scatter(list(mesh.LM)).items():
        setattr(mesh, _GLOBAL_NAME[name], matrix)
    K1c = None
    if mesh.is_coupled and mesh.K1_coupling is not None and mesh.K1_coupling.nnz:
        coo = mesh.K1_coupling.tocoo()
        hb = max(hb, int(np.max(np.abs(coo.row - coo.col))))
        K1c = np.zeros([N, 2 * hb + 1], complex)
        np.add.at(K1c, (coo.row, coo.col - coo.row + hb), coo.data)
    _, K0s, Cb, Mb, K6d, S = run(hb, K1c, omega)
    return K0s, Cb, Mb, K6d, K1c, hb, S.transpose(0, 2, 1)        # column-major -> [f, r, c]


@pytest.mark.parametrize('name,n,kw', GOLDEN_CASES, ids=CASE_IDS)
def test_assembly_kernels_on_emulated_gpu(name, n, kw):
    """Kernel 1 without a GPU: k_element_matrices, k_scatter_operators and k_build_S (the product sources on the
    emulated thread block) inside the product's own host assembly, against the fixtures written by the
    unmodified reference - the CPU twin of test_gpu_parity.py::test_global_matrices_vs_reference and
    ::test_S_batch_vs_reference (same bars: global matrices 2e-12, complex64 mass entries bit-exact, banded
    operators 2e-12 against the oracle, S(w) 1e-11 against the reference's SuperLU-built S)."""
    import axisafe_b200 as ax
    from oracle import safe_oracle as so
    g = load_golden(name, n)
    c = make_case(name, n, kw)
    m = ax.mesh.Mesh()
    m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    w = 2 * np.pi * c['f'][g['pick']]
    K0s, Cb, Mb, K6d, K1c, hb, S = _assemble_emulated(m, n, w)
    for key in ('K1', 'K2', 'K3', 'K4', 'K5', 'K6', 'M', 'KF', 'KFt', 'KFz'):
        ref = golden_dense(g, key)
        got = getattr(m, key).toarray()
        assert np.abs(got - ref).max() <= 2e-12 * max(np.abs(ref).max(), 1e-300), key
    ref_M = golden_dense(g, 'M')
    sol = [d for d in g['solid_dofs'] if ref_M[d, d] == np.complex64(ref_M[d, d])]
    assert np.array_equal(m.M.toarray()[sol, sol], ref_M[sol, sol])       # complex64 quirk, exact
    if bool(g['is_coupled']):
        ref = golden_dense(g, 'K1_coupling')
        assert np.abs(m.K1_coupling.toarray() - ref).max() <= 1e-12 * np.abs(ref).max()
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), n)
    op = so.operators(G, n)
    N = m.no_of_dofs

    def dense(band):
        out = np.zeros([N, N], complex)
        for r in range(N):
            for col in range(max(0, r - hb), min(N, r + hb + 1)):
                out[r, col] = band[r, col - r + hb]
        return out
    for band, key in ((K0s, 'K0s'), (Cb, 'C'), (Mb, 'M')):
        assert np.abs(dense(band) - op[key]).max() <= 2e-12 * max(np.abs(op[key]).max(), 1e-300), key
    assert np.abs(K6d - np.diag(op['K6'])).max() <= 2e-12 * np.abs(op['K6']).max()
    for j in range(len(w)):
        assert np.abs(S[j] - g['S'][j]).max() <= 1e-11 * np.abs(g['S'][j]).max()


@pytest.mark.parametrize('n', [0, 1, 2])
@pytest.mark.parametrize('m', [4, 5, 9, 14, 23])
def test_element_kernel_on_emulated_gpu(m, n):
    """Every element class x circumferential order x element order through k_element_matrices on the
    emulated thread block against the oracle's index-form restatement: the CPU twin of
    test_gpu_parity.py::test_element_matrices_vs_oracle (2e-12; complex64 mass diagonal bit-exact)."""
    import axisafe_b200 as ax
    from oracle import safe_oracle as so
    E = ax.elements
    lam, mu = ax.misc.young2lame(156e9, 0.2)
    solid, water, pml = [lam, mu, 7800, 0.03], [2.25e9, 1000, 0], [0.1, 0.01, 6 + 7j]
    gll = 0.1 + 0.01 * (so.gll(m)[0] + 1) / 2
    glj = 0.1 * (so.glj(m)[0] + 1) / 2
    glj[0] = 0.0
    cases_ = [('SLAX6', E.SLAX6(gll, n), gll, solid, None), ('ALAX6', E.ALAX6(gll, n), gll, water, None),
              ('SLAX6_core', E.SLAX6_core(glj, n), glj, solid, None),
              ('ALAX6_core', E.ALAX6_core(glj, n), glj, water, None),
              ('SLAX6_PML', E.SLAX6_PML(gll, n, pml), gll, solid, pml),
              ('ALAX6_PML', E.ALAX6_PML(gll, n, pml), gll, water, pml)]
    for name, el, nodes, mat, p in cases_:
        el.add_properties(mat)
    _elements_emulated([c[1] for c in cases_])
    for name, el, nodes, mat, p in cases_:
        ref = so.element_matrices(name, nodes, mat, n, p)
        for key in ('K_1', 'K_2', 'K_3', 'K_4', 'K_5', 'K_6', 'K_F', 'K_Ft', 'K_Fz', 'M'):
            a, b = np.asarray(getattr(el, key)), ref[key]
            assert a.shape == b.shape, (name, key, a.shape, b.shape)
            assert np.abs(a - b).max() <= 2e-12 * max(np.abs(b).max(), 1e-300), (name, key)
        if name in ('SLAX6', 'SLAX6_core'):
            assert el.M.dtype == np.complex64
            assert np.array_equal(np.diag(el.M), np.diag(ref['M'])), name   # fp32 rounding, bit exact
        for key in ('Hs0', 'Hs1', 'Hf0', 'Hf1'):
            if key in ref:
                assert np.allclose(getattr(el, key), ref[key], rtol=1e-15, atol=0), (name, key)


def test_tracking_kernels_on_emulated_gpu():
    """k_bpsi_tiled / k_bpsi (Y = B psi), k_track_mma (P = psi_prev^T Y on FP64 tensor-core tiles - the one
    inline-PTX instruction replaced by a lane-collective model of the m8n8k4 fragment layout,
    tests/host/shim/dmma.h - with the fused row arg-max) and k_apply_order on the emulated thread block, for six
    consecutive frequencies of the flagship sweep with the oracle's B-normalised zgeev vectors as input: the
    arg-max table is bit-exact against NumPy's |psi^T B psi|.argmax(axis=1), the row maxima agree to 1e-9 and
    the column gather equals NumPy's take - the CPU twin of test_gpu_parity.py::test_track_pairs_vs_numpy."""
    import struct
    import subprocess
    import tempfile
    from axisafe_b200 import cases
    from oracle import safe_oracle as so
    exe = _emul_exe('track_emul', ('track_kernels.cuh', 'common.cuh'))
    c = cases.pipe_soil_pml(nf=2000)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    N, n2, nf = op['N'], 2 * op['N'], 6
    vals, psis = zip(*[so.eig_normalised(op, 2 * np.pi * f) for f in c['f'][700:700 + nf]])
    k, psi = np.array(vals), np.array(psis)
    C = op['C']
    r, cc = np.nonzero(C)
    hb = int(np.abs(r - cc).max())
    Cb = np.zeros([N, 2 * hb + 1], dtype=complex)
    for d in range(-hb, hb + 1):
        idx = np.arange(max(0, -d), min(N, N - d))
        Cb[idx, d + hb] = C[idx, idx + d]
    rng = np.random.default_rng(3)
    order = np.array([rng.permutation(n2) for _ in range(nf)], dtype=np.int32)
    for tiled in (1, 0):
        with tempfile.TemporaryDirectory() as d:
            fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
            with open(fi, 'wb') as f:
                f.write(struct.pack('5i', N, hb, nf, tiled, 0))
                for a in (Cb, np.diag(op['K6']).astype(complex), np.ascontiguousarray(psi), order, np.ascontiguousarray(k)):
                    f.write(a.tobytes())
            subprocess.run([exe, fi, fo], check=True)
            raw = open(fo, 'rb').read()
        o = (nf - 1) * n2
        arg = np.frombuffer(raw[:4 * o], dtype=np.int32).reshape(nf - 1, n2)
        amax = np.frombuffer(raw[4 * o:12 * o], dtype=float).reshape(nf - 1, n2)
        rest = np.frombuffer(raw[12 * o:], dtype=complex)
        k_sorted, psi_sorted = rest[:nf * n2].reshape(nf, n2), rest[nf * n2:].reshape(nf, n2, n2)
        for p in range(nf - 1):
            P = np.abs(psi[p].T @ op['B'] @ psi[p + 1])
            assert np.array_equal(arg[p], P.argmax(axis=1)), (tiled, p)
            assert np.abs(amax[p] - P.max(axis=1)).max() < 1e-9
        for f in range(nf):
            assert np.array_equal(k_sorted[f], k[f][order[f]])
            assert np.array_equal(psi_sorted[f], psi[f][:, order[f]])


@pytest.mark.parametrize('name,kw,i', [SWEEP_EMUL_CASES[0], SWEEP_EMUL_CASES[4]], ids=['pipe_soil_pml', 'muggleton_submerged'])
def test_frequency_point_has_no_shared_memory_race(name, kw, i):
    """The whole kernel chain of a frequency point (k_balance, k_reduce_oc / k_reduce, k_hqr_mb, k_modes,
    k_wy_tfactor, k_backnorm_wy, k_backnorm) with the emulated threads visited in a new pseudo-random order
    between every two barriers: bit-identical output (balancing factors, eigenvalues, mode shapes)."""
    base = _sweep_emul_point(name, kw, i)[3]
    assert _sweep_emul_point(name, kw, i, seed=1)[3] == base


def test_hqr_kernel_at_the_instance_boundaries():
    """The QR cascade on the emulated thread block at the sizes where the launcher changes instance (the
    128-thread kernel holds 8 x 36 row-rotation items in exactly 3 x 96 slots at 2N = 35; the 5 / 4 / 3 / 2 / 1
    CTAs-per-SM windows end at 73 / 82 / 95 / 117; the leader-isolated stage 1 ends at 2N = 144) and at the
    smallest and largest sizes: random dense, row-graded (1e-3 .. 1e3) and nearly decoupled Hessenberg
    matrices against LAPACK.  A full sweep of 37 sizes x 3 kinds x AED on/off gave 4e-10 at worst."""
    import scipy.linalg as la
    from oracle import safe_oracle as so
    exe = _hqr_emul_exe()
    rng = np.random.default_rng(7)
    jobs = [(n, 'dense') for n in (4, 35, 36, 74, 83, 96, 118, 145, 166)] + [(36, 'graded'), (83, 'graded'), (96, 'decoupled')]
    for n, kind in jobs:
        A = rng.standard_normal([n, n]) + 1j * rng.standard_normal([n, n])
        if kind == 'graded':
            A *= (10.0 ** rng.uniform(-3, 3, size=n))[:, None]
        H = np.triu(la.hessenberg(la.matrix_balance(A, permute=False)[0]), -1)
        if kind == 'decoupled':
            H[n // 2, n // 2 - 1] = 0
            H[n - 2, n - 3] = 1e-300
        st, ev = _hqr_emul_run(exe, H)
        assert st['info'] == 0, (n, kind, st)
        assert so.match_eigenvalues(np.linalg.eigvals(H), ev)[1] < 1e-9, (n, kind)


@pytest.mark.parametrize('N', [8, 47, 48, 64, 65])
def test_frequency_point_random_pencils_at_the_boundaries(N):
    """The whole kernel chain on the emulated thread block for RANDOM quadratic pencils (K6 diagonal, C banded
    symmetric, K0 dense symmetric; S = B^-1 A as in solver.py:115-134) at the sizes where the launchers switch
    instance: 2N = 16 (smallest size of the blocked back-transformation), 94 / 96 (L2-resident vs on-chip
    reduction), 128 / 130 (padded 12-warp vs unpadded 8-warp reduction; blocked DMMA vs per-reflector
    back-transformation).  Same checks as on the reference cases; measured 1e-14 .. 5e-12."""
    import struct
    import subprocess
    import tempfile
    import scipy.linalg as la
    from oracle import safe_oracle as so
    exe = _emul_exe('sweep_emul', ('reduce_oc.cuh', 'reduce_l2.cuh', 'hqr_mb.cuh', 'modes.cuh', 'backnorm.cuh', 'backx_kernels.cuh', 'aed.cuh', 'common.cuh'))
    rng = np.random.default_rng(100 + N)
    hb = min(N - 1, int(rng.integers(1, 12)))
    n2 = 2 * N
    K6 = np.diag(rng.uniform(0.5, 2, N) * np.exp(1j * rng.uniform(-0.3, 0.3, N)))
    C = np.zeros([N, N], complex)
    for d in range(-hb, hb + 1):
        idx = np.arange(max(0, -d), min(N, N - d))
        C[idx, idx + d] = rng.standard_normal(len(idx)) + 1j * rng.standard_normal(len(idx))
    C = C + C.T
    K0 = rng.standard_normal([N, N]) + 1j * rng.standard_normal([N, N])
    K0 = (K0 + K0.T) * 10.0 ** rng.uniform(-1, 2)
    B = np.block([[np.zeros([N, N]), K6], [K6, C]])
    S = np.linalg.solve(B, np.block([[K6, np.zeros([N, N])], [np.zeros([N, N]), -K0]]))
    Cb = np.zeros([N, 2 * hb + 1], dtype=complex)
    for d in range(-hb, hb + 1):
        idx = np.arange(max(0, -d), min(N, N - d))
        Cb[idx, d + hb] = C[idx, idx + d]
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', N, hb, 0, 0, 0))
            f.write(np.asarray(S, dtype=complex).tobytes(order='F'))
            f.write(Cb.tobytes())
            f.write(np.diag(K6).astype(complex).tobytes())
        subprocess.run([exe, fi, fo], check=True)
        raw = open(fo, 'rb').read()
    info, _ = struct.unpack('2i', raw[:8])
    wy_diff, = struct.unpack('d', raw[8:16])
    scale = np.frombuffer(raw[16:16 + 8 * n2], dtype=float)
    k = np.frombuffer(raw[16 + 8 * n2:16 + 24 * n2], dtype=complex)
    psi = np.frombuffer(raw[16 + 24 * n2:], dtype=complex).reshape(n2, n2)
    assert info == 0
    assert (0 <= wy_diff < 1e-11) if n2 <= 128 else wy_diff == -1.0
    _, (ref_scale, _) = la.matrix_balance(S, permute=False, separate=True)
    assert np.array_equal(scale, ref_scale)
    val, vec = la.eig(S)
    vec = vec / np.sqrt(np.diag(vec.T @ B @ vec))
    perm, err = so.match_eigenvalues(val, k)
    assert err < 1e-10
    assert np.abs(np.diag(psi.T @ B @ psi) - 1).max() < 1e-9
    assert np.abs(psi[:N] - k * psi[N:]).max() <= 1e-13 * np.abs(psi[:N]).max()
    errs = [so.mode_shape_error(vec[N:, [j]], psi[N:, [perm[j]]])[0] for j in range(n2)]
    assert max(errs) < 1e-8, max(errs)


def _bigred_emul_run(exe, A, nb=32, seed=None):
    import struct
    import subprocess
    import tempfile
    n = A.shape[0]
    env = dict(os.environ)
    env.pop('CTA_SEED', None)
    if seed is not None:
        env['CTA_SEED'] = str(seed)
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', n, nb, 0, 0, 0))
            f.write(np.asarray(A, dtype=complex).tobytes(order='F'))
        subprocess.run([exe, fi, fo], check=True, env=env)
        raw = open(fo, 'rb').read()
    launches, ctas, have_q, _ = struct.unpack('4i', raw[:16])
    a = np.frombuffer(raw[16:], dtype=complex)
    return dict(launches=launches, ctas=ctas, have_q=have_q, A=a[:n * n].reshape(n, n).T, tau=a[n * n:n * n + n],
                Q=a[n * n + n:2 * n * n + n].reshape(n, n), Hc=a[2 * n * n + n:]), raw


@pytest.mark.parametrize('n,nb', [(168, 32), (294, 32), (200, 16)], ids=['2N=168', '2N=294', 'panel16'])
def test_blocked_reduction_on_emulated_gpu(n, nb):
    """North-star kernel (2) without a GPU: the blocked compact-WY Hessenberg reduction with its trailing
    updates on FP64 tensor-core tiles (k_br_col, k_br_gemv, k_br_gemm<YTOP / RIGHT / LEFTW / LEFTA>, k_br_top,
    k_br_tw, k_br_pack*) and the blocked back-transformation (k_br_gemm<BXW / BXA>, k_bx_tw, k_bx_scale) -
    the product sources, launch by launch and CTA by CTA on the emulated thread block, DMMA through the
    lane-collective model - for 2N = 168 (the smallest size of the large path) and 294: H is upper Hessenberg,
    Q = (back-transformation of the identity) is unitary, Q^H A Q = H to 1e-13 ||A||, the spectrum is kept and the
    packed copy handed to the QR kernel equals H.  Panels of 16 keep no T factors (no Q): spectrum only."""
    import scipy.linalg as la
    from oracle import safe_oracle as so
    exe = _emul_exe('bigred_emul', ('bigred_kernels.cuh', 'common.cuh'))
    rng = np.random.default_rng(n)
    A = rng.standard_normal([n, n]) + 1j * rng.standard_normal([n, n])
    A *= (10.0 ** rng.uniform(-1, 1, size=n))[:, None]
    A = la.matrix_balance(A, permute=False)[0]
    out, _ = _bigred_emul_run(exe, A, nb)
    H = np.triu(out['A'], -1)
    assert so.match_eigenvalues(np.linalg.eigvals(A), np.linalg.eigvals(H))[1] < 1e-11
    packed = np.concatenate([H[:min(j + 2, n), j] for j in range(n)])          # column j: rows 0 .. j+1
    assert np.array_equal(out['Hc'][:len(packed)], packed)
    if nb == 32:
        Q = out['Q']
        assert out['have_q'] == 1
        assert np.abs(Q.conj().T @ Q - np.eye(n)).max() < 1e-13
        assert np.abs(Q.conj().T @ A @ Q - H).max() < 1e-13 * np.abs(A).max()


def test_blocked_reduction_has_no_shared_memory_race():
    """Same kernels with the emulated threads of every CTA visited in a pseudo-random order between barriers:
    bit-identical output (the reductions of the reduction have a fixed summation order by construction)."""
    import scipy.linalg as la
    exe = _emul_exe('bigred_emul', ('bigred_kernels.cuh', 'common.cuh'))
    rng = np.random.default_rng(9)
    n = 168
    A = la.matrix_balance(rng.standard_normal([n, n]) + 1j * rng.standard_normal([n, n]), permute=False)[0]
    assert _bigred_emul_run(exe, A, seed=4)[1] == _bigred_emul_run(exe, A)[1]


def test_large_path_frequency_point_on_emulated_gpu():
    """One frequency point of cfg4's mesh family (coated pipe, 2 elements of order 8 per layer, 2N = 294) through
    the LARGE path on the emulated thread block, ~700 launches of the product's kernel sources: k_reduce_big
    (balancing) -> blocked compact-WY reduction on DMMA tiles -> k_hqr_big<false> (windowed multi-bulge QR) ->
    k_hqr_mb cascade -> k_modes_col -> k_backx_slab + k_norm_big, and the blocked DMMA back-transformation on
    the same vectors.  The CPU twin of test_gpu_big.py::test_coated_pipe_stages_vs_oracle (same bars); the
    thread-block-cluster instance of the windowed QR stays GPU-only."""
    import struct
    import subprocess
    import tempfile
    import scipy.linalg as la
    from axisafe_b200 import cases
    from oracle import safe_oracle as so
    exe = _emul_exe('big_emul', ('big_kernels.cuh', 'bigred_kernels.cuh', 'hqr_mb.cuh', 'aed.cuh', 'common.cuh'))
    c = cases.coated_pipe(nf=4, n=1, elements_per_layer=2, order=8)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    N, w = op['N'], 2 * np.pi * c['f'][2]
    n2 = 2 * N
    assert n2 == 294
    S = so.S_matrix(op, w)
    C = op['C']
    r, cc = np.nonzero(C)
    hb = int(np.abs(r - cc).max())
    Cb = np.zeros([N, 2 * hb + 1], dtype=complex)
    for d in range(-hb, hb + 1):
        idx = np.arange(max(0, -d), min(N, N - d))
        Cb[idx, d + hb] = C[idx, idx + d]
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', N, hb, 0, 0, 0))
            f.write(np.asarray(S, dtype=complex).tobytes(order='F'))
            f.write(Cb.tobytes())
            f.write(np.diag(op['K6']).astype(complex).tobytes())
        subprocess.run([exe, fi, fo], check=True)
        raw = open(fo, 'rb').read()
    info, launches = struct.unpack('2i', raw[:8])
    blocked_vs_slab, = struct.unpack('d', raw[8:16])
    scale = np.frombuffer(raw[16:16 + 8 * n2], dtype=float)
    k = np.frombuffer(raw[16 + 8 * n2:16 + 24 * n2], dtype=complex)
    psi = np.frombuffer(raw[16 + 24 * n2:], dtype=complex).reshape(n2, n2)
    assert info == 0 and launches > 600
    assert blocked_vs_slab < 1e-11
    _, (ref_scale, _) = la.matrix_balance(S, permute=False, separate=True)
    assert np.array_equal(scale, ref_scale)
    val, vec = so.eig_normalised(op, w)
    perm, _ = so.match_eigenvalues(val, k)
    assert np.all(np.abs(val - k[perm]) <= so.eig_tolerance(S, val))
    assert np.abs(psi[:N] - k * psi[N:]).max() <= 1e-13 * np.abs(psi[:N]).max()
    assert np.abs(np.einsum('ij,ij->j', psi, op['B'] @ psi) - 1).max() < 1e-9
    errs = so.mode_shape_error(vec[N:], psi[N:][:, perm])
    assert errs.max() <= 1e-6, errs.max()


def test_post_processing_kernels_on_emulated_gpu():
    """k_energy_ratio, k_quad_forms, k_modal_projection (rows N1 / N3 / N4: the dense batched products of
    reference solver.py:248-266, :374-408, :428-435) and k_rank_modes + k_zero_unselected (the selection behind
    solve(no_of_waves=m), solver.py:137-140) on the emulated thread block against NumPy, random data."""
    import struct
    import subprocess
    import tempfile
    exe = _emul_exe('post_emul', ('post_kernels.cuh', 'common.cuh'))
    rng = np.random.default_rng(21)
    N, nf, m = 37, 3, 20
    n = 2 * N
    crand = lambda *shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    psi = crand(nf, n, n)
    tclass = rng.integers(0, 2, N).astype(np.int32)

    def coo(nnz):
        return (rng.integers(0, N, nnz).astype(np.int32), rng.integers(0, N, nnz).astype(np.int32), crand(nnz))
    Mt, Mp = coo(150), coo(60)
    mats = [coo(int(z)) for z in (40, 1, 75)]
    sel = rng.permutation(n)[:11].astype(np.int32)
    mat_off = np.cumsum([0] + [len(x[0]) for x in mats]).astype(np.int32)
    rows, cols, vals = (np.concatenate([x[q] for x in mats]) for q in range(3))
    fvec = crand(n)
    fvec[:N] = 0
    k = crand(nf, n)
    k[1, 5] = k[1, 9]                                            # a tie: the lower native index ranks first
    sigma = rng.standard_normal(nf)
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('8i', N, nf, len(Mt[0]), len(Mp[0]), len(sel), len(mats), len(rows), m))
            for a in (psi, tclass, Mt[0], Mt[1], Mt[2], Mp[0], Mp[1], Mp[2], sel, mat_off, rows, cols, vals, fvec, k, sigma):
                f.write(np.ascontiguousarray(a).tobytes())
        subprocess.run([exe, fi, fo], check=True)
        raw = open(fo, 'rb').read()
    o = [0]

    def take(dtype, *shape):
        cnt = int(np.prod(shape)) * np.dtype(dtype).itemsize
        a = np.frombuffer(raw[o[0]:o[0] + cnt], dtype=dtype).reshape(shape)
        o[0] += cnt
        return a
    er, forms, proj = take(float, nf, n), take(complex, nf, len(sel), len(mats)), take(complex, nf, n)
    rank, psi_sel = take(np.int32, nf, n), take(complex, nf, n, n)
    q = np.where(tclass[None, :, None] == 1, -1j, 1) * psi[:, N:, :]          # conj(T) psi[N:]
    form = lambda t, qq: np.einsum('fij,ik,fkj->fj', qq.conj(), _coo_dense(t, N), qq)
    ref_er = np.abs(form(Mp, q)) / np.abs(form(Mt, q))
    assert np.abs(er - ref_er).max() <= 1e-12 * np.abs(ref_er).max()
    for mi, t in enumerate(mats):
        ref = form(t, q[:, :, sel])
        assert np.abs(forms[:, :, mi] - ref).max() <= 1e-12 * np.abs(ref).max()
    ref_proj = np.einsum('frj,r->fj', psi, fvec)
    assert np.abs(proj - ref_proj).max() <= 1e-12 * np.abs(ref_proj).max()
    for f in range(nf):
        d2 = (k[f].real - sigma[f]) ** 2 + k[f].imag ** 2
        ref_rank = np.empty(n, dtype=np.int32)
        ref_rank[np.lexsort((np.arange(n), d2))] = np.arange(n)
        assert np.array_equal(rank[f], ref_rank)
        keep = ref_rank < m
        assert np.array_equal(psi_sel[f][:, keep], psi[f][:, keep]) and not psi_sel[f][:, ~keep].any()


def _coo_dense(t, N):
    out = np.zeros([N, N], complex)
    np.add.at(out, (t[0], t[1]), t[2])
    return out
