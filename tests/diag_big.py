"""Stage-by-stage GPU diagnostic of the large (global-memory) eigen path; not a pytest file."""
import contextlib, io, sys, time, traceback
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200'); sys.path.insert(0, 'tests')
import torch
from oracle import safe_oracle as so
import axisafe_b200 as ax
from axisafe_b200 import _native as nat, cases


def stage(name):
    def deco(fn):
        try:
            t = time.time(); out = fn(); torch.cuda.synchronize()
            print('[ok  ] %-26s %.3fs %s' % (name, time.time() - t, out if out is not None else ''), flush=True)
        except Exception:
            print('[FAIL] %-26s' % name); traceback.print_exc(); sys.stdout.flush()
        return fn
    return deco


def ev_err(ref, got):
    perm, e = so.match_eigenvalues(ref, got)
    return e


sizes = [int(a) for a in sys.argv[1:]] or [100, 150, 300]
rng = np.random.default_rng(1)
for N in sizes:
    n = 2 * N
    nf = 3
    print('===== random dense, 2N =', n)
    S = (rng.standard_normal((nf, n, n)) + 1j * rng.standard_normal((nf, n, n)))
    S *= np.exp(rng.uniform(-3, 3, (nf, n, 1))) * np.exp(rng.uniform(-3, 3, (nf, 1, n)))   # badly scaled
    ref = [np.linalg.eigvals(S[j]) for j in range(nf)]
    ctx = {}

    @stage('reduce')
    def _():
        Sd = torch.from_numpy(np.ascontiguousarray(S.transpose(0, 2, 1))).cuda().reshape(-1)
        ctx['work'] = nat.reduce_dense(Sd, N, nf)
        scale, H = nat.get_reduction(N, nf, ctx['work'])
        ctx['H'], ctx['scale'] = H, scale
        out = []
        for j in range(nf):
            out.append('eig %.1e tril %.1e log2d[%d,%d]' % (ev_err(ref[j], np.linalg.eigvals(H[j])), np.abs(np.tril(H[j], -2)).max(),
                                                        np.log2(scale[j]).min(), np.log2(scale[j]).max()))
        return out

    @stage('hqr')
    def _():
        k, info = nat.hqr(N, nf, ctx['work'])
        ctx['k'] = k
        kh = k.cpu().numpy().reshape(nf, n)
        return ['%.1e' % ev_err(ref[j], kh[j]) for j in range(nf)], info.cpu().numpy().tolist()

# ---- the SAFE case: coated pipe with a mid-size mesh ------------------------------------------
for epl, order in ((2, 8), (3, 10)):
    c = cases.coated_pipe(nf=5, n=1, elements_per_layer=epl, order=order)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n']); N = op['N']; n2 = 2 * N
    print('===== coated_pipe epl=%d order=%d 2N=%d' % (epl, order, n2))
    m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    with contextlib.redirect_stdout(io.StringIO()):
        m.assemble_matrices(c['n'])
    ops = m._device_ops
    w = 2 * np.pi * c['f']
    omega = torch.from_numpy(w).cuda()
    refs = [so.eig_normalised(op, wi) for wi in w]
    ctx = {}

    @stage('reduce(ops)')
    def _():
        work = nat.EigWorkspace(N, len(w))
        nat.check(nat.lib().axs_reduce(N, ops.hb, nat._ptr(ops.K0s), nat._ptr(ops.Cb), nat._ptr(ops.Mb), nat._ptr(ops.K1c),
                                       nat._ptr(ops.K6d), nat._ptr(omega), len(w), None, nat._ptr(work.buf), work.bytes,
                                       nat._stream(torch)), 'reduce')
        ctx['work'] = work
        scale, H = nat.get_reduction(N, len(w), work)
        return ['%.1e' % ev_err(refs[j][0], np.linalg.eigvals(H[j])) for j in range(len(w))]

    @stage('hqr')
    def _():
        k, info = nat.hqr(N, len(w), ctx['work'])
        ctx['k'] = k
        kh = k.cpu().numpy().reshape(len(w), n2)
        tol = []
        for j in range(len(w)):
            perm, e = so.match_eigenvalues(refs[j][0], kh[j])
            S = so.S_matrix(op, w[j])
            ok = np.all(np.abs(refs[j][0] - kh[j][perm]) <= so.eig_tolerance(S, refs[j][0]))
            tol.append('%.1e/%s' % (e, ok))
        return tol, info.cpu().numpy().tolist()

    @stage('modes')
    def _():
        psi = nat.modes(ops, len(w), ctx['k'], ctx['work'])
        ctx['psi'] = psi
        ph = psi.cpu().numpy().reshape(len(w), n2, n2)
        kh = ctx['k'].cpu().numpy().reshape(len(w), n2)
        out = []
        for j in range(len(w)):
            S = so.S_matrix(op, w[j])
            res = np.linalg.norm(S @ ph[j] - ph[j] * kh[j], axis=0) / (np.linalg.norm(ph[j], axis=0) * np.abs(kh[j]))
            gram = np.einsum('ij,ij->j', ph[j], op['B'] @ ph[j])
            out.append('res med %.1e max %.1e gram %.1e' % (np.median(res), res.max(), np.abs(gram - 1).max()))
        return out

    @stage('track')
    def _():
        arg, amax = nat.track_pairs(ops, ctx['psi'], len(w) - 1)
        arg = arg.cpu().numpy().reshape(len(w) - 1, n2); amax = amax.cpu().numpy().reshape(len(w) - 1, n2)
        ph = ctx['psi'].cpu().numpy().reshape(len(w), n2, n2)
        bad = 0
        for p in range(len(w) - 1):
            P = np.abs(ph[p].T @ op['B'] @ ph[p + 1])
            bad += int(np.sum(arg[p] != P.argmax(axis=1)))
            assert np.abs(amax[p] - P.max(axis=1)).max() < 1e-8 * P.max()
        return 'argmax mismatches %d' % bad

    @stage('solve API')
    def _():
        sol = ax.solver.WaveElementAxisym(m)
        with contextlib.redirect_stdout(io.StringIO()):
            sol.solve(c['f'])
        e = [so.match_eigenvalues(refs[j][0], sol.k_native[j])[1] for j in range(len(w))]
        psi = sol.psi_hat
        j = 2
        S = so.S_matrix(op, w[j])
        res = np.linalg.norm(S @ psi[j] - psi[j] * sol.k[j], axis=0) / (np.linalg.norm(psi[j], axis=0) * np.abs(sol.k[j]))
        return 'eig %.1e tracked-res med %.1e' % (max(e), np.median(res))

    @stage('timing nf=64')
    def _():
        om = torch.from_numpy(2 * np.pi * np.linspace(1e3, 50e3, 64)).cuda()
        work = nat.EigWorkspace(N, 64)
        k = torch.empty(64 * n2, dtype=torch.complex128, device='cuda')
        psi = torch.empty(64 * n2 * n2, dtype=torch.complex128, device='cuda')
        info = torch.zeros(64, dtype=torch.int32, device='cuda')
        out = []
        L = nat.lib()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        ev[0].record()
        nat.check(L.axs_reduce(N, ops.hb, nat._ptr(ops.K0s), nat._ptr(ops.Cb), nat._ptr(ops.Mb), nat._ptr(ops.K1c), nat._ptr(ops.K6d),
                               nat._ptr(om), 64, None, nat._ptr(work.buf), work.bytes, nat._stream(torch)), 'reduce')
        ev[1].record()
        nat.check(L.axs_hqr(N, 64, nat._ptr(k), nat._ptr(info), nat._ptr(work.buf), work.bytes, nat._stream(torch)), 'hqr')
        ev[2].record()
        nat.check(L.axs_modes(N, ops.hb, nat._ptr(ops.Cb), nat._ptr(ops.K6d), 64, nat._ptr(k), nat._ptr(psi), nat._ptr(work.buf),
                              work.bytes, nat._stream(torch)), 'modes')
        ev[3].record()
        torch.cuda.synchronize()
        return 'ms reduce %.1f hqr %.1f modes %.1f' % tuple(ev[i].elapsed_time(ev[i + 1]) for i in range(3))
