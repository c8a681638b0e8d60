"""Stage-by-stage GPU diagnostic (not a pytest file): prints error norms per stage."""
import contextlib, io, sys, time, traceback
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200'); sys.path.insert(0, 'tests')
import torch
from conftest import GOLDEN_CASES, load_golden, golden_dense, make_case
from oracle import safe_oracle as so
import axisafe_b200 as ax
from axisafe_b200 import _native as nat

def stage(name):
    def deco(fn):
        def run(*a, **k):
            try:
                t = time.time(); out = fn(*a, **k); torch.cuda.synchronize()
                print('[ok  ] %-28s %.3fs %s' % (name, time.time() - t, out if out is not None else ''), flush=True)
            except Exception:
                print('[FAIL] %-28s' % name); traceback.print_exc(); sys.stdout.flush()
        return run
    return deco

which = sys.argv[1:] or ['pipe_soil_pml', 'free_steel_pipe', 'gravenkamp', 'nguyen']
for name, n, kw in GOLDEN_CASES:
    if name not in which: continue
    print('=====', name, n)
    g = load_golden(name, n); c = make_case(name, n, kw)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), n)
    op = so.operators(G, n); N = op['N']; n2 = 2 * N
    m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    ctx = {}
    @stage('assemble')
    def s1():
        with contextlib.redirect_stdout(io.StringIO()):
            m.assemble_matrices(n)
        errs = {}
        for key in ('K1', 'K2', 'K3', 'K4', 'K5', 'K6', 'M', 'KF', 'KFt', 'KFz'):
            ref = golden_dense(g, key); got = getattr(m, key).toarray()
            errs[key] = '%.1e' % (np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))
        ops = m._device_ops
        for dev, key in ((ops.K0s, 'K0s'), (ops.Cb, 'C'), (ops.Mb, 'M')):
            errs['dev' + key] = '%.1e' % (np.abs(ops.band_to_dense(dev) - op[key]).max() / np.abs(op[key]).max())
        errs['K6d'] = '%.1e' % (np.abs(ops.K6d.cpu().numpy() - np.diag(op['K6'])).max() / np.abs(op['K6']).max())
        errs['hb'] = ops.hb
        return errs
    s1()
    pick = g['pick']; w = 2 * np.pi * c['f'][pick]
    @stage('build_S')
    def s2():
        S = nat.build_S(m._device_ops, torch.from_numpy(w).cuda()).cpu().numpy().reshape(len(w), n2, n2).transpose(0, 2, 1)
        return ['%.1e' % (np.abs(S[j] - g['S'][j]).max() / np.abs(g['S'][j]).max()) for j in range(len(w))]
    s2()
    @stage('reduce')
    def s3():
        Sd = torch.from_numpy(np.ascontiguousarray(g['S'].transpose(0, 2, 1))).cuda().reshape(-1)
        work = nat.reduce_dense(Sd, N, len(pick)); ctx['work'] = work
        scale, H = nat.get_reduction(N, len(pick), work)
        out = []
        for j in range(len(pick)):
            ev = np.linalg.eigvals(H[j]); perm, e = so.match_eigenvalues(g['eig_untracked'][j], ev)
            out.append('eigerr %.1e tril %.1e log2d[%d,%d]' % (e, np.abs(np.tril(H[j], -2)).max(), np.log2(scale[j]).min(), np.log2(scale[j]).max()))
        return out
    s3()
    @stage('hqr')
    def s4():
        k, info = nat.hqr(N, len(pick), ctx['work'])
        k = k.cpu().numpy().reshape(len(pick), n2); ctx['k'] = k
        return ['%.1e' % so.match_eigenvalues(g['eig_untracked'][j], k[j])[1] for j in range(len(pick))], info.cpu().numpy()
    s4()
    @stage('eig_sweep+modes')
    def s5():
        k, psi, info, work = nat.eig_sweep(m._device_ops, torch.from_numpy(w).cuda())
        k = k.cpu().numpy().reshape(len(pick), n2); psi = psi.cpu().numpy().reshape(len(pick), n2, n2)
        out = []
        for j, i in enumerate(pick):
            val, vec = so.eig_normalised(op, w[j])
            perm, e = so.match_eigenvalues(val, k[j])
            err = so.mode_shape_error(vec[N:], psi[j][N:, perm])
            gram = psi[j].T @ op['B'] @ psi[j]
            out.append('eig %.1e modes max %.1e med %.1e n>1e-6 %d gram %.1e' % (e, err.max(), np.median(err), (err > 1e-6).sum(), np.abs(np.diag(gram) - 1).max()))
        return out
    s5()
    @stage('solve api')
    def s6():
        sol = ax.solver.WaveElementAxisym(m)
        with contextlib.redirect_stdout(io.StringIO()):
            t = time.time(); sol.solve(c['f']); torch.cuda.synchronize(); dt = time.time() - t
        sys.path.insert(0, 'tests')
        from test_oracle_golden import tracked_agreement
        cover, frac = tracked_agreement(g['k'], sol.k)
        return 'cover %.1e frac %.3f  solve %.3fs for %d freqs' % (cover, frac, dt, len(c['f']))
    s6()

# timing of the flagship at 2000 points, per stage
c = make_case('pipe_soil_pml', 0, {'nf': 2000})
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
ops = m._device_ops; N = ops.N; nf = 2000
omega = torch.from_numpy(2 * np.pi * c['f']).cuda()
work = nat.EigWorkspace(N, nf)
def timed(label, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print('%-14s %8.2f ms' % (label, best), flush=True); return out
L = nat.lib(); st = nat._stream(torch)
timed('reduce', lambda: nat.check(L.axs_reduce(N, ops.hb, nat._ptr(ops.K0s), nat._ptr(ops.Cb), nat._ptr(ops.Mb), nat._ptr(ops.K1c), nat._ptr(ops.K6d), nat._ptr(omega), nf, None, nat._ptr(work.buf), work.bytes, st), 'reduce'))
k, info = timed('hqr', lambda: nat.hqr(N, nf, work))
psi = timed('modes', lambda: nat.modes(ops, nf, k, work))
timed('track', lambda: nat.track_pairs(ops, psi, nf - 1))
sol = ax.solver.WaveElementAxisym(m)
for _ in range(2):
    with contextlib.redirect_stdout(io.StringIO()):
        t = time.time(); sol.solve(c['f']); torch.cuda.synchronize(); print('solve() wall %.1f ms' % ((time.time() - t) * 1e3), file=sys.stderr)
