"""Diagnostic (not a pytest file): cfg4 at full size under the kernel-selection switches of the large path."""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import torch
import axisafe_b200 as ax
from axisafe_b200 import _native as nat, cases
from oracle import safe_oracle as so
g = np.load('tests/golden/big_cfg4.npz')
c = cases.coated_pipe(nf=2, n=1)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(1)
ops, N = m._device_ops, m.no_of_dofs
n2 = 2 * N
w = torch.from_numpy(2 * np.pi * g['f']).cuda()
combos = [dict(big_blocked=1, big_qr_cluster=0), dict(big_blocked=1, big_qr_cluster=1), dict(big_blocked=0, big_qr_cluster=0)]
if len(sys.argv) > 1:
    combos = combos[:int(sys.argv[1])]
for combo in combos:
    for k_, v in combo.items():
        nat.set_tuning(k_, v)
    for rep in range(2 if combo['big_qr_cluster'] == 0 else 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        work = nat.EigWorkspace(N, 2)
        L = nat.lib()
        nat.check(L.axs_reduce(N, ops.hb, nat._ptr(ops.K0s), nat._ptr(ops.Cb), nat._ptr(ops.Mb), nat._ptr(ops.K1c), nat._ptr(ops.K6d),
                               nat._ptr(w), 2, None, nat._ptr(work.buf), work.bytes, nat._stream(torch)), 'reduce')
        torch.cuda.synchronize(); t1 = time.perf_counter()
        k, info = nat.hqr(N, 2, work)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        k = k.cpu().numpy().reshape(2, n2)
        out = []
        for j in range(2):
            ref = g['eig_untracked'][j]
            perm, worst = so.match_eigenvalues_large(ref, k[j])
            rel = np.abs(ref - k[j][perm]) / np.abs(ref)
            out.append('f%d: worst %.2e, >1e-8: %d, >1e-6: %d, trace err %.1e' % (j, worst, (rel > 1e-8).sum(), (rel > 1e-6).sum(),
                       abs(k[j].sum() - g['trS'][j]) / np.abs(k[j]).sum()))
        print(combo, 'rep', rep, 'reduce %.2f s hqr %.2f s info %s |' % (t1 - t0, t2 - t1, info.cpu().numpy().tolist()), ' | '.join(out), flush=True)
        del work
