"""A/B timing + accuracy of the Hessenberg-reduction variants (AXS_REDUCE_OC env: 0 = L2-resident
k_reduce, 8/12/16 = on-chip k_reduce_oc with that many warps).  python tests/ab_reduce.py [nf] [variants...]"""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200'); sys.path.insert(0, 'tests')
import torch
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat
from oracle import safe_oracle as so
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
c = cases.pipe_soil_pml(nf=nf)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
ops = m._device_ops; N = ops.N
omega = torch.from_numpy(2 * np.pi * c['f']).cuda()
work = nat.EigWorkspace(N, nf)
L = nat.lib(); st = nat._stream(torch); P = nat._ptr
G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), 0); op = so.operators(G, 0)
ref = {i: so.eig_normalised(op, 2 * np.pi * c['f'][i])[0] for i in (0, nf // 2, nf - 1)}
for var in sys.argv[2:] or ['0', '8', '12', '16']:
    nat.set_tuning('reduce_oc', int(var))
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.check(L.axs_reduce(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d), P(omega), nf, None,
                               P(work.buf), work.bytes, st), 'reduce')
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    k, info = nat.hqr(N, nf, work)
    kk = k.cpu().numpy().reshape(nf, 2 * N)
    err = max(so.match_eigenvalues(ref[i], kk[i])[1] for i in ref)
    print('reduce variant %s: %.2f ms for %d matrices (%.2f ms per 2000)  max eig err %.1e  failed %d'
          % (var, best, nf, best * 2000 / nf, err, int(info.sum())), flush=True)
