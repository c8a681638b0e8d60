"""A/B timing + agreement of the mode-recovery variants (AXS_BACKNORM_WY = 0 slab / 1 blocked WY on DMMA).
python tests/ab_modes.py [nf]"""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200'); sys.path.insert(0, 'tests')
import torch
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
c = cases.pipe_soil_pml(nf=nf)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
ops = m._device_ops; N = ops.N
omega = torch.from_numpy(2 * np.pi * c['f']).cuda()
work = nat.EigWorkspace(N, nf)
L = nat.lib(); st = nat._stream(torch); P = nat._ptr
nat.check(L.axs_reduce(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d), P(omega), nf, None,
                       P(work.buf), work.bytes, st), 'reduce')
k, info = nat.hqr(N, nf, work)
out = {}
for var in sys.argv[2:] or ['0', '1']:
    nat.set_tuning('backnorm_wy', int(var))
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); psi = nat.modes(ops, nf, k, work); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[var] = psi.cpu().numpy().reshape(nf, 2 * N, 2 * N)
    print('modes variant %s: %.2f ms per %d' % (var, best, nf), flush=True)
if len(out) == 2:
    a, b = out['0'], out['1']
    # columns are defined up to sign (principal sqrt of the normaliser)
    num = np.minimum(np.abs(a - b).max(axis=1), np.abs(a + b).max(axis=1)); den = np.abs(a).max(axis=1)
    rel = num / den
    print('max relative column difference %.2e (median %.1e), nan %d' % (np.nanmax(rel), np.nanmedian(rel), int(np.isnan(b).sum())))
