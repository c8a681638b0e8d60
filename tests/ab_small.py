"""Timing of the eigenvalue stage on a small case (Muggleton submerged pipe, 2N = 68): AXS_HQR_SMALL_DIRECT=0/1."""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200'); sys.path.insert(0, 'tests')
import torch
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat
name = sys.argv[1] if len(sys.argv) > 1 else 'muggleton_submerged'
nf = 4000
c = getattr(cases, name)(nf=nf)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(c['n'])
ops = m._device_ops; N = ops.N
omega = torch.from_numpy(2 * np.pi * c['f']).cuda()
work = nat.EigWorkspace(N, nf)
L = nat.lib(); st = nat._stream(torch); P = nat._ptr
nat.check(L.axs_reduce(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d), P(omega), nf, None,
                       P(work.buf), work.bytes, st), 'reduce')
for var in ('0', '1'):
    nat.set_tuning('reduce_prebal', int(var))
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.check(L.axs_reduce(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d), P(omega), nf, None,
                               P(work.buf), work.bytes, st), 'reduce')
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print('%s 2N=%d prebal=%s: reduce %.2f ms per %d points' % (name, 2 * N, var, best, nf), flush=True)
res = {}
for var in ('0', '1'):
    nat.set_tuning('hqr_small_direct', int(var))
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); k, info = nat.hqr(N, nf, work); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    res[var] = np.sort_complex(k.cpu().numpy().reshape(nf, 2 * N), axis=1) if False else k.cpu().numpy().reshape(nf, 2 * N)
    print('%s 2N=%d direct=%s: hqr %.2f ms per %d points, failed %d' % (name, 2 * N, var, best, nf, int(info.sum())), flush=True)
a, b = res['0'], res['1']
print('max |dk|/|k| between the two (sorted per frequency): %.1e' % max(
    np.abs(np.sort_complex(a[i]) - np.sort_complex(b[i])).max() / np.abs(a[i]).max() for i in range(0, nf, 97)))
