import contextlib, io, os, sys, ctypes
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import torch
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat
nf = 148
c = cases.pipe_soil_pml(nf=nf)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
ops = m._device_ops; N = ops.N
omega = torch.from_numpy(2 * np.pi * c['f']).cuda()
work = nat.EigWorkspace(N, nf)
L = nat.lib(); st = nat._stream(torch); P = nat._ptr
nat.check(L.axs_reduce(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d), P(omega), nf, None, P(work.buf), work.bytes, st), 'reduce')
os.environ['AXS_HQR_BULGES'] = sys.argv[1] if len(sys.argv) > 1 else '0'
k, info = nat.hqr(N, nf, work); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
L.axs_debug_counters(buf)
tot, pro, left, right, nsw, nrot = [buf[i] for i in range(6)]; print('defl', buf[6], 'load', buf[7])
print('block0: total %d cyc; prologue %d (%.0f/sweep); left %d (%.0f/rot); right %d (%.0f/rot); sweeps %d rots %d' % (tot, pro, pro / nsw, left, left / nrot, right, right / nrot, nsw, nrot))
