"""Short profiling driver (ncu target): one device-resident sweep of the flagship case."""
import contextlib, io, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import torch
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = cases.pipe_soil_pml(nf=nf)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
ops = m._device_ops
omega = torch.from_numpy(2 * np.pi * c['f']).cuda()
work = nat.EigWorkspace(ops.N, nf)
for _ in range(reps):
    k, psi, info, work = nat.eig_sweep(ops, omega, work=work)
    arg, amax = nat.track_pairs(ops, psi, nf - 1, scratch=work.buf)
torch.cuda.synchronize()
print('done', nf, int(info.max()))
