"""torchrun target: sharded solve() must reproduce the single-GPU tracked k bit for bit."""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import torch
import torch.distributed as dist

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import axisafe_b200 as ax
from axisafe_b200 import cases
c = cases.pipe_soil_pml(nf=301)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
    a = ax.solver.WaveElementAxisym(m); a.solve(c['f'])                       # sharded
    b = ax.solver.WaveElementAxisym(m); b.solve(c['f'], distributed=False)    # every rank alone
ok = np.array_equal(a.k, b.k) and np.array_equal(a.order, b.order)
lo, hi = a.local_range
psi_ok = np.array_equal(a.psi_hat, b.psi_hat[lo:hi])       # sharded: local block only
t = torch.tensor([int(ok and psi_ok)], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0:
    print('MGPU_CHECK', 'PASS' if int(t.item()) else 'FAIL', 'ranks', dist.get_world_size(), 'block', a.local_range)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
