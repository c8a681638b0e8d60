"""torchrun target: sharded solve() must reproduce the single-GPU tracked k bit for bit.

Three sharded variants against every rank solving the whole sweep alone:
  * nf = 301  - interleaved chunk pipeline with 2 chunks (solver._solve_sharded_pipelined),
  * nf = 1300 - the same with 4 chunks (>= 512 points per rank at 2 ranks),
  * nf = 301 with AXS_SHARD_PIPE=0 - one contiguous block per rank (solver._track_sharded),
and the downstream energy ratio computed from the sharded mode shapes.
"""
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

world = dist.get_world_size()
all_ok = True
report = []
for nf, pipe in ((301, '1'), (650 * world, '1'), (301, '0')):
    os.environ['AXS_SHARD_PIPE'] = pipe
    c = cases.pipe_soil_pml(nf=nf)
    m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    with contextlib.redirect_stdout(io.StringIO()):
        m.assemble_matrices(0)
        a = ax.solver.WaveElementAxisym(m); a.solve(c['f'])                       # sharded
        b = ax.solver.WaveElementAxisym(m); b.solve(c['f'], distributed=False)    # every rank alone
        a.energy_ratio(); b.energy_ratio()
        a.k_propagating(15, 0.935); b.k_propagating(15, 0.935)
        f_ext = np.zeros(m.no_of_dofs); f_ext[m.dof_domains['solid'][0]] = 1.0
        frf_a = a.calculate_frf(0.01, 0.05, f_ext, distance=0.3)
        frf_b = b.calculate_frf(0.01, 0.05, f_ext, distance=0.3)
    loc = a.local_index
    ok = np.array_equal(a.k, b.k) and np.array_equal(a.order, b.order) and np.array_equal(a.k_native, b.k_native)
    ok = ok and np.array_equal(a.info, b.info)
    psi_ok = np.array_equal(a.psi_hat, b.psi_hat[loc])         # sharded: this rank's frequencies only
    # downstream results are complete on every rank (er / FRF rows of the other ranks are exchanged)
    er_ok = np.array_equal(a.er, b.er, equal_nan=True) and len(loc) < nf
    prop_ok = np.array_equal(a.propagating_indices, b.propagating_indices) and np.array_equal(a.k_ready, b.k_ready)
    frf_ok = frf_a.shape == frf_b.shape and np.array_equal(frf_a, frf_b, equal_nan=True)
    contiguous = a.local_range is not None
    report.append('rank %d nf=%d pipe=%s k/order/info=%s psi=%s er=%s k_propagating=%s frf=%s contiguous=%s'
                  % (dist.get_rank(), nf, pipe, ok, psi_ok, er_ok, prop_ok, frf_ok, contiguous))
    all_ok = all_ok and ok and psi_ok and er_ok and prop_ok and frf_ok and (contiguous == (pipe == '0' or nf // world < 128))
# a sweep with fewer than 2 points per rank is solved by every rank alone (no empty blocks, no hang)
c = cases.pipe_soil_pml(nf=2 * world - 1)
with contextlib.redirect_stdout(io.StringIO()):
    tiny = ax.solver.WaveElementAxisym(m); tiny.solve(c['f'])
report.append('rank %d tiny sweep of %d points: local_index=%s' % (dist.get_rank(), len(tiny.w), list(tiny.local_index)))
all_ok = all_ok and len(tiny.local_index) == len(tiny.w)
t = torch.tensor([int(all_ok)], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MIN)
lines = [None] * world
dist.all_gather_object(lines, report)
if dist.get_rank() == 0:
    for r in range(world):
        print('\n'.join(lines[r]))
    print('MGPU_CHECK', 'PASS' if int(t.item()) else 'FAIL', 'ranks', world)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
