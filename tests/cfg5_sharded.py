"""torchrun target (not a pytest file): BASELINE.json configs[4] - the ten-layer waveguide, 2N = 7806,
frequencies of the 10 000-point sweep sharded over the ranks.  A bounded sample: ``per`` frequencies per
rank (default 1), every rank solving its own block on its GPU (no data-path collective), checked through
sum(k) = tr S and, on the rank that owns 20 kHz... (the zgeev fixture frequency is solved by rank 0 in
addition).  Prints one line per rank and a summary with the aggregate rate."""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import torch
import torch.distributed as dist

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat
from oracle import safe_oracle as so

rank, world = dist.get_rank(), dist.get_world_size()
per = int(sys.argv[1]) if len(sys.argv) > 1 else 1
c = cases.ten_layer(nf=10000, n=1)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(1)
n2 = 2 * m.no_of_dofs
block = 10000 // world
f = c['f'][rank * block:rank * block + per].copy()
if rank == 0:
    f[0] = 20e3                                   # the frequency of tests/golden/big_cfg5.npz
sol = ax.solver.WaveElementAxisym(m)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    sol.solve(f, distributed=False)
torch.cuda.synchronize(); secs = time.perf_counter() - t0
ops = m._device_ops
trS = -np.sum(ops.Cb.cpu().numpy().reshape(ops.N, -1)[:, ops.hb] / ops.K6d.cpu().numpy())
tr_err = float(np.abs(sol.k.sum(axis=1) - trS).max() / np.abs(sol.k).sum(axis=1).max())
line = 'rank %d: %d frequencies of 2N = %d in %.1f s, info %d, trace identity %.1e' % (rank, per, n2, secs, int(sol.info.sum()), tr_err)
ok = tr_err < 1e-9 and not sol.info.any()
if rank == 0:
    g = np.load('tests/golden/big_cfg5.npz')
    perm, worst = so.match_eigenvalues_large(g['eig_untracked'][0], sol.k_native[0])
    err = np.abs(g['eig_untracked'][0] - sol.k_native[0][perm])
    tol = so.eig_tolerance_from_kappa(g['eig_untracked'][0], g['kappa'][0], float(g['norm2_balanced'][0]))
    ok = ok and bool(np.all(err <= tol))
    line += '; vs zgeev at 20 kHz: max rel err %.2e, within the bar: %s' % (worst, bool(np.all(err <= tol)))
t = torch.tensor([secs, float(ok)], dtype=torch.float64, device='cuda')
tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
lines = [None] * world
dist.all_gather_object(lines, line)
if rank == 0:
    print('\n'.join(lines))
    print('CFG5_SHARDED %s: %d ranks x %d frequencies in %.1f s (max over ranks) = %.3f points/s aggregate' %
          ('PASS' if tmin[1].item() > 0 else 'FAIL', world, per, tmax[0].item(), world * per / tmax[0].item()))
dist.destroy_process_group()
