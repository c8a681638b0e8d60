"""A/B: whole device pipeline on 1 stream vs frequency chunks on several streams."""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import torch
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat
nf = 2000
c = cases.pipe_soil_pml(nf=nf)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
ops = m._device_ops; N = ops.N; n2 = 2 * N
omega = torch.from_numpy(2 * np.pi * c['f']).cuda()
L = nat.lib(); P = nat._ptr
def run(nchunks):
    bounds = [nf * i // nchunks for i in range(nchunks + 1)]
    streams = [torch.cuda.Stream() for _ in range(nchunks)]
    works = [nat.EigWorkspace(N, bounds[i + 1] - bounds[i]) for i in range(nchunks)]
    k = torch.empty(nf * n2, dtype=torch.complex128, device='cuda')
    psi = torch.empty(nf * n2 * n2, dtype=torch.complex128, device='cuda')
    info = torch.zeros(nf, dtype=torch.int32, device='cuda')
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i, s in enumerate(streams):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                lo, hi = bounds[i], bounds[i + 1]
                st = nat._stream(torch)
                nat.check(L.axs_eig_sweep(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d),
                                          P(omega[lo:hi]), hi - lo, P(k[lo * n2:hi * n2]), P(psi[lo * n2 * n2:hi * n2 * n2]),
                                          P(info[lo:hi]), P(works[i].buf), works[i].bytes, st), 'sweep')
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for nch in [int(x) for x in (sys.argv[1:] or '1 2 3 4 6 8'.split())]:
    print('chunks/streams %d: eig_sweep %.2f ms' % (nch, run(nch)), flush=True)
