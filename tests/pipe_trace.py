"""Timeline of the pipelined solve (not a pytest file): AXS_PIPE_TRACE=1 python tests/pipe_trace.py"""
import contextlib, io, os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import numpy as np, torch
import axisafe_b200 as ax
from axisafe_b200 import cases
c = cases.pipe_soil_pml(nf=2000)
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(c['n'])
sol = ax.solver.WaveElementAxisym(m)
for K in (sys.argv[1:] or ['4']):
    if ',' in K:
        os.environ['AXS_PIPE_BOUNDS'] = K
    else:
        os.environ.pop('AXS_PIPE_BOUNDS', None)
        os.environ['AXS_PIPE_CHUNKS'] = K
    for rep in range(4):
        torch.cuda.synchronize(); t = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            sol.solve(c['f'])
        _ = sol.psi_hat
        torch.cuda.synchronize()
        print('K=%s rep %d: %.1f ms' % (K, rep, (time.perf_counter() - t) * 1e3), file=sys.stderr)
        sol.psi_hat = None
