"""Per-kernel timing of the large eigen path (not a pytest file).
usage: python tests/bench_big.py epl:order:nf [epl:order:nf ...]"""
import contextlib, io, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200')
import torch
import axisafe_b200 as ax
from axisafe_b200 import _native as nat, cases

for spec in sys.argv[1:] or ['3:10:296']:
    epl, order, nf = (int(x) for x in spec.split(':'))
    c = cases.coated_pipe(nf=nf, n=1, elements_per_layer=epl, order=order)
    m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    with contextlib.redirect_stdout(io.StringIO()):
        m.assemble_matrices(c['n'])
    ops, N = m._device_ops, m.no_of_dofs
    n2 = 2 * N
    om = torch.from_numpy(2 * np.pi * c['f']).cuda()
    work = nat.EigWorkspace(N, nf)
    k = torch.empty(nf * n2, dtype=torch.complex128, device='cuda')
    psi = torch.empty(nf * n2 * n2, dtype=torch.complex128, device='cuda')
    info = torch.zeros(nf, dtype=torch.int32, device='cuda')
    L = nat.lib()
    best = None
    for rep in range(int(__import__("os").environ.get("BIG_REPS", 2))):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda.synchronize()
        ev[0].record()
        nat.check(L.axs_reduce(N, ops.hb, nat._ptr(ops.K0s), nat._ptr(ops.Cb), nat._ptr(ops.Mb), nat._ptr(ops.K1c), nat._ptr(ops.K6d),
                               nat._ptr(om), nf, None, nat._ptr(work.buf), work.bytes, nat._stream(torch)), 'reduce')
        ev[1].record()
        nat.check(L.axs_hqr(N, nf, nat._ptr(k), nat._ptr(info), nat._ptr(work.buf), work.bytes, nat._stream(torch)), 'hqr')
        ev[2].record()
        nat.check(L.axs_modes(N, ops.hb, nat._ptr(ops.Cb), nat._ptr(ops.K6d), nf, nat._ptr(k), nat._ptr(psi), nat._ptr(work.buf),
                              work.bytes, nat._stream(torch)), 'modes')
        ev[3].record()
        arg, amax = nat.track_pairs(ops, psi, nf - 1)
        ev[4].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        best = t if best is None or sum(t) < sum(best) else best
    tot = sum(best)
    print('2N=%d nf=%d ms: reduce %.1f hqr %.1f modes %.1f track %.1f total %.1f -> %.1f points/s, %.2f TFLOP/s (100 n^3)  info=%d'
          % (n2, nf, best[0], best[1], best[2], best[3], tot, nf / tot * 1e3, 100.0 * n2 ** 3 * nf / tot / 1e9,
             int(info.cpu().numpy().sum())), flush=True)
    del work, psi
    torch.cuda.empty_cache()
