"""A/B of the aggressive-early-deflation settings of the QR cascade on the flagship matrices (GPU).

    python tests/ab_aed.py [nf] [aed:reg:stages ...] > gpurun_out/ab_aed.log
    AXS_LIBRARY=<path of another build of the library> python tests/ab_aed.py ...

One process, one reduction of nf cfg1 matrices, then for every setting (hqr_aed window, register /
shared-memory window routine, per-stage mask): best-of-3 time of axs_hqr, the axs_hqr_stats counters
(macro steps, window calls, kilo-cycles per stage) and the eigenvalue error of 5 frequencies against
scipy.linalg.eig (bar 1e-8).  The last line is JSON: the fastest setting that met the bar, for the caller.
"""
import contextlib, io, json, os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'axisym-safe-python_b200'); sys.path.insert(0, 'tests')
import torch
import axisafe_b200 as ax
from axisafe_b200 import cases, _native as nat
from oracle import safe_oracle as so

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
full = cases.pipe_soil_pml(nf=2000)
c = dict(full, f=full['f'][::2000 // nf][:nf])               # every other point of the north-star sweep
m = ax.mesh.Mesh(); m.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
with contextlib.redirect_stdout(io.StringIO()):
    m.assemble_matrices(0)
ops = m._device_ops; N = ops.N; n2 = 2 * N
omega = torch.from_numpy(2 * np.pi * np.ascontiguousarray(c['f'])).cuda()
work = nat.EigWorkspace(N, nf)
L = nat.lib(); st = nat._stream(torch); P = nat._ptr
nat.check(L.axs_reduce(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d), P(omega), nf, None,
                       P(work.buf), work.bytes, st), 'reduce')
G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), 0); op = so.operators(G, 0)
ref = {i: so.eig_normalised(op, 2 * np.pi * c['f'][i])[0] for i in (0, nf // 4, nf // 2, (3 * nf) // 4, nf - 1)}

SETTINGS = [(0, 0, 15), (8, 0, 15), (7, 0, 15), (6, 0, 15), (5, 0, 15), (4, 0, 15),
            (6, 0, 13), (6, 0, 9), (6, 0, 5), (6, 0, 1), (6, 0, 7), (6, 0, 11), (6, 0, 3),
            (8, 0, 13), (8, 0, 9), (8, 0, 5), (8, 0, 1), (7, 0, 13), (5, 0, 13), (5, 0, 9)]
if len(sys.argv) > 2:
    SETTINGS = [tuple(int(x) for x in a.split(':')) for a in sys.argv[2:]]
rows = []
for aed, reg, mask in SETTINGS:
    nat.set_tuning('hqr_aed', aed); nat.set_tuning('hqr_aed_reg', reg); nat.set_tuning('hqr_aed_stages', mask)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); k, info = nat.hqr(N, nf, work); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    kk = k.cpu().numpy().reshape(nf, n2)
    err = max(so.match_eigenvalues(ref[i], kk[i])[1] for i in ref)
    failed = int(info.sum())
    s = nat.hqr_stats(N, nf, work).astype(float).mean(axis=0)
    rows.append({'aed': aed, 'reg': reg, 'stages': mask, 'ms_per_2000': best * 2000 / nf, 'err': float(err), 'failed': failed,
                 'macro': s[0], 'batches': s[1], 'calls': s[2], 'deflated': s[3],
                 'kcycles_window_apply_sweeps': [[round(s[4 + 3 * g + q]) for q in range(3)] for g in range(4)]})
    r = rows[-1]
    print('aed=%d reg=%d stages=%2d  %.2f ms per 2000  err %.1e failed %d | macro %.0f batches %.0f calls %.1f deflated %.0f | kcycles %s'
          % (aed, reg, mask, r['ms_per_2000'], err, failed, s[0], s[1], s[2], s[3], r['kcycles_window_apply_sweeps']), flush=True)
good = [r for r in rows if r['err'] < 1e-8 and r['failed'] == 0]
best = min(good, key=lambda r: r['ms_per_2000'])
base = rows[0]
print(json.dumps({'baseline_ms': base['ms_per_2000'], 'best': best, 'gain': 1.0 - best['ms_per_2000'] / base['ms_per_2000']}))
