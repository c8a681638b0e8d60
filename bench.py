#!/usr/bin/env python
"""Benchmark of the per-frequency SAFE dispersion sweep (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU algorithm on host cores

A "step" is ONE full dispersion sweep of the flagship workload - the water-filled cast-iron
pipe buried in soil with a PML (examples/water_filled_pipe_in_soil.py, 2N = 126) at the
north-star resolution of 2000 frequency points - per GPU (weak scaling: every rank sweeps its
own 2000-point block of one N*2000-point frequency axis).

JSON keys follow the driver contract; see DESIGN.md section "Measurement" for what each timed
region contains.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'axisym-safe-python_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

POINTS_PER_GPU = 2000
WORKLOAD = 'pipe_soil_pml (water-filled cast-iron pipe in soil + PML, n=0, 2N=126), 2000 frequency points per GPU'
METRIC = 'frequency points/sec (SAFE sweep)'
UNIT = 'freq_points/s'


def f_alg(n2):
    """Nominal real FP64 flops per frequency point: 100 n2^3 for the eigen-solve
    (zgehrd 10/3 + zunghr 4/3 + zhseqr ~20, complex -> x4); the QR kernel's share is 80 n2^3."""
    return 100.0 * n2 ** 3


# ------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """nvidia-smi clock / throttle sampling during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index, enabled=True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:                 # N > 1: rank 0 samples its GPU; the other ranks keep their host quiet
            return self
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------- CPU baseline
def _cpu_worker(args):
    """Reference loop body (solver.py:119-182 restated in oracle/safe_oracle.py) on a chunk."""
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    freqs, = args
    from oracle import safe_oracle as so
    from axisafe_b200 import cases
    c = cases.pipe_soil_pml(nf=POINTS_PER_GPU)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    t0 = time.perf_counter()
    so.solve(G, c['n'], freqs)
    return time.perf_counter() - t0


def cpu_baseline(points_per_core=100, as_shipped=False):
    """Oracle (port of the reference algorithm: S build + zgeev + normalisation + tracking) on
    all host cores: P single-threaded worker processes, contiguous frequency chunks
    ('tuned' variant - zgeev does not scale with BLAS threads at 2N=126)."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    from axisafe_b200 import cases
    f = cases.pipe_soil_pml(nf=POINTS_PER_GPU)['f']
    chunks = [(f[i * points_per_core:(i + 1) * points_per_core],) for i in range(cores)]
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(c[0][:2],) for c in chunks])          # warm-up: imports, mesh
        t0 = time.perf_counter()
        pool.map(_cpu_worker, chunks)
        wall = time.perf_counter() - t0
    npts = cores * points_per_core
    shipped = as_shipped_baseline() if as_shipped else None
    return {'as_shipped': shipped, 'value': npts / wall, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d of the 2000 sweep frequencies (%d per core, contiguous chunks, tracked within '
                      'a chunk), %d single-threaded processes, %.1f s wall' % (npts, points_per_core, cores, wall)}


_AS_SHIPPED = """
import sys, time
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oracle import safe_oracle as so
from axisafe_b200 import cases
c = cases.pipe_soil_pml(nf=%d)
G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
so.solve(G, c['n'], c['f'][:2])
t0 = time.perf_counter(); so.solve(G, c['n'], c['f'][:%d]); print(time.perf_counter() - t0)
"""


def as_shipped_baseline(points=40):
    """the reference loop the way its examples run it - ONE process, the BLAS
    library's default thread count (which at 2N = 126 is slower than one thread: oversubscription)."""
    env = {k: v for k, v in os.environ.items() if k not in ('OPENBLAS_NUM_THREADS', 'OMP_NUM_THREADS', 'MKL_NUM_THREADS')}
    try:
        out = subprocess.run([sys.executable, '-c', _AS_SHIPPED % (ROOT, os.path.join(ROOT, 'axisym-safe-python_b200'),
                                                                   POINTS_PER_GPU, points)],
                             env=env, capture_output=True, text=True, timeout=300)
        wall = float(out.stdout.strip().splitlines()[-1])
    except Exception as exc:                                            # pragma: no cover
        return {'value': None, 'note': 'as-shipped run failed: %s' % exc}
    return {'value': points / wall, 'unit': UNIT, 'processes': 1, 'blas_threads': 'library default',
            'sample': 'first %d sweep frequencies, %.1f s wall' % (points, wall)}


# ------------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its version banner on stdout: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import axisafe_b200 as ax
    from axisafe_b200 import cases, _native as nat

    nf_total = POINTS_PER_GPU * world
    c = cases.pipe_soil_pml(nf=nf_total)
    mesh = ax.mesh.Mesh()
    mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    with contextlib.redirect_stdout(io.StringIO()):
        mesh.assemble_matrices(c['n'])
    ops = mesh._device_ops
    N, n2 = ops.N, 2 * ops.N
    lo = rank * POINTS_PER_GPU
    halo = 1 if lo > 0 else 0
    w_local = 2 * np.pi * c['f'][lo - halo:lo + POINTS_PER_GPU]
    nloc = len(w_local)
    omega = torch.from_numpy(w_local).cuda()
    work = nat.EigWorkspace(N, nloc)
    order_dev = torch.arange(n2, dtype=torch.int32, device='cuda').repeat(nloc)
    L, st = nat.lib(), nat._stream(torch)
    k = torch.empty(nloc * n2, dtype=torch.complex128, device='cuda')
    psi = torch.empty(nloc * n2 * n2, dtype=torch.complex128, device='cuda')
    k_out, psi_out = torch.empty_like(k), torch.empty_like(psi)
    info = torch.zeros(nloc, dtype=torch.int32, device='cuda')
    arg = torch.empty((nloc - 1) * n2, dtype=torch.int32, device='cuda')
    amax = torch.empty((nloc - 1) * n2, dtype=torch.float64, device='cuda')
    need = int(L.axs_track_worksize(N, nloc - 1))
    ev = {name: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for name in ('reduce', 'hqr', 'modes', 'track', 'order')}
    per_kernel = {name: 0.0 for name in ev}

    def device_step(timed):
        """Whole hot path with inputs resident in HBM on one stream: 12 kernel launches."""
        P = nat._ptr
        ev['reduce'][0].record()
        nat.check(L.axs_reduce(N, ops.hb, P(ops.K0s), P(ops.Cb), P(ops.Mb), P(ops.K1c), P(ops.K6d), P(omega),
                               nloc, None, P(work.buf), work.bytes, st), 'axs_reduce')
        ev['reduce'][1].record(); ev['hqr'][0].record()
        nat.check(L.axs_hqr(N, nloc, P(k), P(info), P(work.buf), work.bytes, st), 'axs_hqr')
        ev['hqr'][1].record(); ev['modes'][0].record()
        nat.check(L.axs_modes(N, ops.hb, P(ops.Cb), P(ops.K6d), nloc, P(k), P(psi), P(work.buf), work.bytes, st),
                  'axs_modes')
        ev['modes'][1].record(); ev['track'][0].record()
        nat.check(L.axs_track_pairs(N, ops.hb, P(ops.Cb), P(ops.K6d), P(psi), nloc - 1, P(arg), P(amax),
                                    P(work.buf), need, st), 'axs_track_pairs')
        ev['track'][1].record(); ev['order'][0].record()
        nat.check(L.axs_apply_order(n2, nloc, P(order_dev), P(k), P(psi), P(k_out), P(psi_out), st),
                  'axs_apply_order')
        ev['order'][1].record()
        if timed:
            torch.cuda.synchronize()
            for name in ev:
                per_kernel[name] += ev[name][0].elapsed_time(ev[name][1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def product_step():
        """The product path: two frequency chunks on two streams (what solve() runs), 9 launches per chunk + 3."""
        P = nat._ptr
        kk, pp, ii, _ = nat.eig_sweep(ops, omega, want_modes=True, work=work)
        nat.check(L.axs_track_pairs(N, ops.hb, P(ops.Cb), P(ops.K6d), P(pp), nloc - 1, P(arg), P(amax),
                                    P(work.buf), need, nat._stream(torch)), 'axs_track_pairs')
        nat.check(L.axs_apply_order(n2, nloc, P(order_dev), P(kk), P(pp), P(k_out), P(psi_out),
                                    nat._stream(torch)), 'axs_apply_order')
        return ii

    for _ in range(args.warmup):
        product_step()
        device_step(False)
    barrier()
    with ClockSampler(local, enabled=(rank == 0)) as clocks:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            info = product_step()
        t1.record()
        barrier()
        # second timed pass on ONE stream with CUDA events around every kernel (roofline / shares)
        for _ in range(args.steps):
            device_step(True)
        barrier()
        dev_ms = t0.elapsed_time(t1)
        # ---- end to end through the public API with host buffers ------------------------------
        sol = ax.solver.WaveElementAxisym(mesh)
        f_host = c['f']
        e2e_times, e2e_k_times = [], []
        with contextlib.redirect_stdout(io.StringIO()):
            sol.solve(f_host); _ = sol.psi_hat                        # warm-up (pinned allocation)
            for _ in range(max(1, min(args.steps, 5))):
                barrier()
                ta = time.perf_counter()
                sol.solve(f_host)            # H2D omega, kernels, D2H k + tracking tables, scan, gather
                tb = time.perf_counter()
                psi_host = sol.psi_hat       # D2H of this rank's tracked mode shapes (pinned)
                torch.cuda.synchronize()
                tc = time.perf_counter()
                e2e_k_times.append(tb - ta); e2e_times.append(tc - ta)
                del psi_host
                sol.psi_hat = None
    if int(info.max().item()) != 0:
        raise RuntimeError('QR did not converge for some frequency point')

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev_ms = reduce_max(dev_ms)
    e2e_s = reduce_max(float(np.median(e2e_times)))
    e2e_k_s = reduce_max(float(np.median(e2e_k_times)))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = dev_ms / args.steps
    value = nf_total / (ms_per_step * 1e-3)
    hqr_ms = per_kernel['hqr'] / args.steps
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    fp64_peak, how = measure_fp64_peak(torch)
    achieved = 80.0 * n2 ** 3 * nloc / (hqr_ms * 1e-3) / 1e12
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'points_per_gpu': POINTS_PER_GPU, 'n2': n2,
                   'l2': 'working set 1.5 GB per sweep >> 126 MB L2 (no flush needed)',
                   'parallelism': 'frequency axis sharded, %d rank(s), 1-point halo; 2 frequency chunks on 2 CUDA streams per rank (value); e2e = solve(): 4-chunk pipeline with psi_hat download overlapped at N=1; at N>1 the same pipeline over interleaved pieces of the frequency axis (cross-rank tracking chain per chunk over gloo, psi_hat pieces downloaded behind the next chunk, NCCL all-gather of k / order at the end)' % world},
        'kernel_ms': {k_: v / args.steps for k_, v in per_kernel.items()},
        'roofline': {'bound': 'tensor', 'kernel': 'k_hqr_mb (4 stage launches)', 'achieved': achieved, 'peak': fp64_peak,
                     'unit': 'TFLOP/s', 'frac': achieved / fp64_peak,
                     'traffic': int(nloc * 472.4e3),
                     'traffic_note': 'DRAM bytes (read + write) of the QR cascade per sweep, from ncu --set full '
                                     '(profiles/r1c_ncu_summary.csv): 130.7 + 112.5 + 74.6 + 55.9 MB read and 98.7 MB '
                                     'written per 1000 matrices over its four stage launches; the algorithmic bytes '
                                     'are the packed Hessenberg matrix once (130 KB per matrix) - the rest is the '
                                     'leading block parked and re-read between stages',
                     'note': 'FP64 pipe (DFMA/DMMA share one roof on B200); algorithmic flops = 80*n2^3 per '
                             'point for the QR kernel (zhseqr share of F_alg = 100*n2^3); peak = ' + how,
                     'pipeline_frac': value / world * f_alg(n2) / 1e12 / fp64_peak},
        'e2e': {'value': nf_total / e2e_s, 'unit': UNIT,
                # per rank.  up: omega + the tracked column order of the rank's block; down: tracking tables
                # (arg int32 + amax f64) + info + psi_hat block, and k: native + tracked table at N=1,
                # the tracked [nf_total, 2N] table + the order table at N>1 (k_native stays on the device)
                'h2d_bytes_per_step': int(8 * nloc + POINTS_PER_GPU * n2 * 4),
                'd2h_bytes_per_step': int((nloc - 1) * n2 * 12 + POINTS_PER_GPU * 4 + POINTS_PER_GPU * n2 * n2 * 16
                                          + (2 * nloc * n2 * 16 if world == 1 else nf_total * n2 * 20)),
                'k_only': {'value': nf_total / e2e_k_s, 'note': 'solve() without reading psi_hat back'}},
        'gpu_launches': (2 * 9 + 3) * args.steps,   # per chunk: balance, reduce_oc, hqr x4, modes, wy_tfactor, backnorm_wy; then bpsi, track_mma, apply_order
        'clocks': clocks.summary(),
        'hbm_peak_gbs': peaks.get('hbm_gbs'),
    }
    out['cpu_baseline'] = cpu_baseline(as_shipped=True) if world == 1 and not args.no_cpu else None
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_fp64_peak(torch):
    """FP64 peak is not in MEASURED_PEAKS.json: cuBLAS DGEMM 8192^3, best of 5."""
    try:
        a = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
        b = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12, 'measured live: torch.matmul fp64 8192^3 best of 5'
    except Exception as exc:                                            # pragma: no cover
        return 37.0, 'nominal 37 TFLOP/s (live DGEMM failed: %s)' % exc


# ------------------------------------------------------------------------ reference arm
def reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if rank != 0:
        return
    rates = []
    for _ in range(args.warmup + args.steps):
        rates.append(cpu_baseline(points_per_core=60))
    timed = rates[args.warmup:]
    value = float(np.mean([r['value'] for r in timed]))
    base = dict(timed[-1], value=value)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': POINTS_PER_GPU * world / value * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'points_per_gpu': POINTS_PER_GPU, 'n2': 126,
                   'note': 'reference algorithm (scipy.linalg.eig loop) on the host cores; each step is a bounded '
                           'sample, throughput extrapolates linearly (frequencies are independent)'},
        'cpu_baseline': base,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
