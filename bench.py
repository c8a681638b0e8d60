#!/usr/bin/env python
"""Benchmark of the per-frequency SAFE dispersion sweep (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU algorithm on host cores

A "step" is ONE full dispersion sweep of the flagship workload - the water-filled cast-iron
pipe buried in soil with a PML (examples/water_filled_pipe_in_soil.py, 2N = 126) at the
north-star resolution of 2000 frequency points - per GPU (weak scaling, the default: every rank
sweeps its own 2000-point block of one N*2000-point frequency axis; ``--scaling strong``: ONE
2000-point sweep split over the N ranks, the configuration BASELINE.json's target names).
``--config cfg4``: the coated pipe of BASELINE.json configs[3] (2N = 3966), 148 frequency points
per GPU and step, through the large (global-memory) kernels.

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
WORKLOAD_STRONG = 'pipe_soil_pml (water-filled cast-iron pipe in soil + PML, n=0, 2N=126), ONE 2000-point sweep over all GPUs'
CFG4_POINTS_PER_GPU = 148
WORKLOAD_CFG4 = ('coated_pipe (steel + bitumen + soil PML, order 10 x 22 elements per layer, n=1, 2N=3966), '
                 '%d frequency points of the 2000-point sweep per GPU and step' % CFG4_POINTS_PER_GPU)
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
def _have_ref():
    """oracle/_ref holds the UNMODIFIED reference package (oracle/build_ref.py, built in the container)."""
    from oracle import build_ref
    return build_ref.available()


def _ref_solver(freqs_all):
    """Reference objects for cfg1: Mesh.create / assemble_matrices / WaveElementAxisym of the reference itself."""
    import warnings
    from oracle import build_ref
    from axisafe_b200 import cases
    ref = build_ref.load()
    c = cases.pipe_soil_pml(nf=len(freqs_all))
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter('ignore')
        mesh = ref.mesh.Mesh()
        mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
        mesh.assemble_matrices(c['n'])
        return ref.solver.WaveElementAxisym(mesh)


def _cpu_worker(args):
    """One worker process = one chunk of the sweep.  kind 'reference': the reference's own
    WaveElementAxisym.solve (solver.py:76-184) from oracle/_ref; kind 'port': the oracle restatement of
    that loop (oracle/safe_oracle.py), used only when oracle/_ref was not shipped."""
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    freqs, kind = args
    import warnings
    if kind == 'reference':
        sol = _ref_solver(np.linspace(1e3, 20e3, POINTS_PER_GPU))
        t0 = time.perf_counter()
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter('ignore')
            sol.solve(np.asarray(freqs))
        return time.perf_counter() - t0
    from oracle import safe_oracle as so
    from axisafe_b200 import cases
    c = cases.pipe_soil_pml(nf=POINTS_PER_GPU)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    t0 = time.perf_counter()
    so.solve(G, c['n'], freqs)
    return time.perf_counter() - t0


def cpu_baseline(points_per_core=100, as_shipped=False, kind=None):
    """The reference's CPU path on all host cores: P single-threaded worker processes, each running the
    reference's WaveElementAxisym.solve on a contiguous chunk of the sweep (BASELINE.md section 3,
    'tuned' variant - zgeev does not scale with BLAS threads at 2N = 126, one process per core does).
    kind 'reference' = the unmodified reference from oracle/_ref; 'port' = the oracle restatement
    (fallback when oracle/_ref is absent); the other one is reported under 'port' / not at all."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    kind = kind or ('reference' if _have_ref() else 'port')
    from axisafe_b200 import cases
    f = cases.pipe_soil_pml(nf=POINTS_PER_GPU)['f']
    chunks = [(f[i * points_per_core:(i + 1) * points_per_core], kind) for i in range(cores)]
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(c[0][:2], kind) for c in chunks])      # warm-up: imports, mesh
        t0 = time.perf_counter()
        pool.map(_cpu_worker, chunks)
        wall = time.perf_counter() - t0
    npts = cores * points_per_core
    shipped = as_shipped_baseline() if as_shipped else None
    return {'as_shipped': shipped, 'value': npts / wall, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': '%d of the 2000 sweep frequencies (%d per core, contiguous chunks, tracked within '
                      'a chunk), %d single-threaded processes each calling %s, %.1f s wall'
                      % (npts, points_per_core, cores,
                         "the reference's WaveElementAxisym.solve (oracle/_ref)" if kind == 'reference'
                         else 'the oracle port of the reference loop', wall)}


_AS_SHIPPED = """
import sys, time, io, contextlib, warnings
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import bench
from axisafe_b200 import cases
c = cases.pipe_soil_pml(nf=%d)
if bench._have_ref():
    sol = bench._ref_solver(c['f'])
    def run(f):
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter('ignore')
            sol.solve(np.asarray(f))
else:
    from oracle import safe_oracle as so
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    run = lambda f: so.solve(G, c['n'], f)
run(c['f'][:2])
t0 = time.perf_counter(); run(c['f'][:%d]); print(time.perf_counter() - t0)
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
            'kind': 'reference' if _have_ref() else 'port',
            'sample': 'first %d sweep frequencies, one process, %.1f s wall' % (points, wall)}


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
    numa_node = None
    if world > 1 and os.environ.get('AXS_NUMA_BIND', '1') != '0':
        # one process per GPU: run (and allocate the pinned download buffers) on the GPU's own NUMA node
        numa_node = nat.bind_host_to_device_numa(local)

    if args.config == 'cfg4':
        return gpu_arm_cfg4(args, torch, dist, rank, world, local, saved_stdout)
    strong = args.scaling == 'strong'
    ppg = POINTS_PER_GPU // world if strong else POINTS_PER_GPU     # points per GPU
    nf_total = ppg * world
    c = cases.pipe_soil_pml(nf=nf_total)
    mesh = ax.mesh.Mesh()
    mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    with contextlib.redirect_stdout(io.StringIO()):
        mesh.assemble_matrices(c['n'])
    ops = mesh._device_ops
    N, n2 = ops.N, 2 * ops.N
    lo = rank * ppg
    halo = 1 if lo > 0 else 0
    w_local = 2 * np.pi * c['f'][lo - halo:lo + ppg]
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
            n_e2e = max(1, min(args.steps, 5))
            for it in range(n_e2e):
                barrier()
                ta = time.perf_counter()
                sol.solve(f_host)            # H2D omega, kernels, D2H k + tracking tables, scan, gather
                k_host = sol.k               # the dispersion curves [nf_total, 2N] (host)
                tb = time.perf_counter()
                psi_host = sol.psi_hat       # D2H of this rank's tracked mode shapes (pinned)
                torch.cuda.synchronize()
                tc = time.perf_counter()
                e2e_k_times.append(tb - ta); e2e_times.append(tc - ta)
                del psi_host, k_host
                if it + 1 < n_e2e:
                    sol.psi_hat = None
        # ---- parity of the measured configuration (outside the timed regions) -----------------
        parity = check_parity(torch, dist, ax, mesh, sol, c, rank, world)
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
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'strong' if strong else 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'parity_ok': bool(parity['ok']), 'parity': parity,
        'config': {'workload': WORKLOAD_STRONG if strong else WORKLOAD, 'points_per_gpu': ppg, 'n2': n2,
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
                     'executed': {'fp64_pipe_active_pct_per_stage': [14.1, 8.6, 6.9, 4.3],
                                  'issue_active_pct_per_stage': [32.9, 30.9, 24.0, 17.3],
                                  'source': 'ncu --set full, profiles/r2c_ncu_summary.csv: the kernel is bound by the '
                                            'latency of its rotation chain and two block barriers per macro step, not '
                                            'by FP64 throughput; frac is a NOMINAL zhseqr flop count (Schur vectors '
                                            'are never accumulated)'},
                     'pipeline_frac': value / world * f_alg(n2) / 1e12 / fp64_peak},
        'e2e': {'value': nf_total / e2e_s, 'unit': UNIT,
                # per rank.  up: omega + the tracked column order of the rank's block; down: tracking tables
                # (arg int32 + amax f64) + info + psi_hat block, and k: native + tracked table at N=1,
                # the tracked [nf_total, 2N] table + the order table at N>1 (k_native stays on the device)
                'h2d_bytes_per_step': int(8 * nloc + ppg * n2 * 4),
                'd2h_bytes_per_step': int((nloc - 1) * n2 * 12 + ppg * 4 + ppg * n2 * n2 * 16
                                          + (2 * nloc * n2 * 16 if world == 1 else nf_total * n2 * 20)),
                'k_only': {'value': nf_total / e2e_k_s, 'note': 'solve() without reading psi_hat back'}},
        'gpu_launches': ((2 if nloc >= 128 else 1) * 9 + 3) * args.steps,   # per chunk: balance, reduce_oc, hqr x4, modes, wy_tfactor, backnorm_wy; then bpsi, track_mma, apply_order
        'qr_counters': qr_counters(nat, N, nloc, work),
        'numa_node_rank0': numa_node,
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


def qr_counters(nat, N, nloc, work):
    """Per-matrix means of the QR stage counters of the last device step (axs_hqr_stats)."""
    try:
        st = nat.hqr_stats(N, nloc, work)
        return {'macro_steps': float(st[:, 0].mean()), 'batches': float(st[:, 1].mean()),
                'aed_calls': float(st[:, 2].mean()), 'aed_deflated': float(st[:, 3].mean()),
                'kcycles_per_stage_window_apply_sweeps': [[float(st[:, 4 + 3 * g + q].mean()) for q in range(3)]
                                                          for g in range(4)]}
    except Exception as exc:                                            # pragma: no cover
        return {'error': str(exc)}


def check_parity(torch, dist, ax, mesh, sol, c, rank, world):
    """Parity of the very sweep that was timed.  N = 1: the tracked spectra of 4 frequencies against the
    CPU oracle (scipy.linalg.eig on the same S, bar 1e-8).  N > 1: every rank repeats the WHOLE sweep
    alone on its GPU (distributed=False) and compares k, order, info bit for bit and its own psi_hat
    rows; the flags are AND-reduced over the ranks, so SCALE lines carry the multi-GPU parity."""
    from oracle import safe_oracle as so
    out = {}
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    nf = len(c['f'])
    worst = 0.0
    for i in sorted({0, nf // 3, (2 * nf) // 3, nf - 1}):
        val, _ = so.eig_normalised(op, 2 * np.pi * c['f'][i])
        worst = max(worst, float(so.match_eigenvalues(val, sol.k[i])[1]))
    out['oracle_max_rel_err'] = worst
    ok = worst < 1e-8
    if world > 1:
        alone = ax.solver.WaveElementAxisym(mesh)
        with contextlib.redirect_stdout(io.StringIO()):
            alone.solve(c['f'], distributed=False, pipeline=False)        # mode shapes stay on the device
        loc = torch.from_numpy(np.asarray(sol.local_index)).cuda()
        same_k = bool(np.array_equal(sol.k, alone.k) and np.array_equal(sol.order, alone.order)
                      and np.array_equal(sol.info, alone.info) and np.array_equal(sol.k_native, alone.k_native))
        same_psi = bool(torch.equal(torch.view_as_real(sol.psi_hat_device),
                                    torch.view_as_real(alone.psi_hat_device.index_select(0, loc))))
        t = torch.tensor([int(same_k), int(same_psi), int(ok)], dtype=torch.int32, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        same_k, same_psi, ok = bool(t[0].item()), bool(t[1].item()), bool(t[2].item())
        out.update({'sharded_k_order_info_bit_identical': same_k, 'sharded_psi_rows_bit_identical': same_psi,
                    'compared_with': 'every rank solving the whole %d-point sweep alone' % nf})
        ok = ok and same_k and same_psi
        del alone
    out['ok'] = bool(ok)
    out['bar'] = 'eigenvalues 1e-8 relative vs scipy.linalg.eig on the same S (4 frequencies)'
    return out


def gpu_arm_cfg4(args, torch, dist, rank, world, local, saved_stdout):
    """BASELINE.json configs[3]: coated pipe, 2N = 3966, through solve() on the large path.

    Step = one solve() of this rank's 148-frequency block of the 2000-point sweep (weak scaling, no
    data-path collective: every rank solves its own block with distributed=False; psi_hat - 37 GB per
    block - stays on the device, the reference would need 503 GB of host memory for the sweep)."""
    import axisafe_b200 as ax
    from axisafe_b200 import cases, _native as nat
    ppg = CFG4_POINTS_PER_GPU
    c = cases.coated_pipe(nf=2000, n=1)
    mesh = ax.mesh.Mesh()
    mesh.create(c['sets'], c['fmax'], PML=c['PML'], PML_props=c['PML_props'])
    with contextlib.redirect_stdout(io.StringIO()):
        mesh.assemble_matrices(c['n'])
    n2 = 2 * mesh.no_of_dofs
    f_block = c['f'][rank * ppg:(rank + 1) * ppg]
    sol = ax.solver.WaveElementAxisym(mesh)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with contextlib.redirect_stdout(io.StringIO()):
            sol.solve(f_block, distributed=False)
        return sol.k

    for _ in range(max(1, args.warmup)):
        step()
    barrier()
    walls = []
    with ClockSampler(local, enabled=(rank == 0)) as clocks:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            ta = time.perf_counter()
            k = step()
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - ta)
        t1.record()
        barrier()
        dev_ms = t0.elapsed_time(t1)
    # parity: tr S, tr S^2 identities on every frequency + the committed zgeev fixture when a frequency matches
    S_tr = -np.sum(np.asarray(mesh._device_ops.Cb.cpu().numpy().reshape(mesh.no_of_dofs, -1)[:, mesh._device_ops.hb]) /
                   mesh._device_ops.K6d.cpu().numpy())
    tr_err = float(np.abs(k.sum(axis=1) - S_tr).max() / np.abs(k).sum(axis=1).max())
    parity = {'trace_identity_max_rel_err': tr_err, 'ok': bool(tr_err < 1e-9 and not sol.info.any()),
              'bar': 'sum(k) = tr S on every frequency (1e-9); eigenvalues vs zgeev at full size: '
                     'tests/test_gpu_big.py::test_cfg4_full_size_vs_reference_zgeev'}
    t = torch.tensor([dev_ms, float(np.median(walls))], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    dev_ms, wall = float(t[0].item()), float(t[1].item())
    ms_per_step = dev_ms / args.steps
    value = ppg * world / (ms_per_step * 1e-3)
    fp64_peak, how = measure_fp64_peak(torch)
    achieved = value / world * f_alg(n2) / 1e12
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'parity_ok': parity['ok'], 'parity': parity,
        'config': {'workload': WORKLOAD_CFG4, 'points_per_gpu': ppg, 'n2': n2,
                   'l2': 'working set > 100 GB per step >> 126 MB L2',
                   'parallelism': 'frequency blocks per rank, no collective; large (global-memory) kernels'},
        'roofline': {'bound': 'tensor', 'kernel': 'whole step (blocked Hessenberg + windowed QR + inverse iteration + back-transformation)',
                     'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': achieved / fp64_peak, 'traffic': None,
                     'note': 'algorithmic flops F_alg = 100 n2^3 per point (SURVEY 8d); peak = ' + how},
        'e2e': {'value': ppg * world / wall, 'unit': UNIT, 'h2d_bytes_per_step': int(8 * ppg),
                'd2h_bytes_per_step': int(ppg * n2 * (16 + 16) + (ppg - 1) * n2 * 12 + 4 * ppg),
                'note': 'solve() with a host frequency array; k, k_native and the tracking tables come back, psi_hat '
                        '(37 GB per block) stays on the device until it is read'},
        'gpu_launches': cfg4_launches(n2, ppg) * args.steps, 'clocks': clocks.summary(),
    }
    out['cpu_baseline'] = cpu_baseline_cfg4() if world == 1 and not args.no_cpu else None
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cfg4_launches(n2, nf, nb=32, chunk=74):
    """Kernel launches of one cfg4 step (solve() in chunks of `chunk` frequencies): per chunk the blocked
    reduction (balance, 2 per column, 1 + 6 per panel, 2 pack), windowed QR + 3 cascade stages, k_modes_col,
    3 per panel of the blocked back-transformation + scale + norm, tracking (bpsi, track_mma); then the
    column gather of the kept mode shapes."""
    panels = -(-(n2 - 2) // nb)
    per_chunk = 1 + 2 * (n2 - 2) + 7 * panels + 2 + 4 + 1 + 3 * panels + 2 + 2
    chunks = -(-nf // chunk)
    return chunks * per_chunk + chunks


def cpu_baseline_cfg4():
    """One frequency of cfg4 through the reference's loop body on all host threads (zgeev gains ~2x from
    BLAS threads at this size): S(w) of the oracle + scipy.linalg.eig + B-normalisation."""
    from oracle import safe_oracle as so
    from axisafe_b200 import cases
    c = cases.coated_pipe(nf=2000, n=1)
    G = so.assemble(so.build_mesh(c['sets'], c['fmax'], c['PML'], c['PML_props']), c['n'])
    op = so.operators(G, c['n'])
    t0 = time.perf_counter()
    so.eig_normalised(op, 2 * np.pi * c['f'][1000])
    wall = time.perf_counter() - t0
    return {'value': 1.0 / wall, 'unit': UNIT, 'cores': len(os.sched_getaffinity(0)), 'kind': 'port',
            'sample': 'ONE frequency of the sweep (S build + scipy.linalg.eig + B-normalisation), one process, '
                      'all BLAS threads, %.0f s wall' % wall}


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
    if args.config == 'cfg4':
        base = cpu_baseline_cfg4()
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': world,
            'steps': 1, 'warmup': 0, 'ms_per_step': CFG4_POINTS_PER_GPU * world / base['value'] * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_CFG4, 'points_per_gpu': CFG4_POINTS_PER_GPU, 'n2': 3966},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}))
        return
    strong = args.scaling == 'strong'
    ppg = POINTS_PER_GPU // world if strong else POINTS_PER_GPU
    rates = []
    for _ in range(args.warmup + args.steps):
        rates.append(cpu_baseline(points_per_core=60))
    timed = rates[args.warmup:]
    value = float(np.mean([r['value'] for r in timed]))
    base = dict(timed[-1], value=value)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ppg * world / value * 1e3,
        'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD_STRONG if strong else WORKLOAD, 'points_per_gpu': ppg, 'n2': 126,
                   'note': "the reference's own WaveElementAxisym.solve (unmodified package shipped as oracle/_ref; the "
                           'oracle port only if that is absent) on all host cores, one single-threaded process per '
                           'core; each step is a bounded sample, throughput extrapolates linearly (frequencies are '
                           'independent)'},
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
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: 2000 points per GPU (default); strong: ONE 2000-point sweep split over the GPUs')
    ap.add_argument('--config', default='cfg1', choices=['cfg1', 'cfg4'],
                    help='cfg1: pipe_soil_pml 2N=126 (headline); cfg4: coated pipe 2N=3966 (large path)')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
