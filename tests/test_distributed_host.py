"""world_size=2 gloo test (CPU) of the host-side sharding logic of WaveElementAxisym.solve:
block bounds + halo convention, all-gather of the per-rank tables, permutation scan."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _global_tables(nf, n2, seed=3):
    rng = np.random.default_rng(seed)
    k = rng.standard_normal([nf, n2]) + 1j * rng.standard_normal([nf, n2])
    arg = np.array([rng.permutation(n2) for _ in range(nf - 1)], dtype=np.int32)
    amax = 0.93 + 0.07 * rng.random([nf - 1, n2])
    amax[5, 3] = 0.4                                   # one weak mode
    amax[17, 1] = amax[17, 6] = 0.2                    # two weak modes: re-seated by position
    amax[18, 2] = amax[18, 7] = amax[18, 9] = 0.1      # consecutive fix-up rows
    amax[nf // 2 - 1, 4] = amax[nf // 2 - 1, 5] = 0.3  # fix-up on the pair at a block seam (world 2)
    amax[nf - 2, 0] = amax[nf - 2, 8] = 0.5            # fix-up on the very last frequency
    arg[11] = arg[11][0]                               # a non-injective arg-max row (duplicates)
    arg[11, 3] = (arg[11, 0] + 1) % n2
    amax[12, :] = 0.95
    amax[12, (arg[11, 0] + 2) % n2] = 0.1              # weak, but not visited after the collapse
    info = np.zeros(nf, np.int32)
    return k, arg, amax, info


def _worker(rank, world, port, nf, n2, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), 'axisym-safe-python_b200'))
    from axisafe_b200 import solver, _native
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    k, arg, amax, info = _global_tables(nf, n2)
    bounds = [(nf * r) // world for r in range(world + 1)]
    lo, hi = bounds[rank], bounds[rank + 1]
    halo = 1 if lo > 0 else 0
    # what the rank computes locally: its block of k, and the tracking pairs (i-1, i) for
    # i = lo-halo+1 .. hi-1, i.e. global pair indices lo-halo .. hi-2
    k_loc = torch.from_numpy(k[lo:hi].copy())
    arg_loc = torch.from_numpy(arg[lo - halo:hi - 1].copy())
    amax_loc = torch.from_numpy(amax[lo - halo:hi - 1].copy())
    info_loc = torch.from_numpy(info[lo:hi].copy())
    sol = solver.WaveElementAxisym(None)
    order_loc, _ = solver._track_sharded(sol, torch, _native, bounds, rank, n2, k_loc, arg_loc, amax_loc, info_loc)
    assert np.array_equal(order_loc, sol.order[lo:hi])
    assert np.array_equal(sol.k_native, k)
    np.save(os.path.join(out_dir, 'order_%d.npy' % rank), sol.order)
    np.save(os.path.join(out_dir, 'k_%d.npy' % rank), sol.k)
    dist.destroy_process_group()


@pytest.mark.parametrize('nf,world', [(41, 2), (64, 2), (50, 3), (37, 4)])
def test_sharded_tracking_matches_sequential_scan(tmp_path, built_lib, nf, world):
    """Segmented relative scan + chain + NCCL-style gathers == the sequential reference scan."""
    n2 = 12
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nf, n2, str(tmp_path)), nprocs=world, join=True)
    from axisafe_b200 import _native
    k, arg, amax, info = _global_tables(nf, n2)
    ref = _native.track_scan(arg, amax, nf, n2)
    ref_k = np.take_along_axis(k, ref.astype(np.int64), axis=1)
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), 'order_%d.npy' % r)), ref)
        assert np.array_equal(np.load(os.path.join(str(tmp_path), 'k_%d.npy' % r)), ref_k)


def test_segmented_scan_random(built_lib):
    """Host-only: any cut of the frequency axis into blocks reproduces the sequential scan."""
    from axisafe_b200 import _native
    rng = np.random.default_rng(11)
    n2 = 9
    for trial in range(40):
        nf = int(rng.integers(2, 60))
        arg = np.array([rng.permutation(n2) for _ in range(nf - 1)], dtype=np.int32)
        for i in rng.integers(0, nf - 1, size=3):
            arg[i, rng.integers(0, n2)] = rng.integers(0, n2)          # duplicates
        amax = np.where(rng.random([nf - 1, n2]) < 0.04, 0.5, 0.97)
        ref = _native.track_scan(arg, amax, nf, n2)
        world = int(rng.integers(1, min(nf, 6) + 1))
        bounds = [(nf * r) // world for r in range(world + 1)]
        parts, payloads = [], []
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            halo = 1 if lo > 0 else 0
            a, m = arg[lo - halo:hi - 1], amax[lo - halo:hi - 1]
            rel, starts = _native.track_scan_segments(a, m, n2)
            ends = starts[1:] + [a.shape[0] + 1]
            fix = np.asarray(starts[1:], dtype=np.int64) - 1
            payloads.append((np.stack([rel[e - 1] for e in ends]), a[fix], m[fix]))
            parts.append((rel, starts, ends, halo))
        sig = _native.chain_segments(payloads, n2)
        got = np.concatenate([np.concatenate([rel[a:e][:, sig[r][s]] for s, (a, e) in enumerate(zip(starts, ends))])[halo:]
                              for r, (rel, starts, ends, halo) in enumerate(parts)])
        assert np.array_equal(got, ref), trial


@pytest.mark.parametrize('nf,world,unit', [(301, 2, 148), (1200, 2, 148), (4000, 2, 148), (8000, 4, 148), (16000, 8, 148), (1111, 3, 16), (2051, 4, 132)])
def test_interleaved_plan_and_chunk_chain(built_lib, nf, world, unit):
    """Host logic of the sharded pipeline (solver._solve_sharded_pipelined): the interleaved plan
    covers every frequency exactly once, and the per-chunk chain with the carried order reproduces
    the sequential scan on every piece."""
    from axisafe_b200 import solver, _native
    pieces = solver._shard_plan(nf, world, unit)
    flat = [p for chunk in pieces for p in chunk]
    assert flat[0][0] == 0 and flat[-1][1] == nf
    assert all(a[1] == b[0] for a, b in zip(flat, flat[1:])) and all(hi > lo for lo, hi in flat)
    n2 = 10
    rng = np.random.default_rng(nf)
    arg = np.array([rng.permutation(n2) for _ in range(nf - 1)], dtype=np.int32)
    amax = np.where(rng.random([nf - 1, n2]) < 0.004, 0.5, 0.97)
    for lo, hi in flat[1:]:
        amax[lo - 1, :2] = 0.3 if rng.random() < 0.3 else 0.97       # fix-ups on the piece seams
    ref = _native.track_scan(arg, amax, nf, n2)
    got = np.empty_like(ref)
    carry = None
    for chunk in pieces:
        payloads, keep = [], []
        for lo, hi in chunk:
            halo = 1 if lo > 0 else 0
            a, m = arg[lo - halo:hi - 1], amax[lo - halo:hi - 1]
            rel, starts = _native.track_scan_segments(a, m, n2)
            ends = starts[1:] + [a.shape[0] + 1]
            fix = np.asarray(starts[1:], dtype=np.int64) - 1
            payloads.append((np.stack([rel[e - 1] for e in ends]), a[fix], m[fix]))
            keep.append((rel, starts, ends, halo))
        sigma, carry = _native.chain_segments(payloads, n2, sigma0=carry, return_final=True)
        for r, ((lo, hi), (rel, starts, ends, halo)) in enumerate(zip(chunk, keep)):
            rows = np.concatenate([rel[p:e][:, sigma[r][s]] for s, (p, e) in enumerate(zip(starts, ends))])
            got[lo:hi] = rows[halo:]
    assert np.array_equal(got, ref)


def test_shard_plan_properties():
    """Every plan solve() can ask for (>= 128 points per rank) tiles the axis with non-empty pieces."""
    from hypothesis import given, settings, strategies as st
    from axisafe_b200 import solver

    @settings(max_examples=300, deadline=None)
    @given(world=st.integers(2, 8), per=st.integers(128, 5000), extra=st.integers(0, 7),
           unit=st.sampled_from([108, 132, 148]), first=st.sampled_from([None, 0.5, 0.67, 0.77, 0.888, 0.95]))
    def check(world, per, extra, unit, first):
        nf = per * world + extra % world
        pieces = solver._shard_plan(nf, world, unit, first=first)
        flat = [p for chunk in pieces for p in chunk]
        assert 2 <= len(pieces) <= 8 and all(len(chunk) == world for chunk in pieces)
        assert flat[0][0] == 0 and flat[-1][1] == nf
        assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
        assert all(hi > lo for lo, hi in flat)
    check()


def _worker_exchange(rank, world, port, out_dir):
    """Fixed-size segment exchange + completion of per-rank rows (solver._exchange_segments / _complete_rows)."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), 'axisym-safe-python_b200'))
    from axisafe_b200 import solver
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n2, smax = 10, 3
    for round_, big in enumerate((False, True)):
        rng = np.random.default_rng(100 * round_ + rank)
        S = 1 + rank % smax if not (big and rank == 1) else smax + 2          # one rank overflows the fixed buffer
        comps = rng.integers(0, n2, size=[S, n2]).astype(np.int32)
        d_arg = rng.integers(0, n2, size=[S - 1, n2]).astype(np.int32)
        d_amax = rng.random([S - 1, n2])
        cap = 2 + n2 * (4 * smax - 3)
        send = torch.zeros(cap, dtype=torch.int32)
        recv = torch.zeros(world * cap, dtype=torch.int32)
        got = solver._exchange_segments(torch, dist, None, (comps, d_arg, d_amax, rank), n2, smax, send, recv, world)
        assert len(got) == world
        for r, g in enumerate(got):
            rr = np.random.default_rng(100 * round_ + r)
            Sr = 1 + r % smax if not (big and r == 1) else smax + 2
            assert np.array_equal(g[0], rr.integers(0, n2, size=[Sr, n2]).astype(np.int32))
            assert np.array_equal(g[1], rr.integers(0, n2, size=[Sr - 1, n2]).astype(np.int32))
            assert np.array_equal(g[2], rr.random([Sr - 1, n2]))
            assert g[3] == r
    # rows of an interleaved layout completed on every rank, real and complex, 2-D and 3-D
    nf = 23
    sol = solver.WaveElementAxisym(None)
    sol.w = np.zeros(nf)
    sol._sharded = True
    sol.local_index = np.arange(rank, nf, world)
    full_r = np.arange(nf * 4, dtype=float).reshape(nf, 4)
    full_c = (np.arange(nf * 6).reshape(nf, 2, 3) * (1 + 0.5j)).astype(complex)
    assert np.array_equal(solver._complete_rows(sol, full_r[sol.local_index]), full_r)
    out = solver._complete_rows(sol, full_c[sol.local_index])
    assert out.dtype == complex and np.array_equal(out, full_c)
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_segment_exchange_and_row_completion(tmp_path, world):
    mp.spawn(_worker_exchange, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
