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
    amax[5, 3] = 0.4
    amax[17, 1] = amax[17, 6] = 0.2
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
    k_all, arg_all, amax_all, info_all = solver._gather_blocks(torch, bounds, rank, n2, k_loc, arg_loc,
                                                               amax_loc, info_loc)
    assert np.array_equal(k_all.numpy(), k)
    assert np.array_equal(arg_all.numpy(), arg)
    assert np.array_equal(amax_all.numpy(), amax)
    order = _native.track_scan(arg_all.numpy(), amax_all.numpy(), nf, n2)
    np.save(os.path.join(out_dir, 'order_%d.npy' % rank), order)
    dist.destroy_process_group()


@pytest.mark.parametrize('nf', [41, 64])
def test_two_rank_gather_and_scan(tmp_path, built_lib, nf):
    n2, world = 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nf, n2, str(tmp_path)), nprocs=world, join=True)
    from axisafe_b200 import _native
    k, arg, amax, info = _global_tables(nf, n2)
    ref = _native.track_scan(arg, amax, nf, n2)
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), 'order_%d.npy' % r)), ref)
