"""ctypes binding of libaxisafe_b200.so (include/axisafe_b200.h) + torch device buffers.

PyTorch is used only for device memory (``torch.empty(..., device='cuda')``,
``.data_ptr()``), streams and ``torch.distributed``; every numerical kernel is in
the shared library.  There is deliberately no CPU fallback: a missing library or a
missing GPU raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AXS_LIBRARY: another build of the same library (A/B builds of tests/run_ab_aed*.sh); default = the in-tree product build
LIB_PATH = os.environ.get('AXS_LIBRARY') or os.path.join(_HERE, '_lib', 'libaxisafe_b200.so')
_lib = None

_c_int_p = ctypes.c_void_p
_SIGNATURES = {
    'axs_version': (ctypes.c_int, []),
    'axs_strerror': (ctypes.c_char_p, [ctypes.c_int]),
    'axs_max_small_n2': (ctypes.c_int, []),
    'axs_max_n2': (ctypes.c_int, []),
    'axs_element_matrices': (ctypes.c_int, [ctypes.c_int] + [ctypes.c_void_p] * 6),
    'axs_assemble_operators': (ctypes.c_int, [ctypes.c_int] + [ctypes.c_void_p] * 5 +
                               [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5),
    'axs_build_S': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 6 + [ctypes.c_int] +
                    [ctypes.c_void_p] * 2),
    'axs_eig_worksize': (ctypes.c_longlong, [ctypes.c_int, ctypes.c_int]),
    'axs_eig_sweep': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 6 + [ctypes.c_int] +
                      [ctypes.c_void_p] * 4 + [ctypes.c_longlong, ctypes.c_void_p]),
    'axs_reduce': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 6 + [ctypes.c_int] +
                   [ctypes.c_void_p] * 2 + [ctypes.c_longlong, ctypes.c_void_p]),
    'axs_hqr': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 3 + [ctypes.c_longlong, ctypes.c_void_p]),
    'axs_modes': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 2 + [ctypes.c_int] +
                  [ctypes.c_void_p] * 3 + [ctypes.c_longlong, ctypes.c_void_p]),
    'axs_get_reduction': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 4),
    'axs_hqr_stats': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 3),
    'axs_set_tuning': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    'axs_get_tuning': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_void_p]),
    'axs_track_worksize': (ctypes.c_longlong, [ctypes.c_int, ctypes.c_int]),
    'axs_track_pairs': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 3 + [ctypes.c_int] +
                        [ctypes.c_void_p] * 3 + [ctypes.c_longlong, ctypes.c_void_p]),
    'axs_track_scan': (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 3 + [ctypes.c_void_p]),
    'axs_apply_order': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 6),
    'axs_copy_to_pinned': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]),
    'axs_select_modes': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 2 + [ctypes.c_int] +
                         [ctypes.c_void_p] * 3),
    'axs_energy_ratio': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 2 +
                         [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 5),
    'axs_quadratic_forms': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 2 + [ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int] + [ctypes.c_void_p] * 6),
    'axs_modal_projection': (ctypes.c_int, [ctypes.c_int] * 2 + [ctypes.c_void_p] * 4),
}
EXPORTED = tuple(_SIGNATURES)


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('axisafe_b200: CUDA library not built (%s). Run '
                               '`python -c "import __graft_entry__ as g; g.build()"` '
                               'or axisym-safe-python_b200/build.py; there is no CPU fallback.' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def torch_cuda():
    """Return the torch module after checking that a CUDA device is usable."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('axisafe_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch


def bind_host_to_device_numa(device_index=None):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (one process per GPU).

    The sharded sweep downloads 0.5 GB of mode shapes per rank into pinned host memory; pinned pages are
    placed on the node of the CPU that allocates them, so a rank that runs on the other socket pushes its
    whole download across the inter-socket link.  Reads /sys/bus/pci/devices/<bdf>/numa_node and the node's
    cpulist; does nothing when the platform exposes no NUMA information.  Returns the node or None."""
    torch = torch_cuda()
    try:
        dev = torch.cuda.current_device() if device_index is None else int(device_index)
        p = torch.cuda.get_device_properties(dev)
        bdf = '%04x:%02x:%02x.0' % (getattr(p, 'pci_domain_id', 0), p.pci_bus_id, p.pci_device_id)
        with open('/sys/bus/pci/devices/%s/numa_node' % bdf) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError):
        return None


def check(code, what):
    if code != 0:
        msg = lib().axs_strerror(code).decode()
        raise RuntimeError('%s failed: %s (code %d)' % (what, msg, code))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(array, torch, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(array))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def download(tensor):
    """Device -> host copy into freshly allocated PINNED memory; returns a NumPy view of it.

    torch's caching host allocator recycles the pinned block once the array is dropped, so
    repeated sweeps do not pay cudaHostAlloc again (psi_hat is 0.5 GB at 2000 x 126 x 126).
    """
    torch = torch_cuda()
    if tensor.numel() * tensor.element_size() < (1 << 20):
        return tensor.cpu().numpy()
    stage = torch.empty(tensor.numel(), dtype=tensor.dtype, pin_memory=True)
    stage.copy_(tensor.reshape(-1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return stage.numpy()


# ------------------------------------------------------------------------ kernel 1a
class DeviceElements(object):
    """Element matrices resident on the GPU + the tables needed to scatter them."""

    def __init__(self, out, edesc_i, offsets, sizes):
        self.out, self.edesc_i, self.offsets, self.sizes = out, edesc_i, offsets, sizes


def element_matrices(elements):
    """Run axs_element_matrices for a list of element objects (elements.py classes).

    Installs the results on the objects (``_load``) and returns a DeviceElements.
    """
    from . import shape_functions as sf
    torch = torch_cuda()
    L = lib()
    tables, table_off, cache = [], {}, 0
    desc_i, desc_d, nodes, offsets, sizes = [], [], [], [], []
    node_off = out_off = 0
    for el in elements:
        (kind, family, pml, m), dbl = el._descriptor()
        if m > 64:
            raise ValueError('element with %d nodes exceeds the kernel limit of 64' % m)
        key = (family, m)
        if key not in table_off:
            ksi, w = (sf.build_GLJ_quadrature if family == 1 else sf.build_GLL_quadrature)(m)
            D = (sf.lagrange_GLJ if family == 1 else sf.lagrange_poly)(ksi)[1]
            table_off[key] = cache
            block = np.concatenate([ksi, w, D.ravel()])
            tables.append(block)
            cache += len(block)
        nd = (3 if kind == 0 else 1) * m
        desc_i.append([kind, family, pml, m, table_off[key], node_off, out_off, 0])
        desc_d.append(list(dbl) + [0.0] * (12 - len(dbl)))
        nodes.append(np.asarray(el.nodes_at, dtype=np.float64))
        offsets.append(out_off)
        sizes.append(nd)
        node_off += m
        out_off += 10 * nd * nd
    d_i = to_device(np.array(desc_i, dtype=np.int32), torch)
    d_d = to_device(np.array(desc_d, dtype=np.float64), torch)
    d_t = to_device(np.concatenate(tables), torch)
    d_n = to_device(np.concatenate(nodes), torch)
    out = torch.empty(out_off, dtype=torch.complex128, device='cuda')
    check(L.axs_element_matrices(len(elements), _ptr(d_i), _ptr(d_d), _ptr(d_t), _ptr(d_n), _ptr(out),
                                 _stream(torch)), 'axs_element_matrices')
    host = out.cpu().numpy()
    for el, off, nd in zip(elements, offsets, sizes):
        el._load(host[off:off + 10 * nd * nd].reshape(10, nd, nd).copy())
    return DeviceElements(out, d_i, offsets, sizes)


# ------------------------------------------------------------------------ kernel 1b
class DeviceOperators(object):
    """Banded operators of the quadratic pencil, resident in HBM."""

    def __init__(self, N, hb, K0s, Cb, Mb, K1c, K6d):
        self.N, self.hb = N, hb
        self.K0s, self.Cb, self.Mb, self.K1c, self.K6d = K0s, Cb, Mb, K1c, K6d

    def band_to_dense(self, t):
        """Host dense N x N copy of one banded operator (tests / debugging)."""
        W = 2 * self.hb + 1
        band = t.cpu().numpy().reshape(self.N, W)
        dense = np.zeros([self.N, self.N], complex)
        for r in range(self.N):
            for c in range(max(0, r - self.hb), min(self.N, r + self.hb + 1)):
                dense[r, c] = band[r, c - r + self.hb]
        return dense


def assemble_operators(mesh, order, dev_elem):
    """Scatter the element matrices into the banded operators (axs_assemble_operators)."""
    torch = torch_cuda()
    L = lib()
    N = mesh.no_of_dofs
    lm_all, lm_off, hb = [], [], 0
    for el in order:
        obj = mesh.elements[el]
        reduced = iter(mesh.LM[el])
        full = [-1 if i in obj._dropped else next(reduced) for i in range(obj._full_dofs)]
        live = [g for g in full if g >= 0]
        if live:
            hb = max(hb, max(live) - min(live))
        lm_off.append(len(lm_all))
        lm_all.extend(full)
    K1c = None
    if mesh.is_coupled and mesh.K1_coupling is not None and mesh.K1_coupling.nnz:
        coo = mesh.K1_coupling.tocoo()
        hb = max(hb, int(np.max(np.abs(coo.row - coo.col))))
    W = 2 * hb + 1
    if mesh.is_coupled and mesh.K1_coupling is not None and mesh.K1_coupling.nnz:
        band = np.zeros([N, W], complex)
        np.add.at(band, (coo.row, coo.col - coo.row + hb), coo.data)
        K1c = to_device(band, torch)
    tclass = np.array([0 if t == 1 else 1 for t in mesh.Tdiag], dtype=np.int32)
    d_lm = to_device(np.array(lm_all, dtype=np.int32), torch)
    d_off = to_device(np.array(lm_off, dtype=np.int32), torch)
    d_tc = to_device(tclass, torch)
    K0s = torch.zeros(N * W, dtype=torch.complex128, device='cuda')
    Cb = torch.zeros_like(K0s)
    Mb = torch.zeros_like(K0s)
    K6d = torch.zeros(N, dtype=torch.complex128, device='cuda')
    check(L.axs_assemble_operators(len(order), _ptr(dev_elem.edesc_i), _ptr(dev_elem.out), _ptr(d_lm),
                                   _ptr(d_off), _ptr(d_tc), int(mesh.n), N, hb, _ptr(K0s), _ptr(Cb),
                                   _ptr(Mb), _ptr(K6d), _stream(torch)), 'axs_assemble_operators')
    return DeviceOperators(N, hb, K0s, Cb, Mb, K1c, K6d)


# ------------------------------------------------------------------- kernels 1c, 2, 3
def build_S(ops, omega):
    """Dense S(w) batch [nf, n2, n2] (column-major per matrix) as a device tensor."""
    torch = torch_cuda()
    n2 = 2 * ops.N
    S = torch.empty(len(omega) * n2 * n2, dtype=torch.complex128, device='cuda')
    check(lib().axs_build_S(ops.N, ops.hb, _ptr(ops.K0s), _ptr(ops.Cb), _ptr(ops.Mb), _ptr(ops.K1c),
                            _ptr(ops.K6d), _ptr(omega), len(omega), _ptr(S), _stream(torch)), 'axs_build_S')
    return S


class EigWorkspace(object):
    def __init__(self, N, nf):
        torch = torch_cuda()
        self.N, self.nf = N, nf
        self.bytes = int(lib().axs_eig_worksize(N, nf)) + 8192      # slack for chunked sub-workspaces
        self.buf = torch.empty(self.bytes, dtype=torch.uint8, device='cuda')


_SIDE_STREAMS = []


def eig_sweep(ops, omega, want_modes=True, work=None, chunks=2):
    """axs_eig_sweep: returns (k [nf, n2], psi [nf, n2, n2] or None, info [nf], workspace).

    The frequency batch is cut into ``chunks`` contiguous groups that run on separate CUDA
    streams: the kernels of one group (L2-bound reduction, latency-bound QR, ...) fill the gaps
    and wave tails of the other (measured +5 % on B200).  Results land in the same buffers.
    """
    torch = torch_cuda()
    nf, n2 = len(omega), 2 * ops.N
    work = work or EigWorkspace(ops.N, nf)
    k = torch.empty(nf * n2, dtype=torch.complex128, device='cuda')
    psi = torch.empty(nf * n2 * n2, dtype=torch.complex128, device='cuda') if want_modes else None
    info = torch.zeros(nf, dtype=torch.int32, device='cuda')
    chunks = max(1, min(int(chunks), nf // 64 if nf >= 128 else 1))
    L = lib()
    if n2 > L.axs_max_small_n2():
        chunks = 1                       # large path: one CTA per frequency already fills the GPU

    def launch(lo, hi, buf, nbytes):
        check(L.axs_eig_sweep(ops.N, ops.hb, _ptr(ops.K0s), _ptr(ops.Cb), _ptr(ops.Mb), _ptr(ops.K1c),
                              _ptr(ops.K6d), _ptr(omega[lo:hi]), hi - lo, _ptr(k[lo * n2:hi * n2]),
                              _ptr(psi[lo * n2 * n2:hi * n2 * n2]) if psi is not None else None,
                              _ptr(info[lo:hi]), _ptr(buf), nbytes, _stream(torch)), 'axs_eig_sweep')
    if chunks == 1:
        launch(0, nf, work.buf, work.bytes)
        return k, psi, info, work
    while len(_SIDE_STREAMS) < chunks:
        _SIDE_STREAMS.append(torch.cuda.Stream())
    main = torch.cuda.current_stream()
    bounds = [nf * c // chunks for c in range(chunks + 1)]
    offset = 0
    for c in range(chunks):
        lo, hi = bounds[c], bounds[c + 1]
        need = int(L.axs_eig_worksize(ops.N, hi - lo))
        sub = work.buf[offset:offset + need]
        offset += (need + 255) // 256 * 256
        if offset > work.bytes + 256 * chunks:
            raise RuntimeError('workspace too small for the chunked sweep')
        side = _SIDE_STREAMS[c]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            launch(lo, hi, sub, need)
    for c in range(chunks):
        main.wait_stream(_SIDE_STREAMS[c])
    return k, psi, info, work


def eig_sweep_into(ops, omega, k, psi, info, work):
    """axs_eig_sweep into caller-provided device buffers (large path, chunked sweeps)."""
    torch = torch_cuda()
    check(lib().axs_eig_sweep(ops.N, ops.hb, _ptr(ops.K0s), _ptr(ops.Cb), _ptr(ops.Mb), _ptr(ops.K1c),
                              _ptr(ops.K6d), _ptr(omega), len(omega), _ptr(k), _ptr(psi), _ptr(info),
                              _ptr(work.buf), work.bytes, _stream(torch)), 'axs_eig_sweep')


def reduce_dense(S, N, nf, work=None):
    """axs_reduce on an explicit dense batch (parity tests)."""
    torch = torch_cuda()
    work = work or EigWorkspace(N, nf)
    check(lib().axs_reduce(N, 0, None, None, None, None, None, None, nf, _ptr(S), _ptr(work.buf),
                           work.bytes, _stream(torch)), 'axs_reduce')
    return work


def hqr(N, nf, work):
    torch = torch_cuda()
    n2 = 2 * N
    k = torch.empty(nf * n2, dtype=torch.complex128, device='cuda')
    info = torch.zeros(nf, dtype=torch.int32, device='cuda')
    check(lib().axs_hqr(N, nf, _ptr(k), _ptr(info), _ptr(work.buf), work.bytes, _stream(torch)), 'axs_hqr')
    return k, info


def modes(ops, nf, k, work):
    torch = torch_cuda()
    n2 = 2 * ops.N
    psi = torch.empty(nf * n2 * n2, dtype=torch.complex128, device='cuda')
    check(lib().axs_modes(ops.N, ops.hb, _ptr(ops.Cb), _ptr(ops.K6d), nf, _ptr(k), _ptr(psi),
                          _ptr(work.buf), work.bytes, _stream(torch)), 'axs_modes')
    return psi


def set_tuning(name, value):
    """axs_set_tuning: kernel-selection switch (see include/axisafe_b200.h); returns the old value."""
    old = ctypes.c_int(0)
    check(lib().axs_get_tuning(name.encode(), ctypes.addressof(old)), 'axs_get_tuning(%s)' % name)
    check(lib().axs_set_tuning(name.encode(), int(value)), 'axs_set_tuning(%s)' % name)
    return old.value


def hqr_stats(N, nf, work):
    """axs_hqr_stats: host array [nf, 16] per matrix = macro steps, batches, AED calls, AED deflations, then
    for each of the four cascade stages the kilo-cycles thread 0 spent in the AED window, in its
    application and in the sweeps."""
    torch = torch_cuda()
    st = torch.zeros(nf * 16, dtype=torch.int32, device='cuda')
    check(lib().axs_hqr_stats(N, nf, _ptr(work.buf), _ptr(st), _stream(torch)), 'axs_hqr_stats')
    return st.cpu().numpy().reshape(nf, 16)


def get_reduction(N, nf, work):
    torch = torch_cuda()
    n2 = 2 * N
    scale = torch.empty(nf * n2, dtype=torch.float64, device='cuda')
    H = torch.empty(nf * n2 * n2, dtype=torch.complex128, device='cuda')
    check(lib().axs_get_reduction(N, nf, _ptr(work.buf), _ptr(scale), _ptr(H), _stream(torch)),
          'axs_get_reduction')
    return scale.cpu().numpy().reshape(nf, n2), H.cpu().numpy().reshape(nf, n2, n2).transpose(0, 2, 1)


def track_pairs(ops, psi, npairs, scratch=None):
    """axs_track_pairs: (arg [npairs, n2] int32, amax [npairs, n2] float64) device tensors."""
    torch = torch_cuda()
    n2 = 2 * ops.N
    need = int(lib().axs_track_worksize(ops.N, npairs))
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device='cuda')
    arg = torch.empty(npairs * n2, dtype=torch.int32, device='cuda')
    amax = torch.empty(npairs * n2, dtype=torch.float64, device='cuda')
    check(lib().axs_track_pairs(ops.N, ops.hb, _ptr(ops.Cb), _ptr(ops.K6d), _ptr(psi), npairs, _ptr(arg),
                                _ptr(amax), _ptr(scratch), need, _stream(torch)), 'axs_track_pairs')
    return arg, amax


def track_pairs_into(ops, psi, npairs, arg, amax, scratch):
    """axs_track_pairs into caller-provided tables (pipelined sweep)."""
    torch = torch_cuda()
    need = int(lib().axs_track_worksize(ops.N, npairs))
    if scratch.numel() * scratch.element_size() < need:
        raise RuntimeError('tracking scratch too small')
    check(lib().axs_track_pairs(ops.N, ops.hb, _ptr(ops.Cb), _ptr(ops.K6d), _ptr(psi), npairs, _ptr(arg),
                                _ptr(amax), _ptr(scratch), need, _stream(torch)), 'axs_track_pairs')


def copy_to_pinned(dst_pinned, src_device):
    """axs_copy_to_pinned on the current stream: ``dst_pinned`` (contiguous pinned host tensor) <- ``src_device``
    (contiguous device tensor of the same byte size), stored by a kernel instead of the D2H copy engine."""
    torch = torch_cuda()
    nbytes = src_device.numel() * src_device.element_size()
    if nbytes != dst_pinned.numel() * dst_pinned.element_size() or not dst_pinned.is_pinned():
        raise ValueError('copy_to_pinned: size mismatch or destination not pinned')
    if nbytes == 0:
        return
    if not (src_device.is_contiguous() and dst_pinned.is_contiguous()):
        raise ValueError('copy_to_pinned: tensors must be contiguous')
    check(lib().axs_copy_to_pinned(_ptr(src_device), ctypes.c_void_p(dst_pinned.data_ptr()), nbytes, _stream(torch)),
          'axs_copy_to_pinned')


def fixup_row(prev, a, m, n2):
    """One tracking step with the reference's re-seating of weakly matched modes (solver.py:159-175):
    ``prev`` = column order at the previous frequency, ``a``/``m`` = arg-max / max tables of the pair.
    The very same Python set expression as the reference, so CPython's set iteration order is kept."""
    new_order = a[prev].copy()
    valid = np.round(m[prev], 2) > 0.9
    valid_entries = np.delete(new_order, np.where(valid == False))   # noqa: E712 (as the reference)
    unassigned = list(set(range(n2)) - set(valid_entries))
    j = 0
    for ii in range(n2):
        if not valid[ii]:
            new_order[ii] = unassigned[j]
            j += 1
    return new_order


def track_scan_range(arg, amax, order, start, stop, n2):
    """Rows [start, stop) of the reference column order (see track_scan); ``order`` [nf, n2] int32 is
    updated in place and must already hold rows < start.  ``arg``/``amax`` are the full host tables."""
    L = lib()
    while True:
        i = L.axs_track_scan(arg.ctypes.data, amax.ctypes.data, stop, n2, start, order.ctypes.data)
        if i >= stop:
            return
        order[i] = fixup_row(order[i - 1], arg[i - 1], amax[i - 1], n2)
        start = i + 1


def track_scan(arg, amax, nf, n2):
    """Reference column order per frequency (solver.py:159-182) from the arg-max tables.

    ``arg``/``amax`` are host arrays [nf-1, n2].  The hot loop is the C function
    axs_track_scan; frequencies where some tracked mode has round(|biorth|, 2) <= 0.9
    take the reference's fix-up with the very same Python set expression, so the
    (CPython-specific) iteration order of ``set`` is reproduced exactly.
    """
    arg = np.ascontiguousarray(arg, dtype=np.int32)
    amax = np.ascontiguousarray(amax, dtype=np.float64)
    order = np.zeros([nf, n2], dtype=np.int32)
    track_scan_range(arg, amax, order, 0, nf, n2)
    return order


def track_scan_segments(arg, amax, n2):
    """Relative tracking scan of one rank's frequency block (sharded sweep).

    ``arg``/``amax`` [P, n2]: the block's pairs; row j of the result belongs to the frequency j
    steps after the block's base frequency.  The scan o_j = a_j[o_{j-1}] is a composition of maps,
    so a block can be scanned from the identity without knowing the absolute column order sigma at
    its base: order_j = rel_j[sigma].  That holds up to the first frequency where a visited mode
    falls below the 0.9 threshold - there the reference re-seats modes by their *position*
    (fixup_row), which needs the absolute order - so the block is cut into segments at those rows
    and every segment restarts from the identity.  Returns (rel [P+1, n2] int32, starts): segment s
    covers rows starts[s] .. starts[s+1]-1 (starts[0] = 0; every later start is a fix-up row).
    A row that is clean for the identity start is clean for every sigma (the visited set can only
    shrink), so no fix-up is ever missed; a row flagged here but clean under sigma goes through
    fixup_row with all modes valid, which is the plain composition again.
    """
    arg = np.ascontiguousarray(arg, dtype=np.int32).reshape(-1, n2)
    amax = np.ascontiguousarray(amax, dtype=np.float64).reshape(-1, n2)
    P = arg.shape[0]
    rel = np.empty([P + 1, n2], dtype=np.int32)
    ident = np.arange(n2, dtype=np.int32)
    rel[0] = ident
    starts = [0]
    L = lib()
    start = 0
    while True:
        i = L.axs_track_scan(arg.ctypes.data, amax.ctypes.data, P + 1, n2, start, rel.ctypes.data)
        if i >= P + 1:
            return rel, starts
        rel[i] = ident
        starts.append(i)
        start = i + 1


def chain_segments(payloads, n2, sigma0=None, return_final=False):
    """Absolute column order at the start of every segment of every rank.

    ``payloads[r]`` = (comps [S_r, n2], d_arg [S_r - 1, n2], d_amax [S_r - 1, n2]) as produced from
    track_scan_segments: the last relative row of each segment and the pair tables of the fix-up
    row that ends it.  ``sigma0`` = absolute order at the base frequency of the first block
    (identity when the first block starts at frequency 0).  Returns ``sigma[r][s]`` (int32 [n2]) -
    and the order at the last frequency of the last block with ``return_final``; cost
    O(total segments * n2).
    """
    sigma = np.arange(n2, dtype=np.int32) if sigma0 is None else np.asarray(sigma0, dtype=np.int32)
    out = []
    for comps, d_arg, d_amax in payloads:
        mine = []
        S = comps.shape[0]
        for s in range(S):
            mine.append(sigma)
            prev = comps[s][sigma]                      # order at the last frequency of the segment
            sigma = fixup_row(prev, d_arg[s], d_amax[s], n2) if s < S - 1 else prev
        out.append(mine)
    return (out, sigma) if return_final else out


def select_modes(n2, nf, k, sigma, m, psi, rank=None):
    """axs_select_modes: rank of every native mode by |k - sigma| (device int32 [nf, n2]); the
    mode-shape columns of rank >= m are zeroed in ``psi`` (in place, may be None)."""
    torch = torch_cuda()
    if rank is None:
        rank = torch.empty(nf * n2, dtype=torch.int32, device='cuda')
    check(lib().axs_select_modes(n2, nf, _ptr(k), _ptr(sigma), int(m), _ptr(rank), _ptr(psi), _stream(torch)),
          'axs_select_modes')
    return rank


def apply_order(n2, nf, order, k, psi):
    torch = torch_cuda()
    d_order = order if hasattr(order, 'data_ptr') else to_device(np.ascontiguousarray(order, dtype=np.int32), torch)
    k_out = torch.empty_like(k)
    psi_out = torch.empty_like(psi) if psi is not None else None
    check(lib().axs_apply_order(n2, nf, _ptr(d_order), _ptr(k), _ptr(psi), _ptr(k_out), _ptr(psi_out),
                                _stream(torch)), 'axs_apply_order')
    return k_out, psi_out


def energy_ratio(N, nf, psi, tclass, total, pml):
    """axs_energy_ratio: ``total`` / ``pml`` are (rows, cols, vals) COO triplets (host arrays)."""
    torch = torch_cuda()

    def up(triplet):
        r, c, v = triplet
        return (to_device(np.asarray(r, np.int32), torch), to_device(np.asarray(c, np.int32), torch),
                to_device(np.asarray(v, np.complex128), torch), len(r))
    rt, ct, vt, nt = up(total)
    rp, cp, vp, npml = up(pml)
    d_tc = to_device(np.asarray(tclass, np.int32), torch)
    er = torch.empty(nf * 2 * N, dtype=torch.float64, device='cuda')
    check(lib().axs_energy_ratio(N, nf, _ptr(psi), _ptr(d_tc), nt, _ptr(rt), _ptr(ct), _ptr(vt),
                                 npml, _ptr(rp) if npml else None, _ptr(cp) if npml else None,
                                 _ptr(vp) if npml else None, _ptr(er), _stream(torch)), 'axs_energy_ratio')
    return er.cpu().numpy().reshape(nf, 2 * N)


def quadratic_forms(N, nf, psi, tclass, sel, matrices):
    """axs_quadratic_forms: ``matrices`` is a list of (rows, cols, vals) COO triplets (host arrays);
    returns the forms as a host array [nf, n_sel, n_mat] (complex)."""
    torch = torch_cuda()
    off = np.cumsum([0] + [len(m[0]) for m in matrices]).astype(np.int32)
    rows = np.concatenate([np.asarray(m[0], np.int32) for m in matrices] + [np.zeros(1, np.int32)])
    cols = np.concatenate([np.asarray(m[1], np.int32) for m in matrices] + [np.zeros(1, np.int32)])
    vals = np.concatenate([np.asarray(m[2], np.complex128) for m in matrices] + [np.zeros(1, np.complex128)])
    n_sel = 2 * N if sel is None else len(sel)
    d_sel = None if sel is None else to_device(np.asarray(sel, np.int32), torch)
    d_off, d_r, d_c, d_v = (to_device(a, torch) for a in (off, rows, cols, vals))
    d_tc = to_device(np.asarray(tclass, np.int32), torch)
    out = torch.empty(nf * n_sel * len(matrices), dtype=torch.complex128, device='cuda')
    check(lib().axs_quadratic_forms(N, nf, _ptr(psi), _ptr(d_tc), n_sel, _ptr(d_sel), len(matrices), _ptr(d_off),
                                    _ptr(d_r), _ptr(d_c), _ptr(d_v), _ptr(out), _stream(torch)),
          'axs_quadratic_forms')
    return out.cpu().numpy().reshape(nf, n_sel, len(matrices))


def modal_projection(N, nf, psi, fvec):
    """axs_modal_projection: psi^T fvec per frequency; returns a host array [nf, 2N]."""
    torch = torch_cuda()
    d_f = to_device(np.asarray(fvec, np.complex128), torch)
    out = torch.empty(nf * 2 * N, dtype=torch.complex128, device='cuda')
    check(lib().axs_modal_projection(N, nf, _ptr(psi), _ptr(d_f), _ptr(out), _stream(torch)),
          'axs_modal_projection')
    return out.cpu().numpy().reshape(nf, 2 * N)
