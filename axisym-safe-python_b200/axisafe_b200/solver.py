"""Per-frequency SAFE dispersion sweep on the GPU.

Host-side mirror of the reference's ``axisafe/solver.py`` ``WaveElementAxisym``
(same constructor, ``solve(f, no_of_waves='full', central_wavespeed=0)``, result
attributes ``k, psi_hat, w, B, T, K2_hat..., evaluated, k_ready``) with the loop body
of solver.py:119-182 replaced by batched sm_100a kernels:

    reference (per frequency, serial)                     here (all frequencies at once)
    ---------------------------------------------------   --------------------------------
    A(w), S = Binv A          solver.py:121-134           k_reduce (S built in-kernel)
    scipy.linalg.eig (zgeev)  solver.py:136               k_reduce + k_hqr + k_modes
    vec^T B vec normalisation solver.py:142               fused into k_modes
    biorth = psi^T B psi, argmax   solver.py:157-161      k_bpsi + k_track_mma
    permutation bookkeeping   solver.py:159-182           axs_track_scan (host, integers)
    reorder val / vec         solver.py:176-182           k_apply_order

The frequency axis is embarrassingly parallel; under ``torch.distributed``
(one process per GPU) every rank solves a contiguous block of frequencies plus a
one-frequency halo for the tracking pair at the block boundary, and only the
eigenvalue / arg-max tables are all-gathered (NCCL).  No CPU fallback.
"""
from __future__ import print_function

import os
import warnings

import numpy as np
import scipy.sparse as sps


class _LazyModes(object):
    """Device-resident ``psi_hat`` block that is copied to the host on first use."""

    def __init__(self, tensor, shape, first, total, ncols=None):
        self.tensor, self.shape, self.first, self.total = tensor, shape, first, total
        self.ncols = ncols               # subset mode: only the first ncols columns are modes
        self.host = None

    def materialise(self):
        if self.host is None:
            from . import _native
            # single GPU: the full [nf, 2N, 2N] array of the reference.  Sharded sweep: only this
            # rank's frequencies [first, first + shape[0]) exist here (see local_range); the full
            # array (0.5 GB .. TBs) is never gathered.
            self.host = _native.download(self.tensor).reshape(self.shape)
            if self.ncols is not None:
                self.host = np.ascontiguousarray(self.host[:, :, :self.ncols])
        return self.host


class WaveElementAxisym(object):
    """Axisymmetric wave element: dispersion curves of a layered waveguide (ref solver.py:24-55)."""

    def __init__(self, mesh):
        self.Mesh = mesh
        self.evaluated = False
        self.k_ready = None
        self.k = None
        self._psi = None
        self.er = None
        self.K2_hat = self.K3_hat = self.KF_hat = self.KFt_hat = None
        self.T = None
        self.B = None
        self.w = None
        self.e_p = self.e_k = self.p = self.en_vel = None
        self.propagating_indices = None
        self.info = None                 # per-frequency count of unconverged eigenvalues
        self.order = None                # tracked column order per frequency
        self._k_native = None            # eigenvalues before tracking (native order)
        self.local_range = None          # (first, last+1) frequencies solved by this rank (contiguous block)
        self.local_index = None          # global indices of the frequencies whose mode shapes this rank holds
        self._ncols = None               # subset mode: number of kept waves (device arrays stay 2N wide)
        self._sharded = False            # the last solve() was sharded over torch.distributed ranks
        self.timings = {}

    # psi_hat lives on the GPU until somebody reads it ------------------------------------
    @property
    def psi_hat(self):
        if isinstance(self._psi, _LazyModes):
            return self._psi.materialise()
        return self._psi

    @psi_hat.setter
    def psi_hat(self, value):
        self._psi = value

    @property
    def k_native(self):
        """Untracked spectra [nf, 2N] in the solver's native column order (a sharded sweep keeps the
        gathered table on the device until it is read)."""
        if self._k_native is not None and not isinstance(self._k_native, np.ndarray):
            self._k_native = self._k_native.cpu().numpy()
        return self._k_native

    @k_native.setter
    def k_native(self, value):
        self._k_native = value

    @property
    def psi_hat_device(self):
        """Tracked mode shapes of this rank's frequency block as a CUDA tensor [nf_local, 2N, 2N]."""
        if isinstance(self._psi, _LazyModes):
            return self._psi.tensor.view(self._psi.shape)
        return None

    def t_transform(self):
        """K_hat = T K conj(T) / (-i) for the skew matrices (ref solver.py:59-74)."""
        T = sps.diags(self.Mesh.Tdiag, offsets=0)
        hat = lambda K: T.dot(K).dot(T.conj()) / (-1j)
        self.K2_hat, self.K3_hat = hat(self.Mesh.K2), hat(self.Mesh.K3)
        self.KF_hat, self.KFt_hat = hat(self.Mesh.KF), hat(self.Mesh.KFt)
        self.T = T

    # ------------------------------------------------------------------------- the sweep
    def solve(self, f, no_of_waves='full', central_wavespeed=0, distributed='auto', pipeline=True):
        """All wavenumbers and mode shapes at every frequency of ``f`` (ref solver.py:76-184).

        ``no_of_waves`` = m (the reference's sparse shift-invert mode, solver.py:137-140) keeps the m
        wavenumbers nearest to ``w / central_wavespeed`` per frequency, tracked among themselves:
        ``k`` is [nf, m], ``psi_hat`` [nf, 2N, m].  The GPU computes the whole spectrum anyway, so the
        subset is a selection (axs_select_modes), not an Arnoldi iteration.
        ``distributed``: 'auto' shards the frequency axis over the ranks of an initialised
        ``torch.distributed`` process group; False forces a single-GPU sweep.
        Non-finite frequencies raise ``ValueError`` (SciPy's ``check_finite``), QR non-convergence raises
        ``numpy.linalg.LinAlgError`` - the two errors the reference's ``la.eig`` call can raise.
        ``pipeline``: single-GPU sweeps of >= 128 points run as a 4-chunk pipeline that copies the
        tracked ``psi_hat`` to (pinned) host memory while the next chunk computes; False keeps the
        mode shapes on the device until ``psi_hat`` is read.
        """
        from . import _native
        print('Calculating dispersion curves...')
        if not np.isfinite(np.asarray(f, dtype=float)).all():
            # what scipy.linalg.eig(check_finite=True) raises on the reference's S(w), solver.py:136
            raise ValueError('array must not contain infs or NaNs')
        subset = None
        if not (isinstance(no_of_waves, str) and no_of_waves == 'full'):
            # the reference's shift-invert mode (solver.py:137-140): the no_of_waves eigenvalues nearest
            # to sigma = w / central_wavespeed.  Here: full spectrum on the GPU, then a selection.
            subset = int(no_of_waves)
            if subset < 1 or subset > 2 * self.Mesh.no_of_dofs - 2:
                raise ValueError('no_of_waves must be between 1 and 2N - 2 (ARPACK: k < n - 1)')
            if central_wavespeed == 0:
                raise ZeroDivisionError('central_wavespeed must be non-zero for a subset solution')
            distributed = False                          # every rank solves the whole (small) sweep
        torch = _native.torch_cuda()
        mesh = self.Mesh
        ops = mesh._device_ops
        if ops is None:
            raise RuntimeError('Mesh.assemble_matrices(n) must be called before solve()')
        n = mesh.n
        f = np.asarray(f, dtype=float)
        self.w = 2 * np.pi * f
        N, n2, nf = mesh.no_of_dofs, 2 * mesh.no_of_dofs, len(f)

        def host_attributes():
            # the reference's sparse by-products (solver.py:102-117): T, K*_hat, B.  The GPU path does
            # not need them (the device operators were assembled with the mesh), so the pipelined
            # sweep builds them while the first chunks compute.
            self.t_transform()
            self.B = sps.bmat([[None, mesh.K6.tocsr()],
                               [mesh.K6.tocsr(), n * mesh.K4.tocsr() + self.K3_hat.tocsr()]])
        if n2 > _native.lib().axs_max_n2():
            raise NotImplementedError('2N = %d exceeds the largest supported problem (2N <= %d)'
                                      % (n2, _native.lib().axs_max_n2()))

        rank, world = 0, 1
        if distributed in ('auto', True):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                rank, world = dist.get_rank(), dist.get_world_size()
            if nf < 2 * world:
                # too few points to shard (a rank with an empty block would leave the others waiting in
                # a collective): decided from nf and the world size alone, so identically on every rank -
                # every rank solves the whole sweep
                rank, world = 0, 1
        self._sharded = world > 1
        bounds = [(nf * r) // world for r in range(world + 1)]
        lo, hi = bounds[rank], bounds[rank + 1]
        halo = 1 if lo > 0 else 0
        self.local_range = (lo, hi)

        small = n2 <= _native.lib().axs_max_small_n2()
        self._ncols = subset
        if subset is not None:
            _solve_subset(self, torch, _native, ops, n2, subset, central_wavespeed, small, host_attributes)
            self.evaluated = True
            self.k_ready = 0
            return
        self.local_index = np.arange(lo, hi)
        if small and world > 1 and pipeline and nf // world >= 128 and os.environ.get('AXS_SHARD_PIPE', '1') != '0':
            # sharded + pipelined: interleaved pieces, order fixed chunk by chunk (see the function)
            _solve_sharded_pipelined(self, torch, _native, ops, nf, n2, rank, world, host_attributes)
            self.evaluated = True
            self.k_ready = 0
            return
        if small and world == 1 and pipeline and nf >= 128:
            # single GPU, shared-memory kernels: chunk pipeline (compute of chunk c+1 overlaps the
            # host scan, the column gather and the psi_hat download of chunk c)
            _solve_pipelined(self, torch, _native, ops, nf, n2, host_attributes)
            self.evaluated = True
            self.k_ready = 0
            return
        if not small:
            host_attributes()
        # host -> device: this rank's angular frequencies (with the halo point in front)
        omega = torch.from_numpy(self.w[lo - halo:hi].copy()).cuda()
        nloc = hi - lo + halo
        if not small:
            # large path: matrices stream from HBM, the sweep is cut into memory-sized chunks
            k_nat, psi_nat, info, arg, amax = _sweep_large(torch, _native, ops, omega, halo)
        else:
            k_nat, psi_nat, info, work = _native.eig_sweep(ops, omega, want_modes=True)
            if nloc > 1:
                arg, amax = _native.track_pairs(ops, psi_nat, nloc - 1, scratch=work.buf)
            else:
                arg = torch.empty(0, dtype=torch.int32, device='cuda')
                amax = torch.empty(0, dtype=torch.float64, device='cuda')
        k_loc = k_nat.view(nloc, n2)[halo:]
        info_loc = info[halo:]
        if world > 1:
            # sharded sweep: every rank scans and reorders its own block, O(nf / world) host work
            if small:
                host_attributes()                        # while the kernels above run
            order_loc, order_dev = _track_sharded(self, torch, _native, bounds, rank, n2, k_loc,
                                                  arg.view(-1, n2), amax.view(-1, n2), info_loc)
        else:
            if small:
                host_attributes()                        # while the kernels above run
            # device -> host: the small tables
            k_host = k_nat.view(nloc, n2).cpu().numpy()
            arg_h = arg.cpu().numpy().reshape(max(nf - 1, 0), n2)
            amax_h = amax.cpu().numpy().reshape(max(nf - 1, 0), n2)
            self.info = info_loc.cpu().numpy()
            _raise_unconverged(self.info)
            self.k_native = k_host       # untracked spectra, solver's native column order
            self.order = _native.track_scan(arg_h, amax_h, nf, n2)
            self.k = np.take_along_axis(k_host, self.order.astype(np.int64), axis=1)
            order_loc = order_dev = self.order
        # reorder this rank's mode shapes on the device; they are copied out lazily
        if psi_nat is None:
            # the mode shapes of the whole block do not fit in HBM (the reference's psi_hat would be
            # nf x 2N x 2N x 16 B of host memory): eigenvalues and tracking are complete, shapes are not kept
            warnings.warn('solve(): the mode shapes of %d frequencies x 2N = %d (%.1f GB) do not fit into the free '
                          'device memory; k is complete, psi_hat is None - solve fewer frequencies per call '
                          'to keep the shapes' % (hi - lo, n2, (hi - lo) * n2 * n2 * 16 / 1e9), RuntimeWarning)
            self._psi = None
        elif not small:
            self._psi = _LazyModes(_reorder_in_place(torch, _native, n2, order_loc, psi_nat, halo),
                                   (hi - lo, n2, n2), lo, nf)
        else:
            psi_loc = psi_nat.view(nloc, n2 * n2)[halo:].reshape(-1)
            _, psi_sorted = _native.apply_order(n2, hi - lo, order_dev,
                                                k_nat.view(nloc, n2)[halo:].reshape(-1), psi_loc)
            self._psi = _LazyModes(psi_sorted, (hi - lo, n2, n2), lo, nf)
        self.evaluated = True
        self.k_ready = 0


    # ------------------------------------------------------------- mode-type filtering (row N1)
    def energy_ratio(self, core_sets=None):
        """Kinetic-energy ratio PML / whole waveguide per mode (ref solver.py:186-269).

        The reference forms q = conj(T) psi_hat[:, N:, :] on the host and two dense batched
        products with the (sub-)assembled mass matrices; here the tracked mode shapes stay on the
        GPU (axs_energy_ratio) and only ``er`` [nf, 2N] comes back.
        """
        if not self.evaluated:
            print('No solution is availble. Solve the system first.')
            return
        from . import _native
        print('Calculating the energy ratio...')
        mesh = self.Mesh
        pml_elements = []
        for key in mesh.dofs_per_elset:
            if 'PML' in key:
                pml_elements.extend(mesh.element_sets[key])
        M_pml = mesh.assemble_subset(sorted(set(pml_elements)))['M'].tocoo() if pml_elements else None
        M_tot = mesh.M.tocoo()
        mult = np.ones(mesh.no_of_dofs)
        if mesh.dof_domains['fluid'] != []:
            mult[mesh.dof_domains['fluid']] = -1             # acts on columns (ref solver.py:234-239)

        def triplets(coo):
            if coo is None:
                return (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, complex))
            coo.sum_duplicates()
            keep = coo.data != 0
            return coo.row[keep], coo.col[keep], (coo.data * mult[coo.col])[keep]
        tclass = [0 if t == 1 else 1 for t in mesh.Tdiag]
        loc = self.local_index
        if self.psi_hat_device is None:
            raise RuntimeError('mode shapes are no longer on the device; call solve() again')
        er_loc = _native.energy_ratio(mesh.no_of_dofs, len(loc), self.psi_hat_device.reshape(-1), tclass,
                                      triplets(M_tot), triplets(M_pml))
        if self._ncols is not None:
            er_loc = er_loc[:, :self._ncols]                  # the padding columns are not modes
        # sharded sweep: the [nf, 2N] table is completed from the other ranks (it is as small as k), so
        # that k_propagating() / energy_velocity() see the same er on every rank as a single-GPU sweep
        self.er = _complete_rows(self, er_loc)

    def k_propagating(self, imag_threshold=10, ER_threshold=0.9):
        """Positive-going propagating solutions (ref solver.py:271-299); host, O(nf * 2N)."""
        if self.er is None:
            print('Warning. Energy ratio must be evaluated first. ' + 'Use the trace_back method')
            return
        k = np.copy(self.k)
        k[abs(k.imag) > imag_threshold] = 0
        if ER_threshold != 0:
            k[self.er > ER_threshold] = 0
        k[k.real < 0] = 0
        k[k.imag > .1] = 0
        self.propagating_indices = np.where(~(k == 0).all(0))[0]
        self.k_ready = k[:, self.propagating_indices]


    # ----------------------------------------------------------------- energy velocity (row N3)
    def energy_velocity(self, core_sets=None, only_propagating=True):
        """Kinetic / potential energy, power flow and energy velocity per mode (ref solver.py:301-409).

        The reference takes q = conj(T) psi_hat[:, N:, :] to the host and evaluates dense batched
        products with the core-subset matrices; here the nine quadratic forms q^H X q
        (X = M, K1, K2, K5, KF, KF^T, KFt^T, KFt, K6 of the non-PML elements, fluid columns x -1 as in
        solver.py:353-362) are evaluated on the GPU from the device-resident tracked mode shapes
        (axs_quadratic_forms) and only [nf, modes, 9] numbers come back; the combination with k, n, w
        (solver.py:395-409) is O(nf * modes) host work.  Sets ``e_k, e_p, p, en_vel``.

        ``core_sets``: the reference's handling of a user list is broken (it ends with empty
        subsets, solver.py:318-341); here the listed element sets are used, as its docstring says.
        ``only_propagating``: needs ``k_propagating()`` first (the reference's own check
        ``propagating_indices == []`` raises on current NumPy; the intent is kept).
        """
        if not self.evaluated:
            print('No solution is availble. Solve the system first.')
            return
        from . import _native
        mesh = self.Mesh
        pml_elements, core_elements = [], []
        for key in mesh.dofs_per_elset:
            if core_sets is None:
                (pml_elements if 'PML' in key else core_elements).extend(mesh.element_sets[key])
            elif key in core_sets:
                core_elements.extend(mesh.element_sets[key])
        core_elements = sorted(set(core_elements) - set(pml_elements))
        sub = mesh.assemble_subset(core_elements)
        mult = np.ones(mesh.no_of_dofs)
        if mesh.dof_domains['fluid'] != []:
            mult[mesh.dof_domains['fluid']] = -1

        def triplets(name, transpose=False):
            coo = sub[name].tocoo()
            coo.sum_duplicates()
            vals = coo.data * mult[coo.col]
            keep = vals != 0
            r, c = (coo.col, coo.row) if transpose else (coo.row, coo.col)
            return r[keep], c[keep], vals[keep]
        mats = [triplets('M'), triplets('K1'), triplets('K2'), triplets('K5'), triplets('KF'),
                triplets('KF', True), triplets('KFt', True), triplets('KFt'), triplets('K6')]
        if only_propagating:
            if self.propagating_indices is None or len(self.propagating_indices) == 0:
                self.k_propagating()
            if self.propagating_indices is None:
                raise RuntimeError('energy_velocity(only_propagating=True) needs energy_ratio() and '
                                   'k_propagating() first')
            sel = np.asarray(self.propagating_indices, dtype=np.int32)
        else:
            sel = None if self._ncols is None else np.arange(self._ncols, dtype=np.int32)
        if self.psi_hat_device is None:
            raise RuntimeError('mode shapes are no longer on the device; call solve() again')
        loc = self.local_index
        tclass = [0 if t == 1 else 1 for t in mesh.Tdiag]
        forms = _native.quadratic_forms(mesh.no_of_dofs, len(loc), self.psi_hat_device.reshape(-1), tclass, sel, mats)
        fM, fK1, fK2, fK5, fKF, fKFT, fKFtT, fKFt, fK6 = (forms[:, :, i] for i in range(9))
        k = self.k[loc] if sel is None else self.k[loc][:, sel]
        w = self.w[loc].reshape(-1, 1)
        n = mesh.n
        e_k = w ** 2 / 4 * fM
        e_p = 0.25 * (fK1 + 1j * n * fK2 + n ** 2 * fK5 + 1j * k.conj() * fKF - 1j * k * fKFT +
                      k * n * fKFtT + k.conj() * n * fKFt + k * k.conj() * fK6)
        p = -w / 2 * np.imag(fKF - 1j * n * fKFt - 1j * k * fK6)

        # sharded sweep: completed from the other ranks, like er
        self.e_k, self.e_p, self.p = (_complete_rows(self, a) for a in (e_k, e_p, p))
        with np.errstate(divide='ignore', invalid='ignore'):
            self.en_vel = self.p / (self.e_k + self.e_p)


    # ------------------------------------------------------------------ forced response (row N4)
    def point_excited_waves(self, theta_0, r_f, f_ext, propagating=False):
        """Excited wave amplitudes for a point force (ref solver.py:411-435).

        ``psi_hat^T [0; f_ext]`` is evaluated on the GPU from the device-resident mode shapes
        (axs_modal_projection); returns [nf, 2N, 1] like the reference.  As in the reference, only
        ``propagating=False`` returns a value.
        """
        from . import _native
        if not propagating:
            N = self.Mesh.no_of_dofs
            force_vector = np.r_[np.zeros(N), np.asarray(f_ext).reshape(-1)]
            if self.psi_hat_device is None:
                raise RuntimeError('mode shapes are no longer on the device; call solve() again')
            loc = self.local_index
            proj = _native.modal_projection(N, len(loc), self.psi_hat_device.reshape(-1), force_vector)
            if self._ncols is not None:
                proj = proj[:, :self._ncols]
            proj = _complete_rows(self, proj)                # sharded sweep: the other ranks' frequencies
            return (1j * np.sinc(self.Mesh.n * theta_0 / np.pi) / (2 * np.pi * r_f) * proj)[:, :, np.newaxis]

    def calculate_frf(self, theta_0, r_f, f_ext, distance=0):
        """Modal contributions to the point-force FRF at ``distance`` (ref solver.py:437-474)."""
        idx = self.propagating_indices
        with np.errstate(over='ignore', invalid='ignore'):       # growing (non-propagating) columns overflow
            amps = self.point_excited_waves(theta_0, r_f, f_ext)[:, idx, 0] * np.exp(-1j * self.k[:, idx] * distance)
        # a sharded sweep holds the mode shapes of local_index only: the product is formed for those
        # frequencies and the [nf, N, modes] result completed from the other ranks
        loc = self.local_index if self.local_index is not None else np.arange(len(self.w))
        shapes = self.psi_hat[:, self.Mesh.no_of_dofs:, idx]
        return _complete_rows(self, shapes * amps[loc][:, np.newaxis, :])


def _complete_rows(sol, rows):
    """[len(local_index), ...] rows of this rank -> the full [nf, ...] host array on every rank.

    Single-GPU sweeps return ``rows`` unchanged.  After a sharded sweep every rank holds the mode
    shapes of ``local_index`` only, so per-mode results derived from them (energy ratio, energy
    velocity, modal amplitudes) are exchanged: one all-gather of the padded blocks + their global
    frequency indices over the default process group (device tensors for NCCL, host tensors for
    gloo).  Collective: every rank of the sweep must make the same call.
    """
    nf = len(sol.w)
    loc = sol.local_index
    if not sol._sharded or loc is None or len(loc) == nf:
        return rows
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    rows = np.ascontiguousarray(rows)
    tail, dtype = rows.shape[1:], rows.dtype
    flat = rows.reshape(len(loc), -1)
    as_real = np.ascontiguousarray(flat).view(np.float64) if np.iscomplexobj(flat) else flat.astype(np.float64, copy=False)
    width = as_real.shape[1]
    cnt = torch.tensor([len(loc)], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    size = max(cnts)
    pad = torch.zeros(size, width + 1, dtype=torch.float64, device=dev)
    if len(loc):
        pad[:len(loc), :width] = torch.from_numpy(as_real).to(dev)
        pad[:len(loc), width] = torch.from_numpy(np.asarray(loc, dtype=np.float64)).to(dev)   # exact below 2^53
    full = torch.empty(world * size, width + 1, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(full, pad)
    full = full.cpu().numpy()
    out = np.full([nf, width], np.nan)
    for r in range(world):
        blk = full[r * size:r * size + cnts[r]]
        out[blk[:, width].astype(np.int64)] = blk[:, :width]
    if np.issubdtype(dtype, np.complexfloating):
        out = out.view(np.complex128)
    return out.reshape((nf,) + tail).astype(dtype, copy=False)


_PIPE = {}


def _pipe_streams(torch):
    """Side streams of the chunk pipelines, one set per device (a process may drive several GPUs)."""
    dev = torch.cuda.current_device()
    if dev not in _PIPE:
        _PIPE[dev] = dict(cs=[torch.cuda.Stream(), torch.cuda.Stream()], xs=torch.cuda.Stream(), ss=torch.cuda.Stream())
    d = _PIPE[dev]
    return d['cs'], d['xs'], d['ss']


def _solve_pipelined(sol, torch, _native, ops, nf, n2, host_work=None):
    """Single-GPU sweep as a software pipeline over K frequency chunks.

    stream A/B (alternating)  reduce, QR, modes of chunk c; tracking pairs of chunk c (+ the seam
                              pair, after chunk c-1's modes on the other stream)
    stream X                  D2H of the small tables of chunk c -> [host: scan rows of chunk c]
                              -> H2D order, column gather (apply_order), D2H of psi_hat chunk c
    The host only ever waits for the small tables of chunk c AFTER the kernels of chunk c+1 are
    queued, so the scan, the gather and the 0.5 GB psi_hat download hide behind compute.
    Results are identical to the one-shot path (same kernels on the same inputs).
    """
    import os
    import time
    K = int(os.environ.get('AXS_PIPE_CHUNKS', 4 if nf >= 512 else 2))
    trace = os.environ.get('AXS_PIPE_TRACE')
    tl = [('start', time.perf_counter())]
    if os.environ.get('AXS_PIPE_BOUNDS'):               # e.g. "0.3,0.6,0.8,1" (A/B testing)
        fr = [float(x) for x in os.environ['AXS_PIPE_BOUNDS'].split(',')]
        bounds = [0] + [int(nf * x) for x in fr[:-1]] + [nf]
        K = len(bounds) - 1
    elif K == 4:
        # chunks run pairwise on two streams; the last pair is smaller so that the un-overlapped
        # tail (scan + gather + download of the final chunk) is short
        # (44 / 44 / 7 / 5 % rounded to whole waves of one CTA per SM: 888 / 888 / 148 / 76 points at
        # nf = 2000 on 148 SMs; within noise of 40 / 40 / 10 / 10 %: 61.5-62.7 ms end to end)
        unit = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        b1 = max(unit, int(round(0.444 * nf / unit)) * unit)
        b3 = 2 * b1 + max(unit, int(round(0.074 * nf / unit)) * unit)
        if nf >= 8 * unit and b3 < nf:
            bounds = [0, b1, 2 * b1, b3, nf]
        else:
            bounds = [0, int(nf * 0.4), int(nf * 0.8), int(nf * 0.9), nf]
    else:
        bounds = [nf * c // K for c in range(K + 1)]
    nn = n2 * n2
    cs, xs, ss = _pipe_streams(torch)
    main = torch.cuda.current_stream()
    cmax = max(bounds[c + 1] - bounds[c] for c in range(K))
    dev = 'cuda'
    omega = torch.from_numpy(sol.w.copy()).cuda()
    k_nat = torch.empty(nf * n2, dtype=torch.complex128, device=dev)
    psi_nat = torch.empty(nf * nn, dtype=torch.complex128, device=dev)
    psi_sorted = torch.empty(nf * nn, dtype=torch.complex128, device=dev)
    info = torch.zeros(nf, dtype=torch.int32, device=dev)
    arg = torch.empty((nf - 1) * n2, dtype=torch.int32, device=dev)
    amax = torch.empty((nf - 1) * n2, dtype=torch.float64, device=dev)
    order_d = torch.empty(nf * n2, dtype=torch.int32, device=dev)
    k_sorted = torch.empty(nf * n2, dtype=torch.complex128, device=dev)
    works = [_native.EigWorkspace(ops.N, cmax) for _ in range(2)]
    ybuf = [torch.empty(int(_native.lib().axs_track_worksize(ops.N, cmax)), dtype=torch.uint8, device=dev)
            for _ in range(2)]
    pin = lambda *shape, dtype: torch.empty(*shape, dtype=dtype, pin_memory=True)
    k_h = pin(nf, n2, dtype=torch.complex128)
    ks_h = pin(nf, n2, dtype=torch.complex128)
    arg_h = pin(max(nf - 1, 1), n2, dtype=torch.int32)
    amax_h = pin(max(nf - 1, 1), n2, dtype=torch.float64)
    info_h = pin(nf, dtype=torch.int32)
    order_h = pin(nf, n2, dtype=torch.int32)
    psi_h = pin(nf, n2, n2, dtype=torch.complex128)
    order = order_h.numpy()
    ev_modes = [torch.cuda.Event(enable_timing=bool(trace)) for _ in range(K)]
    ev_track = [torch.cuda.Event(enable_timing=bool(trace)) for _ in range(K)]
    ev_small = [torch.cuda.Event(enable_timing=bool(trace)) for _ in range(K)]
    ev_done = [torch.cuda.Event(enable_timing=bool(trace)) for _ in range(K)]
    ev0 = torch.cuda.Event(enable_timing=bool(trace))
    ev0.record(main)
    tl.append(('alloc', time.perf_counter()))
    for s_ in cs + [xs, ss]:
        s_.wait_stream(main)

    def compute(c):
        lo, hi = bounds[c], bounds[c + 1]
        st = cs[c % 2]
        with torch.cuda.stream(st):
            _native.eig_sweep_into(ops, omega[lo:hi], k_nat[lo * n2:hi * n2], psi_nat[lo * nn:hi * nn],
                                   info[lo:hi], works[c % 2])
            ev_modes[c].record(st)
            if c > 0:
                st.wait_event(ev_modes[c - 1])           # the seam pair needs the previous chunk's shapes
            p0 = max(lo - 1, 0)
            npairs = hi - 1 - p0
            if npairs > 0:
                _native.track_pairs_into(ops, psi_nat[p0 * nn:hi * nn], npairs, arg[p0 * n2:(hi - 1) * n2],
                                         amax[p0 * n2:(hi - 1) * n2], ybuf[c % 2])
            ev_track[c].record(st)
        with torch.cuda.stream(ss):                      # small tables on their own stream: waiting for the
            ss.wait_event(ev_track[c])                   # tracking of chunk c must not hold up psi downloads
            # (stored by a kernel through the unified address space: a cudaMemcpyAsync would queue behind the
            # bulk psi_hat download of the previous chunk on the D2H engine)
            _native.copy_to_pinned(k_h[lo:hi], k_nat[lo * n2:hi * n2])
            _native.copy_to_pinned(info_h[lo:hi], info[lo:hi])
            if hi - 1 > p0:
                _native.copy_to_pinned(arg_h[p0:hi - 1], arg[p0 * n2:(hi - 1) * n2])
                _native.copy_to_pinned(amax_h[p0:hi - 1], amax[p0 * n2:(hi - 1) * n2])
            ev_small[c].record(ss)

    def finish(c):
        lo, hi = bounds[c], bounds[c + 1]
        ev_small[c].synchronize()
        if info_h[lo:hi].numpy().any():
            torch.cuda.synchronize()
            bad = int(np.count_nonzero(info_h[lo:hi].numpy()))
            raise np.linalg.LinAlgError('eig algorithm (QR) did not converge at %d frequencies' % bad)
        _native.track_scan_range(arg_h.numpy(), amax_h.numpy(), order, lo, hi, n2)
        with torch.cuda.stream(xs):
            xs.wait_event(ev_track[c])
            order_d[lo * n2:hi * n2].copy_(order_h[lo:hi].view(-1), non_blocking=True)
            _native.check(_native.lib().axs_apply_order(
                n2, hi - lo, _native._ptr(order_d[lo * n2:hi * n2]), _native._ptr(k_nat[lo * n2:hi * n2]),
                _native._ptr(psi_nat[lo * nn:hi * nn]), _native._ptr(k_sorted[lo * n2:hi * n2]),
                _native._ptr(psi_sorted[lo * nn:hi * nn]), _native._stream(torch)), 'axs_apply_order')
            ks_h[lo:hi].copy_(k_sorted[lo * n2:hi * n2].view(hi - lo, n2), non_blocking=True)
            psi_h[lo:hi].copy_(psi_sorted[lo * nn:hi * nn].view(hi - lo, n2, n2), non_blocking=True)
            ev_done[c].record(xs)

    # AXS_PIPE_LAG=2: queue the kernels of chunk c+2 (same stream and workspace as chunk c, so in order
    # behind it) before the host blocks on the small tables of chunk c
    lag = max(1, int(os.environ.get('AXS_PIPE_LAG', 1)))      # 2 measured equal (62.0-62.7 ms either way)
    for c in range(K):
        compute(c)
        tl.append(('enq%d' % c, time.perf_counter()))
        if c == min(1, K - 1) and host_work is not None:
            host_work()                                  # sparse host by-products, hidden behind the kernels
            tl.append(('host', time.perf_counter()))
        if c >= lag:
            finish(c - lag)
            tl.append(('fin%d' % (c - lag), time.perf_counter()))
    for c in range(max(K - lag, 0), K):
        finish(c)
        tl.append(('fin%d' % c, time.perf_counter()))
    xs.synchronize()
    tl.append(('sync', time.perf_counter()))
    if trace:
        t0 = tl[0][1]
        print('AXS_PIPE host ms:', ' '.join('%s=%.1f' % (n_, (t_ - t0) * 1e3) for n_, t_ in tl), file=__import__('sys').stderr)
        print('AXS_PIPE device ms since start:', ' '.join(
            'c%d modes=%.1f track=%.1f small=%.1f done=%.1f' % (c, ev0.elapsed_time(ev_modes[c]), ev0.elapsed_time(ev_track[c]),
                                                             ev0.elapsed_time(ev_small[c]), ev0.elapsed_time(ev_done[c]))
            for c in range(K)), file=__import__('sys').stderr)
    main.wait_stream(xs)
    for s_ in cs:
        main.wait_stream(s_)
    sol.info = info_h.numpy().copy()
    sol.k_native = k_h.numpy()
    sol.order = order.copy()
    sol.k = ks_h.numpy()                                 # tracked on the device by k_apply_order
    sol.local_range = (0, nf)
    sol.local_index = np.arange(nf)
    modes = _LazyModes(psi_sorted, (nf, n2, n2), 0, nf)
    modes.host = psi_h.numpy()
    sol._psi = modes


def _subset_scan(arg, amax, rank, m):
    """Tracked native column per (frequency, wave) for the subset mode.

    The reference tracks the m columns ARPACK returned (solver.py:151-182): biorth is m x m,
    ``new_order`` holds *positions* in the current column list, and weakly matched modes are
    re-seated from ``set(range(m))``.  Here position p at frequency i is the selected mode of
    rank p (distance to sigma); ``arg``/``amax`` are the native tables of axs_track_pairs computed
    with the unselected columns zeroed, so ``rank[i][arg]`` is the position the reference's arg-max
    would return.  Returns cols [nf, m] (native column of every tracked wave).
    """
    from . import _native
    nf, n2 = rank.shape
    sel = np.argsort(rank, axis=1, kind='stable')[:, :m].astype(np.int32)     # native index by position
    cols = np.empty([nf, m], dtype=np.int32)
    cols[0] = sel[0]
    for i in range(1, nf):
        prev = cols[i - 1]
        pos = np.minimum(rank[i][arg[i - 1][prev]], m - 1)        # (an all-zero row cannot occur)
        if not (np.rint(amax[i - 1][prev] * 100.0) >= 91.0).all():
            ident = np.arange(m, dtype=np.int32)
            pos = _native.fixup_row(ident, pos.astype(np.int32), amax[i - 1][prev], m)
        cols[i] = sel[i][pos]
    return cols


def _solve_subset(sol, torch, _native, ops, n2, m, central_wavespeed, small, host_attributes):
    """solve(f, no_of_waves=m, central_wavespeed=c): whole spectrum on the GPU, the m wavenumbers
    nearest to w / c selected per frequency (axs_select_modes), tracked among themselves.

    Device arrays keep their [nf, 2N(, 2N)] shape with the tracked waves in the first m columns
    (the remaining mode-shape columns are zero), so the energy-ratio / energy-velocity / forced
    response kernels run unchanged; the host attributes are cut to m columns like the reference's.
    """
    nf = len(sol.w)
    sigma = torch.from_numpy(sol.w / float(central_wavespeed)).cuda()
    omega = torch.from_numpy(sol.w.copy()).cuda()
    if small:
        k_nat, psi_nat, info, work = _native.eig_sweep(ops, omega, want_modes=True)
        rank = _native.select_modes(n2, nf, k_nat, sigma, m, psi_nat)
        if nf > 1:
            arg, amax = _native.track_pairs(ops, psi_nat, nf - 1, scratch=work.buf)
        else:
            arg = torch.empty(0, dtype=torch.int32, device='cuda')
            amax = torch.empty(0, dtype=torch.float64, device='cuda')
        host_attributes()                                # while the kernels above run
    else:
        host_attributes()
        rank = torch.empty(nf * n2, dtype=torch.int32, device='cuda')
        k_nat, psi_nat, info, arg, amax = _sweep_large(torch, _native, ops, omega, 0, select=(sigma, m, rank))
    sol.info = info.cpu().numpy()
    _raise_unconverged(sol.info)
    k_host = k_nat.view(nf, n2).cpu().numpy()
    rank_h = rank.view(nf, n2).cpu().numpy()
    arg_h = arg.cpu().numpy().reshape(max(nf - 1, 0), n2)
    amax_h = amax.cpu().numpy().reshape(max(nf - 1, 0), n2)
    cols = _subset_scan(arg_h, amax_h, rank_h, m)
    # full-width order rows: the tracked waves first, then the unselected (zeroed) native columns
    rest = np.argsort(rank_h, axis=1, kind='stable')[:, m:].astype(np.int32)
    order = np.ascontiguousarray(np.concatenate([cols, rest], axis=1))
    sol.k_native = k_host
    sol.order = cols
    sol.k = np.take_along_axis(k_host, cols.astype(np.int64), axis=1)
    sol.local_range = (0, nf)
    sol.local_index = np.arange(nf)
    if psi_nat is None:
        sol._psi = None
    elif not small:
        sol._psi = _LazyModes(_reorder_in_place(torch, _native, n2, order, psi_nat, 0), (nf, n2, n2), 0, nf, ncols=m)
    else:
        _, psi_sorted = _native.apply_order(n2, nf, order, k_nat, psi_nat)
        sol._psi = _LazyModes(psi_sorted, (nf, n2, n2), 0, nf, ncols=m)


_GLOO = {}


def _host_group(dist):
    """CPU (gloo) process group for the small per-chunk object exchange of the sharded pipeline:
    it must not queue behind (or wait for) the compute streams like an NCCL collective would."""
    if 'g' not in _GLOO:
        try:
            group = dist.new_group(backend='gloo')
        except Exception:                                # no usable CPU transport on this host
            group = None
        # the ranks must agree: if gloo failed anywhere, everybody takes the default (NCCL) group
        import torch
        dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
        flag = torch.tensor([1 if group is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if group is not None:
                warnings.warn('axisafe_b200: no gloo group on some rank; the per-chunk exchange of the sharded '
                              'pipeline uses the NCCL group (it then queues behind the compute streams)',
                              RuntimeWarning)
            group = None
        _GLOO['g'] = group
    return _GLOO['g']


def _shard_plan(nf, world, unit, K=None, first=None):
    """Interleaved sharding of the frequency axis: the axis is cut into K global chunks and every
    chunk into ``world`` contiguous pieces, so that all ranks finish chunk c together and its tracked
    column order can be fixed (and its mode shapes downloaded) while chunk c+1 computes.
    Returns pieces[c][r] = (lo, hi).

    K = 4: chunks run pairwise on two streams; ``first`` = share of the first pair.  The download of
    the first pair has to fit behind the compute of the second one: x * r <= 1 - x with r = D2H time /
    compute time per point.  Measured: 254 KB per point at ~50 GB/s against 27 us of compute on one or
    two B200s (r = 0.19, x = 0.89: the 44/44/7/5 % split of the single-GPU pipeline), but only
    ~19 GB/s per GPU when eight ranks download through one host at the same time (r = 0.5, x = 0.67).
    The second pair is split 2:1 so that the last download is the small one.  Pieces are whole waves
    of ``unit`` matrices (one halo frequency included) when the sweep is long enough."""
    per = nf // world
    plan = os.environ.get('AXS_SHARD_PLAN', 'fine')
    if K is None and first is None and 'AXS_SHARD_K' not in os.environ and 'AXS_SHARD_FIRST' not in os.environ \
            and world >= 6 and per >= 8 * unit and plan != 'pairs2':
        # Many ranks share the host's memory system for their downloads (measured on 8 B200s:
        # ~19 GB/s per GPU, 13 us per frequency point against 27 us of compute), so the downloads have to
        # start early and keep running.  Measured end to end on 8 B200s, 2000 points per rank
        # (profiles/r2_pipe_trace_n8_*.txt):
        #   'pairs2'    two chunk pairs 67 % | 33 % (the round-1 plan)                    183 k points/s
        #   'fine'      seven chunks of two waves of `unit` matrices, remainder last       192 k  (default)
        #   'geometric' three balanced pairs 591 591 | 295 295 | 114 114                   185 k
        # Small chunks lose compute rate (a chunk of 295 matrices fills only half of the 4-CTAs-per-SM QR
        # stage), large ones delay the downloads; 'fine' is the best of the three at eight ranks.  At four
        # ranks (less contention per download) the two-pair plan measured 109.5 k against 106.7 k for 'fine',
        # so the fine plan starts at six ranks.
        if plan == 'balanced':
            # two balanced pairs 66 % | 34 %: the last download (34 % of the block) stays exposed, but every
            # chunk is large enough to compute at the full rate
            c1 = int(round(0.33 * per))
            c2 = per // 2 - c1
            sizes = [c1, c1, c2, per - 2 * c1 - c2]
        elif plan == 'fine':
            kept = 2 * unit - 1
            K = min(8, max(2, int(round(per / float(unit) / 2))))
            while kept * (K - 1) >= per:
                K -= 1
            sizes = [kept] * (K - 1) + [per - kept * (K - 1)]
        else:
            align = lambda x: max(1, int(round((x + 1.0) / unit))) * unit - 1
            c1, c2 = align(per * 4.0 / 7 / 2), align(per * 2.0 / 7 / 2)
            rem = per - 2 * c1 - 2 * c2
            if rem < 2:
                c1, c2 = per * 2 // 7, per // 7
                rem = per - 2 * c1 - 2 * c2
            sizes = [c1, c1, c2, c2, rem // 2, rem - rem // 2]
        gb = [0]
        for sz in sizes[:-1]:
            gb.append(gb[-1] + sz * world)
        gb.append(nf)
        pieces = []
        for c in range(len(sizes)):
            L = gb[c + 1] - gb[c]
            pieces.append([(gb[c] + (L * r) // world, gb[c] + (L * (r + 1)) // world) for r in range(world)])
        return pieces
    if K is None:
        K = int(os.environ.get('AXS_SHARD_K', 4 if per >= 512 else 2))     # A/B: more, equal chunks
    if first is None:
        first = float(os.environ.get('AXS_SHARD_FIRST', 0.888 if world <= 2 else (0.77 if world <= 4 else 0.67)))
    if K == 4:
        b1 = max(unit, int(first / 2 * per / unit + 0.05) * unit)      # rounded down: the download must fit
        b2 = max(unit, int(round((1 - first) * 2 / 3 * per / unit)) * unit)
        if per >= 8 * unit and (2 * b1 + b2) * world < nf:
            # every piece carries one halo frequency: b - 1 kept points + halo = whole waves
            gb = [0, (b1 - 1) * world, 2 * (b1 - 1) * world, (2 * b1 + b2 - 3) * world, nf]
        else:
            gb = [0, int(nf * first / 2), int(nf * first), int(nf * (first + (1 - first) * 2 / 3)), nf]
    else:
        gb = [nf * c // K for c in range(K + 1)]
    pieces = []
    for c in range(K):
        L = gb[c + 1] - gb[c]
        pieces.append([(gb[c] + (L * r) // world, gb[c] + (L * (r + 1)) // world) for r in range(world)])
    return pieces


def _solve_sharded_pipelined(sol, torch, _native, ops, nf, n2, rank, world, host_work=None):
    """Sharded sweep as a chunk pipeline (one process per GPU, torch.distributed).

    With one contiguous block per rank the absolute column order of a block is only known once
    every earlier rank has finished, i.e. at the very end, and the 0.5 GB psi_hat download of every
    rank is exposed.  Here the axis is interleaved (_shard_plan): all ranks work on pieces of the
    same global chunk, after chunk c they exchange the segment composites of their pieces
    (track_scan_segments / chain_segments, a few KB over a gloo group), fix the order of chunk c,
    gather the columns and download that part of psi_hat while chunk c+1 computes - the structure
    of _solve_pipelined with the host scan replaced by the cross-rank chain.  Every piece recomputes
    one halo frequency (the last point of the piece before it) so no eigenvector crosses a link.
    Results are bit-identical to the single-GPU sweep; the rank keeps the mode shapes of the
    frequencies ``sol.local_index``.
    """
    import torch.distributed as dist
    group = _host_group(dist)
    unit = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    pieces = _shard_plan(nf, world, unit)
    K = len(pieces)
    mine = [pieces[c][rank] for c in range(K)]
    halo = [1 if lo > 0 else 0 for lo, hi in mine]
    S = [h + hi - lo for h, (lo, hi) in zip(halo, mine)]            # slots per chunk (halo first)
    so = [0] + list(np.cumsum(S))                                   # slot offsets
    cnt = [hi - lo for lo, hi in mine]
    ko = [0] + list(np.cumsum(cnt))                                 # offsets of the kept rows
    slots, kept = int(so[-1]), int(ko[-1])
    nn = n2 * n2
    cs, xs, ss = _pipe_streams(torch)
    main = torch.cuda.current_stream()
    dev = 'cuda'
    w_loc = np.concatenate([sol.w[lo - h:hi] for h, (lo, hi) in zip(halo, mine)])
    omega = torch.from_numpy(w_loc).cuda()
    k_nat = torch.empty(slots * n2, dtype=torch.complex128, device=dev)
    psi_nat = torch.empty(slots * nn, dtype=torch.complex128, device=dev)
    info = torch.zeros(slots, dtype=torch.int32, device=dev)
    arg = torch.empty(slots * n2, dtype=torch.int32, device=dev)
    amax = torch.empty(slots * n2, dtype=torch.float64, device=dev)
    k_sorted = torch.empty(kept * n2, dtype=torch.complex128, device=dev)
    psi_sorted = torch.empty(kept * nn, dtype=torch.complex128, device=dev)
    order_d = torch.empty(kept * n2, dtype=torch.int32, device=dev)
    smax = max(S)
    works = [_native.EigWorkspace(ops.N, smax) for _ in range(2)]
    ybuf = [torch.empty(int(_native.lib().axs_track_worksize(ops.N, max(smax - 1, 1))), dtype=torch.uint8, device=dev)
            for _ in range(2)]
    pin = lambda *shape, dtype: torch.empty(*shape, dtype=dtype, pin_memory=True)
    arg_h = pin(slots, n2, dtype=torch.int32)
    amax_h = pin(slots, n2, dtype=torch.float64)
    info_h = pin(slots, dtype=torch.int32)
    order_h = pin(kept, n2, dtype=torch.int32)
    psi_h = pin(kept, n2, n2, dtype=torch.complex128)
    ev_track = [torch.cuda.Event() for _ in range(K)]
    # the host SLEEPS on the small-table events instead of spinning (cudaEventBlockingSync): with one
    # process per GPU the spinning threads of 8 ranks compete with the NCCL / gloo service threads for the
    # host cores (AXS_EVENT_BLOCKING=0 restores the spin wait for A/B)
    ev_small = [torch.cuda.Event(blocking=os.environ.get('AXS_EVENT_BLOCKING', '1') == '1') for _ in range(K)]
    SMAX = 6                                             # segments per piece carried by the fixed-size exchange
    xcap = 2 + n2 * (4 * SMAX - 3)
    xdev = 'cpu' if group is not None else dev
    xsend = torch.zeros(xcap, dtype=torch.int32, device=xdev)
    xrecv = torch.zeros(world * xcap, dtype=torch.int32, device=xdev)
    for s_ in cs + [xs, ss]:
        s_.wait_stream(main)

    def compute(c):
        a, b = int(so[c]), int(so[c + 1])
        st = cs[c % 2]
        with torch.cuda.stream(st):
            _native.eig_sweep_into(ops, omega[a:b], k_nat[a * n2:b * n2], psi_nat[a * nn:b * nn], info[a:b],
                                   works[c % 2])
            if b - a > 1:
                _native.track_pairs_into(ops, psi_nat[a * nn:b * nn], b - a - 1, arg[a * n2:(b - 1) * n2],
                                         amax[a * n2:(b - 1) * n2], ybuf[c % 2])
            ev_track[c].record(st)
        with torch.cuda.stream(ss):
            ss.wait_event(ev_track[c])
            # (kernel stores into the pinned tables: a cudaMemcpyAsync would queue behind the bulk psi_hat
            # download of the previous chunk on the D2H engine - measured on 8 B200s)
            _native.copy_to_pinned(info_h[a:b], info[a:b])
            if b - a > 1:
                _native.copy_to_pinned(arg_h[a:b - 1], arg[a * n2:(b - 1) * n2])
                _native.copy_to_pinned(amax_h[a:b - 1], amax[a * n2:(b - 1) * n2])
            ev_small[c].record(ss)

    # index tensors and buffers of the final gather, prepared before the pipeline starts (a pageable H2D
    # copy issued at the end would make the host wait for every queued kernel and download)
    keep_slots = torch.from_numpy(np.concatenate([np.arange(so[c] + halo[c], so[c + 1]) for c in range(K)])).to(dev)
    counts = [sum(pieces[c][r][1] - pieces[c][r][0] for c in range(K)) for r in range(world)]
    size = max(counts)
    index = np.concatenate([np.arange(pieces[c][r][0], pieces[c][r][1]) for r in range(world) for c in range(K)])
    src = np.concatenate([r * size + np.arange(counts[r]) for r in range(world)])
    back = np.empty(nf, dtype=np.int64)
    back[index] = src                                    # global frequency -> row of the gathered table
    back_d = torch.from_numpy(back).to(dev)
    k_h = pin(nf, n2, dtype=torch.complex128)
    o_h = pin(nf, n2 + 1, dtype=torch.int32)

    carry = [None]                                      # absolute order at the base of the next chunk

    def finish(c):
        a, b = int(so[c]), int(so[c + 1])
        ev_small[c].synchronize()
        arg_n, amax_n = arg_h[a:b - 1].numpy(), amax_h[a:b - 1].numpy()
        rel, starts = _native.track_scan_segments(arg_n, amax_n, n2)
        ends = starts[1:] + [b - a]
        comps = np.stack([rel[e - 1] for e in ends])
        fix = np.asarray(starts[1:], dtype=np.int64) - 1
        payload = (comps, arg_n[fix].copy(), amax_n[fix].copy(), int(np.count_nonzero(info_h[a:b].numpy())))
        gathered = _exchange_segments(torch, dist, group, payload, n2, SMAX, xsend, xrecv, world)
        bad = sum(g[3] for g in gathered)
        if bad:
            torch.cuda.synchronize()
            raise np.linalg.LinAlgError('eig algorithm (QR) did not converge at %d frequencies' % bad)
        sigma, carry[0] = _native.chain_segments([g[:3] for g in gathered], n2, sigma0=carry[0], return_final=True)
        rows = np.empty([b - a, n2], dtype=np.int32)
        for s, (p, e) in enumerate(zip(starts, ends)):
            np.take(rel[p:e], sigma[rank][s], axis=1, out=rows[p:e])
        q, r_ = int(ko[c]), int(ko[c + 1])
        order_h[q:r_].copy_(torch.from_numpy(rows[halo[c]:]))
        f0 = a + halo[c]                                 # first kept slot
        with torch.cuda.stream(xs):
            xs.wait_event(ev_track[c])
            order_d[q * n2:r_ * n2].copy_(order_h[q:r_].view(-1), non_blocking=True)
            _native.check(_native.lib().axs_apply_order(
                n2, r_ - q, _native._ptr(order_d[q * n2:r_ * n2]), _native._ptr(k_nat[f0 * n2:b * n2]),
                _native._ptr(psi_nat[f0 * nn:b * nn]), _native._ptr(k_sorted[q * n2:r_ * n2]),
                _native._ptr(psi_sorted[q * nn:r_ * nn]), _native._stream(torch)), 'axs_apply_order')
            psi_h[q:r_].copy_(psi_sorted[q * nn:r_ * nn].view(r_ - q, n2, n2), non_blocking=True)

    import time
    trace = os.environ.get('AXS_PIPE_TRACE')
    tl = [('start', time.perf_counter())]
    for c in range(K):
        compute(c)
        tl.append(('enq%d' % c, time.perf_counter()))
        if c == min(1, K - 1) and host_work is not None:
            host_work()                                  # sparse host by-products, hidden behind the kernels
            tl.append(('host', time.perf_counter()))
        if c >= 1:
            finish(c - 1)
            tl.append(('fin%d' % (c - 1), time.perf_counter()))
    finish(K - 1)
    tl.append(('fin%d' % (K - 1), time.perf_counter()))
    # full [nf, 2N] tables on every rank: all-gather of the kept rows (NCCL), rows back into global order
    main.wait_stream(xs)
    for s_ in cs:
        main.wait_stream(s_)
    k_nat_kept = k_nat.view(slots, n2).index_select(0, keep_slots)
    # (the per-frequency counts of unconverged eigenvalues travel as an extra column of the order table)
    order_info = torch.cat([order_d.view(kept, n2), info.index_select(0, keep_slots).view(kept, 1)], dim=1)
    tables = []
    for t in (k_sorted.view(kept, n2), order_info, k_nat_kept):
        pad = torch.zeros(size, t.shape[1], dtype=t.dtype, device=dev)
        pad[:kept] = t
        full = torch.empty(world * size, t.shape[1], dtype=t.dtype, device=dev)
        if t.dtype.is_complex:
            dist.all_gather_into_tensor(torch.view_as_real(full), torch.view_as_real(pad))
        else:
            dist.all_gather_into_tensor(full, pad)
        tables.append(full.index_select(0, back_d))
    k_h.copy_(tables[0], non_blocking=True)
    o_h.copy_(tables[1], non_blocking=True)
    tl.append(('gather', time.perf_counter()))
    torch.cuda.synchronize()
    tl.append(('sync', time.perf_counter()))
    if trace:                                            # host timeline of this rank (ms since entry)
        print('AXS_PIPE rank %d host ms: %s' % (rank, ' '.join('%s=%.1f' % (n_, (t_ - tl[0][1]) * 1e3) for n_, t_ in tl)),
              file=__import__('sys').stderr)
    sol.k = k_h.numpy()
    sol.order = np.ascontiguousarray(o_h.numpy()[:, :n2])
    sol.k_native = tables[2]                             # stays on the device until it is read
    sol.info = o_h.numpy()[:, n2].copy()                 # gathered per-frequency counts (all zero here: a
                                                         # non-converged chunk has raised LinAlgError above)
    sol.local_index = np.concatenate([np.arange(lo, hi) for lo, hi in mine])
    sol.local_range = None                               # not one contiguous block
    modes = _LazyModes(psi_sorted, (kept, n2, n2), None, nf)
    modes.host = psi_h.numpy()
    sol._psi = modes


def _exchange_segments(torch, dist, group, payload, n2, smax, send, recv, world):
    """All-gather of the per-piece tracking composites of one chunk (see _solve_sharded_pipelined).

    ``payload`` = (comps [S, n2] int32, d_arg [S-1, n2] int32, d_amax [S-1, n2] float64, bad count).
    One fixed-size int32 tensor per rank ([S, bad | comps | d_arg | d_amax bits], room for ``smax``
    segments) instead of pickled objects: a single collective without serialisation.  A piece with
    more than ``smax`` fix-up segments is rare (a segment ends where a tracked mode falls below the 0.9
    threshold); every rank sees every header, so all of them then repeat the exchange with
    all_gather_object.  Returns the list of payload tuples of all ranks.
    """
    comps, d_arg, d_amax, bad = payload
    S = comps.shape[0]
    buf = np.zeros(send.numel(), dtype=np.int32)
    buf[0], buf[1] = S, bad
    if S <= smax:
        o = 2
        buf[o:o + S * n2] = comps.reshape(-1); o += smax * n2
        buf[o:o + (S - 1) * n2] = d_arg.reshape(-1); o += (smax - 1) * n2
        buf[o:o + 2 * (S - 1) * n2] = np.ascontiguousarray(d_amax, dtype=np.float64).reshape(-1).view(np.int32)
    send.copy_(torch.from_numpy(buf))
    dist.all_gather_into_tensor(recv, send, group=group)
    got = recv.cpu().numpy().reshape(world, -1)
    if (got[:, 0] > smax).any():
        gathered = [None] * world
        dist.all_gather_object(gathered, payload, group=group)
        return gathered
    out = []
    for r in range(world):
        Sr, o = int(got[r, 0]), 2
        c = got[r, o:o + Sr * n2].reshape(Sr, n2).copy(); o += smax * n2
        a = got[r, o:o + (Sr - 1) * n2].reshape(Sr - 1, n2).copy(); o += (smax - 1) * n2
        m = got[r, o:o + 2 * (Sr - 1) * n2].copy().view(np.float64).reshape(Sr - 1, n2)
        out.append((c, a, m, int(got[r, 1])))
    return out


def _sweep_large(torch, _native, ops, omega, halo, select=None):
    """Chunked sweep for 2N above the shared-memory kernels (eig_big.cu).

    The per-frequency workspace is ~32 (2N)^2 bytes, so the frequency block is processed in
    chunks sized from the free HBM; consecutive chunks overlap by one frequency whose mode
    shapes are carried over (not recomputed) for the tracking pair at the seam.  Returns the
    same tables as the one-shot path; ``psi`` is None when the whole block of mode shapes does
    not fit next to the workspace.
    """
    n2, nloc = 2 * ops.N, len(omega)
    L = _native.lib()
    free, _ = torch.cuda.mem_get_info()
    # blocks that torch's caching allocator holds but does not use (the workspace of a previous sweep) are
    # available to this one as well
    free += max(0, torch.cuda.memory_reserved() - torch.cuda.memory_allocated())
    per_mat = n2 * n2 * 16
    keep = nloc * per_mat <= 0.45 * free
    budget = 0.8 * (free - (nloc * per_mat if keep else 0))
    fixed = int(L.axs_eig_worksize(ops.N, 1))
    slope = int(L.axs_eig_worksize(ops.N, 2)) - fixed
    fixed -= slope
    # per chunk of C frequencies: workspace + (C + 1) mode-shape slots + (C + 1) slots of Y = B psi
    C = int((budget - fixed - 2 * per_mat) // (slope + 2 * per_mat))
    if C < 1:
        raise MemoryError('2N = %d: one frequency needs %.1f GB of HBM workspace' % (n2, (fixed + slope) / 1e9))
    C = min(C, nloc)
    C = -(-nloc // -(-nloc // C))                   # equal chunks (148 points as 74 + 74, not 112 + 36)
    k = torch.empty(nloc * n2, dtype=torch.complex128, device='cuda')
    info = torch.zeros(nloc, dtype=torch.int32, device='cuda')
    arg = torch.empty(max(nloc - 1, 0) * n2, dtype=torch.int32, device='cuda')
    amax = torch.empty(max(nloc - 1, 0) * n2, dtype=torch.float64, device='cuda')
    psi_all = torch.empty(nloc * n2 * n2, dtype=torch.complex128, device='cuda') if keep else None
    work = _native.EigWorkspace(ops.N, C)
    slots = torch.empty((C + 1) * n2 * n2, dtype=torch.complex128, device='cuda')
    nn = n2 * n2
    for c0 in range(0, nloc, C):
        c1 = min(c0 + C, nloc)
        cur = slots[nn:(c1 - c0 + 1) * nn]
        _native.eig_sweep_into(ops, omega[c0:c1], k[c0 * n2:c1 * n2], cur, info[c0:c1], work)
        if select is not None:                        # subset mode: unselected mode columns zeroed before tracking
            sigma, m_sel, rank = select
            _native.select_modes(n2, c1 - c0, k[c0 * n2:c1 * n2], sigma[c0:c1], m_sel, cur, rank[c0 * n2:c1 * n2])
        first = 0 if c0 > 0 else 1                    # slot 0 holds the last frequency of the previous chunk
        npairs = (c1 - c0) - first
        if npairs > 0:
            a, m = _native.track_pairs(ops, slots[first * nn:(c1 - c0 + 1) * nn], npairs)
            p0 = c0 - 1 + first
            arg[p0 * n2:(p0 + npairs) * n2] = a
            amax[p0 * n2:(p0 + npairs) * n2] = m
        if keep:
            psi_all[c0 * nn:c1 * nn] = cur
        if c1 < nloc:
            slots[:nn] = slots[(c1 - c0) * nn:(c1 - c0 + 1) * nn].clone()
    return k, psi_all, info, arg, amax


def _reorder_in_place(torch, _native, n2, order_loc, psi_nat, halo):
    """Tracked column order applied frequency block by block (no second full-size copy)."""
    nn = n2 * n2
    psi = psi_nat[halo * nn:]
    nf = order_loc.shape[0]
    step = max(1, min(nf, (1 << 30) // (nn * 16)))
    dummy = torch.empty(step * n2, dtype=torch.complex128, device='cuda')
    for c0 in range(0, nf, step):
        c1 = min(c0 + step, nf)
        blk = psi[c0 * nn:c1 * nn]
        _, out = _native.apply_order(n2, c1 - c0, np.ascontiguousarray(order_loc[c0:c1]),
                                     dummy[:(c1 - c0) * n2], blk)
        blk.copy_(out)
    return psi


def _raise_unconverged(info):
    if info.any():
        bad = int(np.count_nonzero(info))
        raise np.linalg.LinAlgError('eig algorithm (QR) did not converge at %d frequencies' % bad)


_PINNED = {}


def _pinned(torch, name, shape, dtype):
    """Pinned host staging buffer, kept between sweeps of the same shape."""
    key = (name, tuple(shape), dtype)
    if key not in _PINNED:
        for old in [k for k in _PINNED if k[0] == name]:
            del _PINNED[old]
        _PINNED[key] = torch.empty(*shape, dtype=dtype, pin_memory=torch.cuda.is_available())
    return _PINNED[key]


def _track_sharded(sol, torch, _native, bounds, rank, n2, k_loc, arg_loc, amax_loc, info_loc):
    """Tracking tail of a sharded sweep: sets ``k, order, info, k_native`` on ``sol`` (full tables,
    identical on every rank) and returns this rank's rows of the column order (host, device).

    The reference's scan (solver.py:159-182) is sequential over the frequencies.  Here every rank
    scans its own block relative to the block's base frequency (_native.track_scan_segments), the
    ranks exchange one small object each - the composed map of every segment plus the pair tables
    of the fix-up rows - every rank replays the chain over those (O(segments), not O(nf)), applies
    its absolute start order to its own rows, reorders its own k / psi on the device, and the
    tracked k and the order are all-gathered over NCCL.  Host work and PCIe traffic per rank are
    O(nf / world) apart from the final [nf, 2N] tables the API exposes on every rank.
    """
    import torch.distributed as dist
    world = len(bounds) - 1
    nf = bounds[-1]
    lo, hi = bounds[rank], bounds[rank + 1]
    halo = 1 if lo > 0 else 0
    P = arg_loc.shape[0]                                    # pairs of this block (halo pair included)
    on_gpu = k_loc.is_cuda
    if on_gpu:
        arg_h = _pinned(torch, 'arg', (max(P, 1), n2), torch.int32)[:P]
        amax_h = _pinned(torch, 'amax', (max(P, 1), n2), torch.float64)[:P]
        info_h = _pinned(torch, 'info', (hi - lo,), torch.int32)
        arg_h.copy_(arg_loc, non_blocking=True)
        amax_h.copy_(amax_loc, non_blocking=True)
        info_h.copy_(info_loc, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    else:
        arg_h, amax_h, info_h = arg_loc, amax_loc, info_loc
    arg_n, amax_n = arg_h.numpy(), amax_h.numpy()
    rel, starts = _native.track_scan_segments(arg_n, amax_n, n2)
    ends = starts[1:] + [P + 1]
    comps = np.stack([rel[e - 1] for e in ends])
    fix = np.asarray(starts[1:], dtype=np.int64) - 1        # pair index of every fix-up row
    payload = (comps, arg_n[fix].copy(), amax_n[fix].copy(), info_h.numpy().copy())
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    sol.info = np.concatenate([g[3] for g in gathered])
    _raise_unconverged(sol.info)
    sigma = _native.chain_segments([g[:3] for g in gathered], n2)[rank]
    order_rows = np.empty([P + 1, n2], dtype=np.int32)
    for s, (a, e) in enumerate(zip(starts, ends)):
        np.take(rel[a:e], sigma[s], axis=1, out=order_rows[a:e])
    order_loc = np.ascontiguousarray(order_rows[halo:])     # frequencies lo .. hi-1
    # tracked k of this block on the device, then the full tables over NCCL
    dev = k_loc.device
    order_dev = torch.from_numpy(order_loc).to(dev, non_blocking=True)
    if on_gpu:
        k_sorted, _ = _native.apply_order(n2, hi - lo, order_dev.reshape(-1), k_loc.reshape(-1), None)
        k_sorted = k_sorted.view(hi - lo, n2)
    else:
        k_sorted = torch.gather(k_loc.reshape(hi - lo, n2), 1, order_dev.long())
    k_all, order_all, k_nat_all = _gather_rows(torch, bounds, rank, n2, [k_sorted, order_dev, k_loc.reshape(hi - lo, n2)])
    if on_gpu:
        k_h = _pinned(torch, 'k', (nf, n2), torch.complex128)
        o_h = _pinned(torch, 'order', (nf, n2), torch.int32)
        k_h.copy_(k_all, non_blocking=True)
        o_h.copy_(order_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        sol.k, sol.order = k_h.numpy().copy(), o_h.numpy().copy()
    else:
        sol.k, sol.order = k_all.numpy(), order_all.numpy()
    sol.k_native = k_nat_all                                # stays on the device until it is read
    return order_loc, order_dev.reshape(-1)


def _gather_rows(torch, bounds, rank, n2, tensors):
    """All-gather [rows_r, n2] blocks of the ranks (padded to equal size) into [nf, n2] tables."""
    import torch.distributed as dist
    world = len(bounds) - 1
    size = max(bounds[r + 1] - bounds[r] for r in range(world))
    rows = bounds[rank + 1] - bounds[rank]
    out = []
    for t in tensors:
        pad = torch.zeros(size, n2, dtype=t.dtype, device=t.device)
        pad[:rows] = t
        full = torch.empty(world * size, n2, dtype=t.dtype, device=t.device)
        if t.dtype.is_complex:
            dist.all_gather_into_tensor(torch.view_as_real(full), torch.view_as_real(pad))
        else:
            dist.all_gather_into_tensor(full, pad)
        if all(bounds[r + 1] - bounds[r] == size for r in range(world)):
            out.append(full)
        else:
            out.append(torch.cat([full[r * size:r * size + bounds[r + 1] - bounds[r]] for r in range(world)]))
    return out
