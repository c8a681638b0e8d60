"""TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy/SciPy) of the reference hot path.

This module is the *checker* for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product package (``axisafe_b200``) never does and
fails loudly when its CUDA library is missing.

Parity status: **pinned** - ``tests/test_oracle_golden.py`` checks every function
below against fixtures produced by importing the unmodified reference
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and, inside the build
container, against the live reference (``oracle/refshim.py``).  The reference
itself ships no tests or golden vectors.

What is restated (reference file:line):
  * quadrature tables            shape_functions.py:101-169, :171-315
  * six element classes          elements.py:148-229, :293-330, :447-586, :683-737,
                                 :810-894, :962-1007  (index form)
  * mesh generation              mesh.py:153-288
  * DOF numbering / LM           mesh.py:337-456
  * coupling blocks + scatter    mesh.py:493-619
  * T-transform, B, A(w), S      solver.py:59-74, :115-134
  * eig + B-normalisation        solver.py:136-149   (scipy.linalg.eig == LAPACK zgeev,
                                 the same third-party routine the reference calls)
  * biorthogonal mode tracking   solver.py:151-182
  * subset (shift-invert) mode   solver.py:137-140   (scipy.sparse.linalg.eigs == ARPACK znaupd)
The only third-party arithmetic is ``scipy.linalg.eig`` / ``scipy.sparse.linalg.eigs``
(SciPy 1.18.1 / OpenBLAS 0.3.30 / ARPACK in this image; the reference pins no version).
"""
import numpy as np
import scipy.linalg as la

TWO_PI = 2 * np.pi

# --------------------------------------------------------------------------- tables


def gll(m):
    """GLL nodes/weights (shape_functions.py:101-140) via eigen-free Newton."""
    p = m - 1
    x = np.cos(np.pi * np.arange(m) / p)
    P = np.zeros([m, m])
    for _ in range(100):
        P[:, 0], P[:, 1] = 1.0, x
        for k in range(2, m):
            P[:, k] = ((2 * k - 1) * x * P[:, k - 1] - (k - 1) * P[:, k - 2]) / k
        dx = (x * P[:, p] - P[:, p - 1]) / (m * P[:, p])
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    P[:, 0], P[:, 1] = 1.0, x
    for k in range(2, m):
        P[:, k] = ((2 * k - 1) * x * P[:, k - 1] - (k - 1) * P[:, k - 2]) / k
    order = np.argsort(x)
    return x[order], (2.0 / (p * m * P[:, p] ** 2))[order]


def _jac(x, n, a, b):
    x = np.atleast_1d(np.asarray(x, float))
    p0, p1 = np.ones_like(x), 0.5 * (a - b + (a + b + 2) * x)
    if n == 0:
        return p0
    for k in range(1, n):
        a1 = 2 * (k + 1) * (k + a + b + 1) * (2 * k + a + b)
        a2 = (2 * k + a + b + 1) * (a * a - b * b)
        a3 = (2 * k + a + b) * (2 * k + a + b + 1) * (2 * k + a + b + 2)
        a4 = 2 * (k + a) * (k + b) * (2 * k + a + b + 2)
        p0, p1 = p1, ((a2 + a3 * x) * p1 - a4 * p0) / a1
    return p1


def glj(Q):
    """GLJ(0,1) nodes/weights (shape_functions.py:142-169)."""
    from scipy.special import gamma
    from math import factorial
    n = Q - 2
    # interior nodes: zeros of P^{(1,2)}_{Q-2}; polish scipy's roots with Newton
    from scipy.special import roots_jacobi
    xi = np.sort(roots_jacobi(n, 1, 2)[0])
    for _ in range(3):
        xi = xi - _jac(xi, n, 1, 2) / (0.5 * (n + 4) * _jac(xi, n - 1, 2, 3))
    ksi = np.r_[-1.0, xi, 1.0]
    C = 2.0 ** 2 * gamma(Q) * gamma(1 + Q) / ((Q - 1) * factorial(Q - 1) * gamma(Q + 2))
    w = C / _jac(ksi, Q - 1, 0, 1) ** 2
    w[0] *= 2.0
    return ksi, w


def dmatrix(ksi):
    """D[a, q] = l_a'(ksi_q)  (shape_functions.py:195-204 / :239-263)."""
    m = len(ksi)
    D = np.zeros([m, m])
    for a in range(m):
        others = np.delete(np.arange(m), a)
        denom = np.prod(ksi[a] - ksi[others])
        for q in range(m):
            if q == a:
                D[a, q] = np.sum(1.0 / (ksi[a] - ksi[others]))
            else:
                rest = others[others != q]
                D[a, q] = np.prod(ksi[q] - ksi[rest]) / denom
    return D


# ------------------------------------------------------------------------- elements

_L = np.zeros([6, 3]); _L[1, 0], _L[5, 1] = 1.0, -1.0
_LR = np.zeros([6, 3]); _LR[0, 0] = _LR[4, 2] = _LR[5, 1] = 1.0
_LT = np.zeros([6, 3]); _LT[1, 1] = _LT[3, 2] = _LT[5, 0] = 1.0
_LZ = np.zeros([6, 3]); _LZ[2, 2] = _LZ[3, 1] = _LZ[4, 0] = 1.0


def pml_exp(r, start, thk, gam):
    """Exponential PML profile and stretched radius (elements.py:37-51, :833-837)."""
    s = (r - start) / thk
    a, b = gam.real, gam.imag
    g = np.exp(a * s) - 1j * (np.exp(b * s) - 1)
    rt = start + thk / a * (np.exp(a * s) - 1) - 1j * thk / b * (np.exp(b * s) - 1) + 1j * s * thk
    return g, rt


def iso_C(lam, mu, eta):
    C = np.zeros([6, 6], complex)
    C[:3, :3] = lam
    C[np.arange(3), np.arange(3)] = lam + 2 * mu
    C[np.arange(3, 6), np.arange(3, 6)] = mu
    return (1 + 1j * eta) * C


def element_matrices(etype, nodes, material, n, pml=None):
    """All element matrices of one element, already reduced by the axis BCs.

    etype in {'SLAX6','ALAX6','SLAX6_core','ALAX6_core','SLAX6_PML','ALAX6_PML'}.
    Returns dict with K_1..K_6, K_F, K_Ft, K_Fz, M and the interface vectors.
    """
    nodes = np.asarray(nodes, float)
    m = len(nodes)
    core = etype.endswith('_core')
    is_pml = etype.endswith('_PML')
    solid = etype.startswith('S')
    ksi, w = glj(m) if core else gll(m)
    D = dmatrix(ksi)
    J = D.T @ nodes
    Nr = D / J
    r = nodes.copy()
    axis = np.round(r, 7) == 0 if core else np.zeros(m, bool)
    if is_pml:
        g, rt = pml_exp(r, pml[0], pml[1], pml[2])
        invr, ginv, wt = 1.0 / rt, 1.0 / g, TWO_PI * w * J * rt * g
    else:
        g = np.ones(m)
        ginv = np.ones(m)
        with np.errstate(divide='ignore'):
            invr = np.where(axis, 0.0, 1.0 / np.where(axis, 1.0, r))
        if core:
            rat = np.where(axis, J, r / np.where(axis, 1.0, 1 + ksi))
            wt = TWO_PI * w * J * rat
        else:
            wt = TWO_PI * w * J * r
    out = {}
    eye = np.eye(m)
    if solid:
        C = iso_C(material[0], material[1], material[3])
        rho = material[2]
        # shape-function matrices at every point: Nm[q] = kron(e_q^T, I3), Nrm[q] = kron(Nr[:,q]^T, I3)
        Nm = np.einsum('aq,cd->qcad', eye, np.eye(3)).reshape(m, 3, 3 * m)
        Nrm = np.einsum('aq,cd->qcad', Nr, np.eye(3)).reshape(m, 3, 3 * m)
        B1 = np.einsum('sc,qcd->qsd', _L, Nm) * invr[:, None, None] + \
            np.einsum('sc,qcd->qsd', _LR, Nrm) * ginv[:, None, None]
        B2 = np.einsum('sc,qcd->qsd', _LT, Nm) * invr[:, None, None]
        B3 = np.einsum('sc,qcd->qsd', _LZ, Nm) + 0j
        for q in np.where(axis)[0]:          # l'Hopital forms on the axis
            B1[q] = (_L + _LR) @ Nrm[q]
            B2[q] = _LT @ Nrm[q]

        def form(X, Y):
            return np.einsum('q,qsa,st,qtb->ab', wt, X, C, Y)
        KF1, KF2, KF3 = form(B1, B3), form(B1, B2), form(B2, B3)
        out['K_1'], out['K_5'], out['K_6'] = form(B1, B1), form(B2, B2), form(B3, B3)
        out['K_2'], out['K_3'], out['K_4'] = KF2.T - KF2, KF1.T - KF1, KF3.T + KF3
        out['K_F'], out['K_Ft'], out['K_Fz'] = KF1.T, KF3.T, out['K_6']
        if is_pml:
            M = np.einsum('q,qca,qcb->ab', wt * rho, Nm, Nm)
        else:
            # bug-compatible complex64 accumulation (elements.py:99, :217, :397, :529)
            M = np.zeros([3 * m, 3 * m], np.complex64)
            for q in range(m):
                scal = TWO_PI * w[q]
                lastf = r[q] if not core else (J[q] if axis[q] else r[q] / (1 + ksi[q]))
                term = scal * (rho * (Nm[q].T @ Nm[q])) * J[q] * lastf
                M += term
        out['M'] = M
        e_r = np.zeros(3 * m); e_r[0] = TWO_PI
        e_l = np.zeros(3 * m); e_l[3 * (m - 1)] = TWO_PI * nodes[-1]
        if not core:
            out['Hs0'] = e_r.reshape(-1, 1)
        if not is_pml:
            out['Hs1'] = e_l.reshape(-1, 1)
        drop = []
        if core and 0 in nodes:
            drop = [0, 1] if n == 0 else ([2] if n == 1 else [0, 1, 2])
        if drop:
            for key in ('K_1', 'K_2', 'K_3', 'K_4', 'K_5', 'K_6', 'M', 'K_F', 'K_Ft', 'K_Fz'):
                out[key] = np.delete(np.delete(out[key], drop, 0), drop, 1)
            out['Hs1'] = np.delete(out['Hs1'], drop, 0)
    else:
        rho = material[1]
        cp2 = material[0] / material[1]
        sc = -rho * wt
        K5w = sc * invr ** 2
        K1 = np.einsum('q,aq,bq->ab', sc * ginv ** 2, Nr, Nr)
        K5 = np.diag(K5w) + 0j
        for q in np.where(axis)[0]:
            K5[q, q] = 0.0
            K5 = K5 + sc[q] * np.outer(Nr[:, q], Nr[:, q])
        K6 = np.diag(sc) + 0j
        M = np.diag(sc / ((cp2 ** 0.5) ** 2))
        if not is_pml:
            M = M.real.copy()
        z = np.zeros([m, m], complex)
        out.update({'K_1': K1 + 0j, 'K_2': z, 'K_3': z, 'K_4': z, 'K_5': K5, 'K_6': K6,
                    'K_F': z, 'K_Ft': K5, 'K_Fz': K6, 'M': M})
        hf0 = np.zeros([m, 1]); hf0[0] = rho
        hf1 = np.zeros([m, 1]); hf1[-1] = rho * nodes[-1]
        if not core:
            out['Hf0'] = hf0
        if not is_pml:
            out['Hf1'] = hf1
        if core and 0 in nodes and n != 0:
            for key in ('K_1', 'K_5', 'K_6', 'M', 'K_F', 'K_Ft', 'K_Fz', 'K_2', 'K_3', 'K_4'):
                out[key] = np.delete(np.delete(out[key], [0], 0), [0], 1)
    return out


# ----------------------------------------------------------------------- mesh + DOFs


def build_mesh(sets, max_frequency, PML=0, PML_props=0):
    """1-D layered mesh (mesh.py:153-288, without lubricated coupling).

    Returns dict(nodes_r [nn], conn [list of node-index arrays, 0-based],
    etype [list], material [list], layer_of [list], pml_elems [list]).
    """
    r0 = sets[0][1]
    radii, conn, etype, mats, layer_of = [r0 * 1.0], [], [], [], []
    first = True
    for layer in sets:
        name, _, thk, mat, typ = layer[:5]
        cmin = (mat[1] / mat[2]) ** 0.5 if len(mat) == 4 else (mat[0] / mat[1]) ** 0.5
        lam_min = cmin / max_frequency
        order = layer[5] if len(layer) > 5 and layer[5] != 0 else int(np.ceil(np.pi * thk / lam_min) + 3)
        nel = layer[6] if len(layer) > 6 and layer[6] != 0 else 1
        h = thk / nel
        for _ in range(nel):
            start = radii[-1] if not first else r0
            ksi = glj(order + 1)[0] if start == 0.0 else gll(order + 1)[0]
            loc = (ksi + 1) * 0.5
            pos = start + loc * h
            base = len(radii) - 1
            if first:
                radii[0] = pos[0]
                first = False
            radii.extend(pos[1:])
            conn.append(np.arange(base, base + order + 1))
            etype.append(typ)
            mats.append(mat)
            layer_of.append(name)
    pml_elems = [e for e, l in enumerate(layer_of) if PML != 0 and l == PML]
    return dict(nodes_r=np.array(radii), conn=conn, etype=etype, material=mats,
                layer_of=layer_of, pml_elems=pml_elems, PML_props=PML_props)


def number_dofs(mesh, n):
    """DOF numbering, T diagonal and per-element location vectors (mesh.py:337-456)."""
    conn, et = mesh['conn'], list(mesh['etype'])
    r = mesh['nodes_r']
    nel = len(conn)
    for e in range(nel):
        if et[e] in ('SLAX6', 'ALAX6') and 0 in r[conn[e]]:
            et[e] += '_core'
    per_node = []
    tcomp = []
    for e in range(nel):
        m = len(conn[e])
        solid = et[e][0] == 'S'
        d = [3 if solid else 1] * m
        t = [[1, 1j, 1j] if solid else [1]] * m
        if et[e].endswith('_core') and 0 in r[conn[e]]:
            if solid:
                d[0], t[0] = {0: (1, [1j]), 1: (2, [1, 1j])}.get(n, (0, []))
            elif n != 0:
                d[0], t[0] = 0, []
        per_node.append(d)
        tcomp.append(t)
    owner = {}
    for e in range(nel):
        for nd in conn[e]:
            owner.setdefault(int(nd), []).append(e)
    ID, Tdiag, nxt = {}, [], 0
    solid_dofs, fluid_dofs = [], []
    for e in range(nel):
        for i, nd in enumerate(conn[e]):
            nd = int(nd)
            es = owner[nd]
            iface = len(es) == 2 and mesh['etype'][es[0]][0] != mesh['etype'][es[1]][0]
            if nd in ID and not iface:
                continue
            new = list(range(nxt, nxt + per_node[e][i]))
            if iface and es.index(e) == 1:
                ID[nd] = ID[nd] + new
            else:
                ID[nd] = new
            Tdiag.extend(tcomp[e][i])
            (solid_dofs if et[e][0] == 'S' else fluid_dofs).extend(new)
            nxt += per_node[e][i]
    LM = []
    for e in range(nel):
        seq = []
        for nd in conn[e]:
            nd = int(nd)
            es = owner[nd]
            iface = len(es) == 2 and mesh['etype'][es[0]][0] != mesh['etype'][es[1]][0]
            if not iface:
                seq.extend(ID[nd])
            elif e == es[0]:
                seq.extend(ID[nd][:per_node[e][-1]])
            else:
                seq.extend(ID[nd][per_node[e - 1][-1]:])
        LM.append(seq)
    pairs = sorted({tuple(sorted(es)) for es in owner.values()
                    if len(es) == 2 and mesh['etype'][es[0]][0] != mesh['etype'][es[1]][0]})
    return dict(N=nxt, etype=et, ID=ID, LM=LM, Tdiag=np.array(Tdiag, complex),
                solid=solid_dofs, fluid=fluid_dofs, pairs=[list(p) for p in pairs])


def assemble(mesh, n):
    """Global dense matrices K1..K6, M, KF, KFt, KFz, K1c (mesh.py:467-619)."""
    num = number_dofs(mesh, n)
    N = num['N']
    em = [element_matrices(num['etype'][e], mesh['nodes_r'][mesh['conn'][e]], mesh['material'][e], n,
                           mesh['PML_props']) for e in range(len(mesh['conn']))]
    G = {k: np.zeros([N, N], complex) for k in
         ('K1', 'K2', 'K3', 'K4', 'K5', 'K6', 'M', 'KF', 'KFt', 'KFz')}
    names = dict(K1='K_1', K2='K_2', K3='K_3', K4='K_4', K5='K_5', K6='K_6', M='M',
                 KF='K_F', KFt='K_Ft', KFz='K_Fz')
    for e, mats in enumerate(em):
        lm = np.array(num['LM'][e], int)
        for gk, ek in names.items():
            G[gk][np.ix_(lm, lm)] += mats[ek]
    K1c = np.zeros([N, N])
    for p in num['pairs']:
        a, b = em[p[0]], em[p[1]]
        if num['etype'][p[0]][0] == 'A':
            H = -b['Hs0'] @ a['Hf1'].T
            blk = np.block([[np.zeros([H.shape[1]] * 2), H.T], [H, np.zeros([H.shape[0]] * 2)]])
        else:
            H = a['Hs1'] @ b['Hf0'].T
            blk = np.block([[np.zeros([H.shape[0]] * 2), H], [H.T, np.zeros([H.shape[1]] * 2)]])
        lm = num['LM'][p[0]] + num['LM'][p[1]]
        ii, jj = np.nonzero(blk)
        for i, j in zip(ii, jj):
            K1c[lm[i], lm[j]] += blk[i, j]
    G['K1c'] = K1c
    G['coupled'] = len(num['pairs']) > 0
    G['num'] = num
    G['elements'] = em
    return G


# ---------------------------------------------------------------------------- solver


def operators(G, n):
    """T-transformed operators of the quadratic pencil (solver.py:59-74, :115-133).

    Returns K0s (static part of K0), C, K6, M, K1c, B with
    K0(w) = K0s - w^2 M + i w K1c and P(k) = K0 + k C + k^2 K6.
    """
    T = G['num']['Tdiag']
    hat = lambda K: (T[:, None] * K * T.conj()[None, :]) / (-1j)
    K0s = G['K1'] + n ** 2 * G['K5'] + n * hat(G['K2'])
    C = n * G['K4'] + hat(G['K3'])
    N = len(T)
    B = np.block([[np.zeros([N, N]), G['K6']], [G['K6'], C]])
    return dict(K0s=K0s, C=C, K6=G['K6'], M=G['M'], K1c=G['K1c'], B=B, N=N)


def S_matrix(op, w):
    """S(w) = B^-1 A(w)  (solver.py:115-134) through a dense solve."""
    N = op['N']
    K0 = op['K0s'] - w ** 2 * op['M'] + (1j * w * op['K1c'] if op['K1c'].any() else 0)
    A = np.block([[op['K6'], np.zeros([N, N])], [np.zeros([N, N]), -K0]])
    return np.linalg.solve(op['B'], A)


def eig_normalised(op, w):
    """zgeev eigenpairs of S(w), B-normalised (solver.py:136-142)."""
    val, vec = la.eig(S_matrix(op, w))
    nf = 1 / np.diag(vec.T @ op['B'] @ vec) ** 0.5
    return val, vec * nf


def track(prev_psi, B, val, psi):
    """One step of biorthogonal mode tracking (solver.py:151-182)."""
    bi = prev_psi.T @ B @ psi
    order = np.argmax(abs(bi), axis=1)
    ok = np.round(abs(bi[np.arange(len(bi)), order]), 2) > 0.9
    if not ok.all():
        free = list(set(range(len(bi))) - set(order[ok]))
        j = 0
        for i in range(len(ok)):
            if not ok[i]:
                order[i] = free[j]
                j += 1
    return val[order], psi[:, order], order


def solve(G, n, f):
    """Full tracked sweep: returns k [nf, 2N], psi_hat [nf, 2N, 2N] (solver.py:76-184)."""
    op = operators(G, n)
    w = 2 * np.pi * np.asarray(f, float)
    k = np.zeros([len(w), 2 * op['N']], complex)
    psi = np.zeros([len(w), 2 * op['N'], 2 * op['N']], complex)
    for i, wi in enumerate(w):
        val, vec = eig_normalised(op, wi)
        if i == 0:
            k[i], psi[i] = val, vec
        else:
            k[i], psi[i], _ = track(psi[i - 1], op['B'], val, vec)
    return k, psi


def solve_subset(G, n, f, no_of_waves, central_wavespeed):
    """The reference's shift-invert mode (solver.py:137-140 inside the loop of :119-182):
    ``scipy.sparse.linalg.eigs(S, k=m, sigma=w/c, which='LM')`` (ARPACK, the same third-party
    routine, started from a fixed vector here), B-normalised and tracked among the m columns.
    Returns k [nf, m], psi_hat [nf, 2N, m]."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    op = operators(G, n)
    w = 2 * np.pi * np.asarray(f, float)
    m = int(no_of_waves)
    k = np.zeros([len(w), m], complex)
    psi = np.zeros([len(w), 2 * op['N'], m], complex)
    v0 = np.ones(2 * op['N'], complex)
    for i, wi in enumerate(w):
        val, vec = spla.eigs(sps.csc_matrix(S_matrix(op, wi)), k=m, sigma=wi / central_wavespeed, which='LM', v0=v0)
        vec = vec / np.diag(vec.T @ op['B'] @ vec) ** 0.5
        if i == 0:
            k[i], psi[i] = val, vec
        else:
            k[i], psi[i], _ = track(psi[i - 1], op['B'], val, vec)
    return k, psi


# ------------------------------------------------------------- energy velocity (row N3)


def energy_velocity(G, mesh, n, w, k, psi):
    """e_k, e_p, p, en_vel for all modes (solver.py:301-409, ``only_propagating=False``).

    ``mesh`` is the build_mesh dict (its ``pml_elems`` are excluded, solver.py:318-341),
    ``k`` [nf, 2N] and ``psi`` [nf, 2N, 2N] a tracked sweep.  Dense NumPy, small cases only.
    """
    num, em = G['num'], G['elements']
    N = num['N']
    names = dict(K1='K_1', K2='K_2', K5='K_5', K6='K_6', M='M', KF='K_F', KFt='K_Ft')
    core = [e for e in range(len(em)) if e not in mesh['pml_elems']]
    sub = {key: np.zeros([N, N], complex) for key in names}
    for e in core:                                            # Mesh.assemble_subset, mesh.py:621-755
        lm = np.array(num['LM'][e], int)
        for gk, ek in names.items():
            sub[gk][np.ix_(lm, lm)] += em[e][ek]
    mult = np.ones(N)
    mult[num['fluid']] = -1                                   # physical pressure sign, solver.py:353-362
    for key in sub:
        sub[key] = sub[key] * mult[None, :]
    core_dofs = sorted({d for e in core for d in num['LM'][e]})
    q = (num['Tdiag'].conj()[None, :, None] * psi[:, N:, :])[:, core_dofs, :]     # [nf, dofs, modes]
    R = {key: m[np.ix_(core_dofs, core_dofs)] for key, m in sub.items()}
    form = lambda X: np.einsum('fam,ab,fbm->fm', q.conj(), X, q)
    w = np.asarray(w, float).reshape(-1, 1)
    e_k = w ** 2 / 4 * form(R['M'])
    e_p = 0.25 * (form(R['K1']) + 1j * n * form(R['K2']) + n ** 2 * form(R['K5']) + 1j * k.conj() * form(R['KF'])
                  - 1j * k * form(R['KF'].T) + k * n * form(R['KFt'].T) + k.conj() * n * form(R['KFt'])
                  + k * k.conj() * form(R['K6']))
    p = -w / 2 * np.imag(form(R['KF']) - 1j * n * form(R['KFt']) - 1j * k * form(R['K6']))
    with np.errstate(divide='ignore', invalid='ignore'):
        return e_k, e_p, p, p / (e_k + e_p)


# ------------------------------------------------------------- comparison utilities


def match_eigenvalues(a, b):
    """One-to-one greedy nearest matching; returns (perm, max relative error).

    ``b[perm]`` is matched to ``a``.  Greedy on the globally smallest distances,
    which is exact whenever the two sets agree far better than their gaps.
    """
    a, b = np.asarray(a), np.asarray(b)
    d = abs(a[:, None] - b[None, :])
    perm = -np.ones(len(a), int)
    used = np.zeros(len(b), bool)
    for idx in np.argsort(d, axis=None):
        i, j = divmod(int(idx), len(b))
        if perm[i] < 0 and not used[j]:
            perm[i] = j
            used[j] = True
    rel = abs(a - b[perm]) / np.maximum(abs(a), 1e-300)
    return perm, rel.max()


def match_eigenvalues_large(a, b):
    """One-to-one nearest matching for thousands of eigenvalues (cfg4 / cfg5 fixtures): nearest
    neighbours through a k-d tree; the (rare) eigenvalues whose nearest partner is claimed twice are
    resolved by the greedy matcher on the left-over sets.  Returns (perm, max relative error)."""
    from scipy.spatial import cKDTree
    a, b = np.asarray(a), np.asarray(b)
    tree = cKDTree(np.c_[b.real, b.imag])
    _, nn = tree.query(np.c_[a.real, a.imag])
    perm = -np.ones(len(a), int)
    cnt = np.bincount(nn, minlength=len(b))
    good = cnt[nn] == 1
    perm[good] = nn[good]
    rest_a = np.where(~good)[0]
    if len(rest_a):
        rest_b = np.setdiff1d(np.arange(len(b)), perm[good])
        sub, _ = match_eigenvalues(a[rest_a], b[rest_b])
        perm[rest_a] = rest_b[sub]
    rel = abs(a - b[perm]) / np.maximum(abs(a), 1e-300)
    return perm, rel.max()


def eig_tolerance_from_kappa(val, kappa, norm2_balanced, rel=1e-8, backward=256.0):
    """eig_tolerance with the conditioning data stored in a fixture (tests/golden/big_*.npz)."""
    bound = backward * np.finfo(float).eps * np.asarray(kappa) * norm2_balanced
    return np.maximum(rel * abs(np.asarray(val)), bound)


def eig_tolerance(S, val, rel=1e-8, backward=256.0):
    """Per-eigenvalue parity tolerance: ``rel*|k|`` (the north-star bar) or, for eigenvalues
    whose own conditioning is worse than that, the first-order perturbation bound
    ``backward * eps * kappa_k * ||S_bal||_2`` of a backward-stable solver on the balanced
    matrix.  At the lowest frequencies of some sweeps LAPACK itself only reproduces its
    eigenvalues to 1e-5..1e-7 under 1e-15 perturbations of S ("parity bar")."""
    Sb, _ = la.matrix_balance(S, permute=False)
    w, vl, vr = la.eig(Sb, left=True, right=True)
    kappa = 1.0 / np.maximum(abs(np.sum(vl.conj() * vr, axis=0)), 1e-300)
    perm, _ = match_eigenvalues(val, w)
    bound = backward * np.finfo(float).eps * kappa[perm] * np.linalg.norm(Sb, 2)
    return np.maximum(rel * abs(val), bound)


def mode_shape_error(u_ref, u_new):
    """min over sign of ||u_new -/+ u_ref|| / ||u_ref|| per column."""
    nr = np.linalg.norm(u_ref, axis=0)
    e1 = np.linalg.norm(u_new - u_ref, axis=0) / nr
    e2 = np.linalg.norm(u_new + u_ref, axis=0) / nr
    return np.minimum(e1, e2)
