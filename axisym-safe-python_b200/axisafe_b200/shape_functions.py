"""Quadrature nodes/weights and Lagrange tables for 1-D spectral line elements.

Host-side table builder for the assembly kernel (``csrc/assemble.cu``): the
tables are tiny (m <= ~32), computed once per (family, m) and staged into
shared memory by the kernel.  The call surface mirrors the reference's
``axisafe/shape_functions.py`` (file:line cited per function); the
implementation is independent:

* GLL nodes are the extrema of P_{m-1}; found by vectorised Newton on
  (1-x^2) P'_{m-1}(x) using the Legendre three-term recurrence.
* GLJ(0,1) interior nodes are the zeros of P^{(1,2)}_{Q-2}; found by Newton with
  deflation from Chebyshev starting points.
* Differentiation matrices for BOTH families come from one barycentric routine
  (the reference has a triple-product formula for GLL and closed forms for GLJ;
  they agree with the barycentric result to a few ulp - SURVEY.md A.1).

Conventions kept from the reference: ``dN_nodal[a, q] = l_a'(xi_q)`` and
``N_nodal`` is the identity at the nodes.
"""
from math import factorial, lgamma, exp

import numpy as np

_EPS = np.finfo(float).eps


def jacobi(x, order, alpha, beta):
    """Jacobi polynomial P_order^{(alpha,beta)}(x) by the three-term recurrence.

    ref shape_functions.py:20-50.  Returns a 1-D array (length 1 for scalar x),
    like the reference.
    """
    xv = np.atleast_1d(np.asarray(x, dtype=float))
    p_prev = np.ones_like(xv)
    if order == 0:
        return p_prev
    p = 0.5 * (alpha - beta + (alpha + beta + 2.0) * xv)
    ab = alpha + beta
    for n in range(1, order):
        c = 2 * n + ab
        a1 = 2.0 * (n + 1) * (n + ab + 1) * c
        a2 = (c + 1.0) * (alpha * alpha - beta * beta)
        a3 = c * (c + 1.0) * (c + 2.0)
        a4 = 2.0 * (n + alpha) * (n + beta) * (c + 2.0)
        p_prev, p = p, ((a2 + a3 * xv) * p - a4 * p_prev) / a1
    return p


def djacobi(x, order, alpha, beta):
    """d/dx of P_order^{(alpha,beta)}(x).  ref shape_functions.py:52-67."""
    if order == 0:
        return np.zeros_like(np.atleast_1d(np.asarray(x, dtype=float)))
    return 0.5 * (alpha + beta + order + 1) * jacobi(x, order - 1, alpha + 1, beta + 1)


def jacobi_roots(order, alpha, beta):
    """Zeros of P_order^{(alpha,beta)}, ascending.  ref shape_functions.py:69-99."""
    roots = np.zeros(order)
    for k in range(order):
        r = -np.cos(np.pi * (2 * k + 1) / (2 * order))
        if k:
            r = 0.5 * (r + roots[k - 1])
        for _ in range(200):
            p = jacobi(r, order, alpha, beta)[0]
            dp = djacobi(r, order, alpha, beta)[0]
            defl = np.sum(1.0 / (r - roots[:k])) if k else 0.0
            step = -p / (dp - defl * p)
            r += step
            if abs(step) <= _EPS:
                break
        roots[k] = r
    return roots


def build_GLL_quadrature(no_of_nodes):
    """Gauss-Lobatto-Legendre nodes and weights on [-1, 1].

    ref shape_functions.py:101-140.  ``w_i = 2 / (p (p+1) P_p(x_i)^2)``, p = m-1.
    """
    p = no_of_nodes - 1
    x = -np.cos(np.pi * np.arange(p + 1) / p)        # ascending Chebyshev-Lobatto guess
    for _ in range(200):
        # Legendre P_{p-1}, P_p at x
        pm, pc = np.ones_like(x), x.copy()
        for k in range(2, p + 1):
            pm, pc = pc, ((2 * k - 1) * x * pc - (k - 1) * pm) / k
        # Newton on q(x) = x P_p - P_{p-1}  (zeros = GLL nodes);  q' = (p+1) P_p
        step = (x * pc - pm) / ((p + 1) * pc)
        x = x - step
        if np.max(np.abs(step)) <= _EPS:
            break
    pm, pc = np.ones_like(x), x.copy()
    for k in range(2, p + 1):
        pm, pc = pc, ((2 * k - 1) * x * pc - (k - 1) * pm) / k
    x[0], x[-1] = -1.0, 1.0
    w = 2.0 / (p * (p + 1) * pc ** 2)
    return x, w


def build_GLJ_quadrature(Q):
    """Gauss-Lobatto-Jacobi (alpha, beta) = (0, 1) nodes and weights.

    ref shape_functions.py:142-169 (Karniadakis & Sherwin, app. B.2).  The weights
    integrate against (1 + xi) d xi and sum to 2.
    """
    alpha, beta = 0, 1
    ksi = np.concatenate(([-1.0], jacobi_roots(Q - 2, alpha + 1, beta + 1), [1.0]))
    lg = lgamma(alpha + Q) + lgamma(beta + Q) - lgamma(alpha + beta + Q + 1)
    coeff = 2.0 ** (alpha + beta + 1) * exp(lg) / ((Q - 1) * factorial(Q - 1))
    w = coeff / jacobi(ksi, Q - 1, alpha, beta) ** 2
    w[0] *= beta + 1
    w[-1] *= alpha + 1
    return ksi, w


def _barycentric_derivative(ksi):
    """D[a, q] = l_a'(ksi_q) for the Lagrange basis through the points ksi."""
    m = len(ksi)
    diff = ksi[:, None] - ksi[None, :]
    np.fill_diagonal(diff, 1.0)
    lam = 1.0 / np.prod(diff, axis=1)                # barycentric weights
    D = np.zeros([m, m])
    for a in range(m):
        for q in range(m):
            if a != q:
                D[a, q] = (lam[a] / lam[q]) / (ksi[q] - ksi[a])
    for a in range(m):
        D[a, a] = np.sum(1.0 / np.delete(ksi[a] - ksi, a))
    return D


def lagrange_poly(ksi, order=0):
    """Lagrange interpolants through ``ksi`` and their derivatives at ``ksi``.

    ref shape_functions.py:171-205.  Only the nodal evaluation (order == 0 or
    order == len(ksi)) is on the hot path and supported.
    """
    ksi = np.asarray(ksi, dtype=float)
    if order not in (0, len(ksi)):
        raise NotImplementedError('lagrange_poly: only nodal tables are supported')
    return np.eye(len(ksi)), _barycentric_derivative(ksi)


def lagrange_GLJ(ksi, Q=0):
    """Same as :func:`lagrange_poly` on GLJ(0,1) nodes.  ref shape_functions.py:207-264."""
    ksi = np.asarray(ksi, dtype=float)
    if Q not in (0, len(ksi)):
        raise NotImplementedError('lagrange_GLJ: only nodal tables are supported')
    return np.eye(len(ksi)), _barycentric_derivative(ksi)


def line_spectral_GLL_lagrange(nodes):
    """(ksi, w, N, dN/dr, J) for a GLL line element with physical nodes ``nodes``.

    ref shape_functions.py:266-290.
    """
    nodes = np.asarray(nodes, dtype=float)
    ksi, weights = build_GLL_quadrature(len(nodes))
    N_nodal, dN_nodal = lagrange_poly(ksi)
    J = dN_nodal.T.dot(nodes)
    return ksi, weights, N_nodal, dN_nodal / J, J


def line_spectral_GLJ_lagrange(nodes):
    """(ksi, w, N, dN/dr, J) for a GLJ(0,1) line element.  ref shape_functions.py:292-315."""
    nodes = np.asarray(nodes, dtype=float)
    ksi, weights = build_GLJ_quadrature(len(nodes))
    N_nodal, dN_nodal = lagrange_GLJ(ksi)
    J = nodes.dot(dN_nodal)
    return ksi, weights, N_nodal, dN_nodal / J, J
