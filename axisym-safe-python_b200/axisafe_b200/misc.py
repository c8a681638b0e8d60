"""Material-constant conversions (host only, scalar arithmetic).

Mirrors the call surface of the reference's ``axisafe/misc.py`` (reference
file:line given per function) so that the example scripts run unchanged.
Nothing here touches the GPU.
"""
import numpy as np


def G_nu2lame(G, nu):
    """(shear modulus, Poisson ratio) -> (lambda, mu).  ref misc.py:21-33."""
    return 2.0 * G * nu / (1.0 - 2.0 * nu), G


def bulkvel2E(cl, ct, rho):
    """(c_L, c_S, rho) -> (Young modulus, Poisson ratio).  ref misc.py:35-51."""
    mu = ct ** 2 * rho
    ratio = cl ** 2 / ct ** 2
    nu = (2.0 - ratio) / (2.0 * (1.0 - ratio))
    return mu * 2.0 * (1.0 + nu), nu


def cLcS2lame(cl, ct, rho):
    """(c_L, c_S, rho) -> (lambda, mu).  ref misc.py:53-68."""
    return (cl ** 2 - 2 * ct ** 2) * rho, ct ** 2 * rho


def young2lame(E, nu):
    """(Young modulus, Poisson ratio) -> (lambda, mu).  ref misc.py:70-85."""
    return nu * E / (1 + nu) / (1 - 2 * nu), E / 2 / (1 + nu)


def lame2C(lame_1, lame_2):
    """Isotropic 6x6 stiffness in [rr, tt, zz, tz, rz, rt] order.  ref misc.py:109-128."""
    C = np.zeros([6, 6], 'complex')
    C[:3, :3] = lame_1
    for i in range(3):
        C[i, i] = lame_1 + 2 * lame_2
        C[3 + i, 3 + i] = lame_2
    return C


def young2C(E, nu):
    """Isotropic 6x6 stiffness from (E, nu).  ref misc.py:87-107."""
    return lame2C(*young2lame(E, nu))
