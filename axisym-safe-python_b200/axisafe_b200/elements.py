"""Element library of the SAFE sweep: solid / acoustic, regular / axis (core) / PML.

Host-side mirror of the reference's ``axisafe/elements.py`` classes
(``SLAX6, ALAX6, SLAX6_core, ALAX6_core, SLAX6_PML, ALAX6_PML``: same constructor
arguments, ``add_properties``, ``calculate_matrices`` and result attributes
``K_1..K_6, K_F, K_Ft, K_Fz, (K_F1..K_F3), M, Hs0/Hs1/Hf0/Hf1``).

The classes only carry the element *description* (nodes, material, PML profile,
DOF tables).  The quadrature itself - the reference's per-point Python loops,
elements.py:157-217 and friends - runs in the CUDA kernel ``axs_element_matrices``
(csrc/assemble.cu, one warp per element); ``calculate_matrices`` launches it for
this one element, ``Mesh.assemble_matrices`` launches it once for all elements.
No CPU fallback exists.
"""
import numpy as np

_SOLID, _FLUID = 0, 1
_GLL, _GLJ = 0, 1


def gamma_PML(x, gamma, PML_start, PML_thk):
    """Quadratic PML stretching profile, ``gamma`` = its mean value (ref elements.py:22-35)."""
    return 1 + 3 * (gamma - 1) * (abs(x - PML_start) / PML_thk) ** 2


def gamma_PML_exp(x, gamma, PML_start, PML_thk):
    """Exponential PML profile, ``gamma = a + 1j*b`` (ref elements.py:37-51)."""
    s = (x - PML_start) / PML_thk
    return np.exp(gamma.real * s) - 1j * (np.exp(gamma.imag * s) - 1)


class _LineElement(object):
    """Description of one 1-D spectral element; matrices come from the GPU."""
    _kind = _SOLID
    _family = _GLL
    _pml = False
    # raw kernel output order (csrc/assemble.cu)
    _RAW = ('K_1', 'K_2', 'K_3', 'K_4', 'K_5', 'K_6', 'K_F1', 'K_F2', 'K_F3', 'M')

    def __init__(self, nodes_at, n):
        self.nodes_at = nodes_at
        self.nodes_per_el = len(nodes_at)
        self.n = n
        per_node = 3 if self._kind == _SOLID else 1
        m = self.nodes_per_el
        self._full_dofs = per_node * m
        if self._kind == _SOLID:
            self.dofs_per_node = [3] * m
            self.dofs_domain = [['s'] * 3] * m
            self.T_components = [[1, 1j, 1j]] * m
            self.C = self.loss_factor = self.rho = None
        else:
            self.dofs_per_node = [1] * m
            self.dofs_domain = [['f']] * m
            self.T_components = [[1]] * m
            self.c_p = self.rho = None
        self._dropped = []                 # local DOFs removed by the axis conditions
        self.no_of_dofs = self._full_dofs
        self.dofs_per_el = sum(self.dofs_per_node)
        self._material = None
        self._pml_desc = (0, 0.0, 1.0, 0.0, 0.0)
        zeros = lambda dtype='complex': np.zeros([self._full_dofs] * 2, dtype=dtype)
        for name in ('K_Fz', 'K_Ft', 'K_F', 'K_1', 'K_2', 'K_3', 'K_4', 'K_5', 'K_6'):
            setattr(self, name, zeros())

    # ------------------------------------------------------------------ properties
    def add_properties(self, list_of_props):
        """Solid: [lambda, mu, rho, eta]; fluid: [K, rho, ...] (ref elements.py:103-121, :282-291)."""
        self._material = list(list_of_props)
        if self._kind == _SOLID:
            lam, mu = list_of_props[0], list_of_props[1]
            self.loss_factor = list_of_props[3]
            C = np.zeros([6, 6])
            C[:3, :3] = lam
            C[np.arange(3), np.arange(3)] = lam + 2 * mu
            C[np.arange(3, 6), np.arange(3, 6)] = mu
            self.C = (1 + self.loss_factor * 1.0j) * C
            self.rho = list_of_props[2]
        else:
            self.c_p = (list_of_props[0] / list_of_props[1]) ** 0.5
            self.rho = list_of_props[1]

    # ------------------------------------------------------------- kernel interface
    def _descriptor(self):
        """(ints, doubles) rows of the element table consumed by axs_element_matrices."""
        if self._material is None:
            raise ValueError('add_properties() must be called before calculate_matrices()')
        if self._kind == _SOLID:
            lam, mu, rho, eta = self._material[:4]
            mat = (lam, mu, eta, rho, 0.0)
        else:
            mat = (0.0, 0.0, 0.0, self.rho, self.c_p ** 2)
        return (self._kind, self._family, self._pml_desc[0], self.nodes_per_el), \
            mat + tuple(self._pml_desc[1:])

    def calculate_matrices(self):
        """Gauss-Lobatto quadrature of all element matrices on the GPU.

        ref elements.py:148-229 / :293-330 / :447-586 / :683-737 / :810-894 / :962-1007.
        """
        from . import _native
        _native.element_matrices([self])

    def _load(self, raw):
        """Install the kernel result ``raw[10, nd, nd]`` as the reference's attributes."""
        got = dict(zip(self._RAW, raw))
        keep = [i for i in range(self._full_dofs) if i not in self._dropped]
        cut = (lambda a: a[np.ix_(keep, keep)]) if self._dropped else (lambda a: a)
        if self._kind == _SOLID:
            for name in ('K_1', 'K_2', 'K_3', 'K_4', 'K_5', 'K_6'):
                setattr(self, name, cut(got[name]))
            if not self._pml:
                self.K_F1, self.K_F2, self.K_F3 = got['K_F1'], got['K_F2'], got['K_F3']
            self.K_F = cut(got['K_F1'].T.copy())
            self.K_Ft = cut(got['K_F3'].T.copy())
            self.K_Fz = cut(got['K_6'].copy())
            # the reference accumulates M of non-PML solids in complex64 (elements.py:99, :397);
            # the kernel reproduces the per-point fp32 rounding, the cast below is exact
            self.M = cut(got['M'] if self._pml else got['M'].astype(np.complex64))
            m = self.nodes_per_el
            first = np.zeros([3 * m, 1]); first[0, 0] = 2 * np.pi
            last = np.zeros([3 * m, 1]); last[3 * (m - 1), 0] = 2 * np.pi
            if self._family == _GLL:
                self.Hs0 = first
            if not self._pml:
                self.Hs1 = np.delete(last, self._dropped, 0) * self.nodes_at[-1]
        else:
            for name in ('K_1', 'K_5', 'K_6'):
                setattr(self, name, cut(got[name]))
            zero = np.zeros_like(self.K_1)
            self.K_2, self.K_3, self.K_4, self.K_F = zero, zero.copy(), zero.copy(), zero.copy()
            self.K_Fz = self.K_6.copy()
            self.K_Ft = self.K_5.copy()
            self.M = cut(got['M'] if self._pml else got['M'].real.copy())
            m = self.nodes_per_el
            first = np.zeros([m, 1]); first[0, 0] = 1.0
            last = np.zeros([m, 1]); last[m - 1, 0] = 1.0
            if self._family == _GLL:
                self.Hf0 = self.rho * first
            if not self._pml:
                # kept at full length even when the axis DOF is dropped (ref elements.py:727-737)
                self.Hf1 = self.rho * last * self.nodes_at[-1]
        if self._family == _GLJ and self._kind == _SOLID:
            self.no_of_dofs = self.K_1.shape[0] if 0 in self.nodes_at else None


class SLAX6(_LineElement):
    """Solid, 3 DOF/node (u_r, u_theta, u_z), GLL nodes (ref elements.py:54-229)."""

    def __init__(self, nodes_at, n):
        super(SLAX6, self).__init__(nodes_at, n)
        self.K_F1, self.K_F2, self.K_F3 = (np.zeros([self._full_dofs] * 2, 'complex') for _ in range(3))
        self.M = np.zeros([self._full_dofs] * 2, dtype='complex64')
        self.Hs0 = self.Hs1 = None


class ALAX6(_LineElement):
    """Acoustic fluid, 1 DOF/node (velocity potential), GLL nodes (ref elements.py:231-330)."""
    _kind = _FLUID

    def __init__(self, nodes_at, n):
        super(ALAX6, self).__init__(nodes_at, n)
        self.M = np.zeros([self._full_dofs] * 2, dtype='float64')
        self.Hf0 = self.Hf1 = None


class SLAX6_core(_LineElement):
    """Solid element touching the axis: GLJ(0,1) nodes, l'Hopital forms at r = 0 and the
    axis conditions per circumferential order (ref elements.py:332-586)."""
    _family = _GLJ

    def __init__(self, nodes_at, n):
        super(SLAX6_core, self).__init__(nodes_at, n)
        m = self.nodes_per_el
        if 0 in self.nodes_at:
            if n == 0:        # only u_z survives on the axis
                kept, self._dropped = [1j], [0, 1]
            elif n == 1:      # u_r, u_theta survive
                kept, self._dropped = [1, 1j], [2]
            else:             # n > 1: the axis node is clamped
                kept, self._dropped = [], [0, 1, 2]
            self.dofs_per_node = [len(kept)] + [3] * (m - 1)
            self.dofs_domain = [['s'] * len(kept)] + [['s'] * 3] * (m - 1)
            self.T_components = [kept] + [[1, 1j, 1j]] * (m - 1)
        self.dofs_per_el = sum(self.dofs_per_node)
        self.no_of_dofs = None
        self.K_F1, self.K_F2, self.K_F3 = (np.zeros([self._full_dofs] * 2, 'complex') for _ in range(3))
        self.M = np.zeros([self._full_dofs] * 2, dtype='complex64')
        self.Hs1 = None


class ALAX6_core(_LineElement):
    """Acoustic element touching the axis (ref elements.py:588-737)."""
    _kind = _FLUID
    _family = _GLJ

    def __init__(self, nodes_at, n):
        super(ALAX6_core, self).__init__(nodes_at, n)
        m = self.nodes_per_el
        if 0 in self.nodes_at and n != 0:
            self._dropped = [0]
            self.dofs_per_node = [0] + [1] * (m - 1)
            self.dofs_domain = [[]] + [['f']] * (m - 1)
            self.T_components = [[]] + [[1]] * (m - 1)
        self.dofs_per_el = sum(self.dofs_per_node)
        self.M = np.zeros([self._full_dofs] * 2, dtype='float64')
        self.Hf1 = None


class _PMLMixin(object):
    def _init_pml(self, PML_params, PML_function):
        self.PML_start = PML_params[0]
        self.PML_thk = PML_params[1]
        self.gamma = PML_params[2]
        self.PML_function = PML_function
        if PML_function is gamma_PML_exp:
            code = 1
        elif PML_function is gamma_PML:
            code = 2
        else:
            print('Error. No such PML function defined')       # ref elements.py:839
            raise ValueError('unsupported PML profile')
        g = complex(self.gamma)
        self._pml_desc = (code, float(self.PML_start), float(self.PML_thk), g.real, g.imag)
        self.M = np.zeros([self._full_dofs] * 2, dtype='complex')


class SLAX6_PML(_PMLMixin, _LineElement):
    """Solid PML element with complex radial stretching (ref elements.py:740-894)."""
    _pml = True

    def __init__(self, nodes_at, n, PML_params, PML_function=gamma_PML_exp):
        _LineElement.__init__(self, nodes_at, n)
        self._init_pml(PML_params, PML_function)
        self.Hs0 = None


class ALAX6_PML(_PMLMixin, _LineElement):
    """Acoustic PML element (ref elements.py:896-1007)."""
    _kind = _FLUID
    _pml = True

    def __init__(self, nodes_at, n, PML_params, PML_function=gamma_PML_exp):
        _LineElement.__init__(self, nodes_at, n)
        self._init_pml(PML_params, PML_function)
        self.Hf0 = None
