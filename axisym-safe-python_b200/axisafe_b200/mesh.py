"""Layered 1-D mesh, DOF numbering and global assembly for the SAFE sweep.

Host-side mirror of the reference's ``axisafe/mesh.py`` call surface
(``suggest_PML_parameters``, ``Mesh.create``, ``Mesh.assemble_matrices``,
``Mesh.assemble_subset`` and the public attributes listed in mesh.py:113-150).

Division of labour:

* mesh generation, DOF numbering and the tiny fluid-structure coupling blocks stay
  in Python (integer book-keeping, a few dozen entries);
* the Gauss-Lobatto quadrature of every element matrix (the reference's hot
  loop 1, mesh.py:488-489 -> elements.py) is ONE launch of the sm_100a kernel
  ``axs_element_matrices`` (csrc/assemble.cu), one warp per element;
* the global operators needed by the eigen-sweep are scatter-assembled on the
  device by ``axs_assemble_operators`` and stay resident in HBM
  (``Mesh._device_ops``); SciPy COO matrices ``K1..K6, M, KF, KFt, KFz,
  K1_coupling`` are materialised on the host from the element matrices for
  callers that read them (mesh.py:575-619).

There is no CPU fallback: ``assemble_matrices`` raises if the CUDA library or a
GPU is missing.
"""
from cmath import exp, pi

import numpy as np
import scipy.sparse as sps

from . import elements as ElementLibrary
from .shape_functions import build_GLL_quadrature, build_GLJ_quadrature

_MATRIX_ATTR = (('k1', 'K_1'), ('k2', 'K_2'), ('k3', 'K_3'), ('k4', 'K_4'), ('k5', 'K_5'),
                ('k6', 'K_6'), ('m', 'M'), ('kfz', 'K_Fz'), ('kft', 'K_Ft'), ('kf', 'K_F'))
_GLOBAL_NAME = dict(k1='K1', k2='K2', k3='K3', k4='K4', k5='K5', k6='K6', m='M',
                    kfz='KFz', kft='KFt', kf='KF', k1_coupling='K1_coupling')


def suggest_PML_parameters(PML_start, wavelenghts_within, waveguide_material,
                           surr_material, low_f, high_f, alpha=6, beta=7, att=1):
    """Heuristic thickness / order / reflection coefficients of an exponential PML.

    ref mesh.py:26-108.  Pure scalar pre-processing, called once by every example
    script; closed-form arithmetic on the radial wavenumbers of the bulk waves of
    the surrounding medium at the two ends of the frequency band.

    Returns (h_long, h_short, order, R_short, R_long, R_short_at_long).
    """
    omega = 2 * pi * np.array([low_f, high_f])
    lam, mu, rho = waveguide_material[:3]
    k_guide = [omega / ((lam + 2 * mu) / rho) ** 0.5 - 1j * att,       # longitudinal
               omega / (mu / rho) ** 0.5 - 1j * att]                    # shear
    if len(surr_material) == 4:
        ls, ms, rs = surr_material[:3]
        k_surr = [omega / ((ls + 2 * ms) / rs) ** 0.5, omega / (ms / rs) ** 0.5]
    else:
        k_surr = [omega / (surr_material[0] / surr_material[1]) ** 0.5]
    # radial wavenumbers for every (surrounding wave, guided wave) combination; the
    # reference orders them LL, LS, SL, SS and python's max/min keep the first extremum
    radial = [(ks ** 2 - kg ** 2) ** 0.5 for ks in k_surr for kg in k_guide]
    kr_short = max([kr[1] for kr in radial])
    kr_long = min([kr[0] for kr in radial])

    stretch = (exp(alpha) - 1) / alpha

    def thickness(kr):
        kp = kr.real * stretch
        disc = 4 * pi ** 2 * wavelenghts_within ** 2 / kp ** 2 + \
            8 * pi * PML_start * wavelenghts_within / kp
        return (2 * pi * wavelenghts_within / kp + disc ** 0.5) / 2

    h_long, h_short = thickness(kr_long), thickness(kr_short)
    damp = (beta - exp(beta) + 1) / beta

    def reflection(h, kr):
        return np.exp(2 * h * kr.imag * stretch) * np.exp(2 * h * kr.real * damp)

    order = int(np.ceil(pi * wavelenghts_within + 3))
    return (h_long.real, h_short.real, order, reflection(h_short, kr_short).real,
            reflection(h_long, kr_long).real, reflection(h_short, kr_long).real)


def _is_interface(types):
    """True when a node joins an acoustic ('A...') and a structural ('S...') element."""
    return len(set(types)) == 2 and types[0][0] != types[1][0]


class Mesh(object):
    """Radial mesh of an axisymmetric layered waveguide (ref mesh.py:110-150)."""

    def __init__(self):
        self.element_types, self.element_size, self.min_wavelengths = dict(), [], []
        self.glob_nodes, self.IEN, self.element_sets = dict(), dict(), dict()
        self.rev_PMLs = dict()
        self.element_props = dict()
        self.node_sets = dict()
        self.uncoupled_nodes, self.uncoupled_dofs = [], []
        self.nodes_types = dict()
        self.nodes_in_element = dict()
        self.struct_acoust = dict()
        self.struct_acoust_pairs = []
        self.n = None
        self.PML_props = None
        self.elements = dict()
        self.is_coupled = None
        self.dofs_per_elset = dict()
        self.no_of_nodes = 0
        self.no_of_dofs = 0
        self.ID = dict()
        self.dof_domains = dict()
        self.LM = dict()
        self.K1 = self.K2 = self.K3 = self.K4 = self.K5 = self.K6 = None
        self.KF = self.KFt = self.KFz = self.M = None
        self.Tdiag = None
        self.K1_coupling = None
        self._device_ops = None          # banded operators resident on the GPU
        self._device_elem = None         # per-element matrices resident on the GPU

    # ------------------------------------------------------------------ generation
    def create(self, sets, max_frequency, lubricated_coupling=0, PML=0, PML_props=0,
               domain='axisymmetric'):
        """Generate nodes and connectivity layer by layer (ref mesh.py:153-288).

        ``sets`` rows: [name, r_start, thickness, material, element type,
        (order), (number of elements)]; solids carry [lambda, mu, rho, eta], fluids
        [K, rho, eta].  Elements touching r = 0 use GLJ(0,1) nodes, all others GLL.
        """
        next_label, element = 1, 1
        r_cursor = sets[0][1]
        for layer in sets:
            name, _, thickness, material, etype = layer[:5]
            if len(material) == 4:
                slowest = (material[1] / material[2]) ** 0.5
            elif len(material) == 3:
                slowest = (material[0] / material[1]) ** 0.5
            self.min_wavelengths.append(slowest / max_frequency)
            order = layer[5] if len(layer) > 5 and layer[5] != 0 else \
                int(np.ceil(np.pi * thickness / self.min_wavelengths[-1]) + 3)
            count = layer[6] if len(layer) > 6 and layer[6] != 0 else 1
            self.element_size.append(thickness / count)
            members = []
            for local in range(count):
                on_axis = r_cursor == 0.0 and domain == 'axisymmetric'
                ksi, _ = (build_GLJ_quadrature if on_axis else build_GLL_quadrature)(order + 1)
                unit = (ksi + 1) * 0.5
                if lubricated_coupling != 0 and lubricated_coupling[1] == name and local == 0:
                    next_label += 1          # duplicate the interface node
                positions = list(r_cursor + unit * self.element_size[-1])
                labels = list(range(next_label, next_label + len(positions)))
                next_label += len(positions) - 1
                for label, position in zip(labels, positions):
                    self.glob_nodes.setdefault(label, [position])
                self.IEN[element] = labels
                self.element_types[element] = etype
                self.element_props[element] = {'material': material}
                members.append(element)
                element += 1
                r_cursor = positions[-1]
            self.element_sets[name] = members
            self.node_sets[name] = [node for el in members for node in self.IEN[el]]
        if PML != 0:
            self.element_sets['PML'] = self.element_sets[PML]
            self.PML_props = PML_props
            for el in self.element_sets['PML']:
                self.rev_PMLs[el] = 'PML'
        if lubricated_coupling != 0:
            inner = self.element_sets[lubricated_coupling[0]][-1]
            outer = self.element_sets[lubricated_coupling[1]][0]
            self.uncoupled_nodes.append([self.IEN[inner][-1], self.IEN[outer][0]])
            self.uncoupled_dofs.append([1, 2])
        for el in sorted(self.IEN):
            for node in self.IEN[el]:
                self.nodes_types.setdefault(node, []).append(self.element_types[el])
                self.nodes_in_element.setdefault(node, []).append(el)
        for node, types in self.nodes_types.items():
            if len(types) == 2 and types[0] != types[1] and types[0][0] != types[1][0]:
                self.struct_acoust_pairs.append(sorted(self.nodes_in_element[node]))
                kind = None
                if 'SL' in types[0] and 'AL' in types[1]:
                    kind = 'solid-fluid'
                elif 'AL' in types[0] and 'SL' in types[1]:
                    kind = 'fluid-solid'
                if kind:
                    for el in self.nodes_in_element[node]:
                        self.struct_acoust[el] = kind

    # --------------------------------------------------------------- host numbering
    def _instantiate_elements(self, n):
        """Element objects by type; elements with a node on the axis become *_core
        (ref mesh.py:307-335)."""
        self.elements = dict()
        for el in sorted(self.IEN):
            nodes_at = np.array([self.glob_nodes[int(node)][0] for node in self.IEN[el]])
            base = self.element_types[el]
            if base in ('SLAX6', 'SLAX6_core', 'ALAX6', 'ALAX6_core'):
                stem = base[:5]
                if 0 in nodes_at:
                    self.element_types[el] = stem + '_core'
                    obj = getattr(ElementLibrary, stem + '_core')(nodes_at, n)
                else:
                    obj = getattr(ElementLibrary, stem)(nodes_at, n)
            elif base in ('SLAX6_PML', 'ALAX6_PML'):
                obj = getattr(ElementLibrary, base)(nodes_at, n, self.PML_props)
            else:
                raise ValueError('unknown element type ' + repr(base))
            obj.add_properties(self.element_props[el]['material'])
            self.elements[el] = obj

    def _number_dofs(self):
        """Global DOF ids per node, T diagonal, domain lists and LM (ref mesh.py:337-464)."""
        t_diag, solid, fluid = [], [], []
        next_id = 0

        def take(el, idx, node, append):
            nonlocal next_id
            obj = self.elements[el]
            fresh = [next_id + i for i in range(obj.dofs_per_node[idx])]
            if append:
                self.ID[int(node)].extend(fresh)
            else:
                self.ID[int(node)] = fresh
            t_diag.extend(obj.T_components[idx])
            for dof, domain in zip(fresh, obj.dofs_domain[idx]):
                (solid if domain == 's' else fluid).append(dof)
            next_id += obj.dofs_per_node[idx]

        for el in sorted(self.IEN):
            for idx, node in enumerate(self.IEN[el]):
                shared_media = _is_interface(self.nodes_types[node])
                if node in self.ID and not shared_media:
                    continue
                second_visit = len(set(self.nodes_types[node])) == 2 and \
                    self.nodes_in_element[node].index(el) == 1
                if self.uncoupled_nodes == [] or self.element_types[el] != 'SLAX6':
                    take(el, idx, node, second_visit)
                    continue
                # lubricated (radial-only) contact, SLAX6 elements only (ref mesh.py:377-422)
                for case_no, case in enumerate(self.uncoupled_nodes):
                    if node not in case:
                        take(el, idx, node, second_visit)
                    elif case.index(node) == 0:
                        take(el, idx, node, False)
                    elif self.uncoupled_dofs[case_no] == [1, 2]:
                        radial = self.ID[case[0]][0]
                        self.ID[int(node)] = [radial, next_id, next_id + 1]
                        solid.extend([next_id, next_id + 1])
                        next_id += 2
                        t_diag.extend([1j, 1j])
        self.no_of_dofs = next_id
        self.no_of_nodes = len(self.glob_nodes)
        self.Tdiag = t_diag
        self.dof_domains.update({'solid': solid, 'fluid': fluid})

        for el, nodes in self.IEN.items():
            sequence = []
            for node in nodes:
                owners = self.nodes_in_element[node]
                if owners not in self.struct_acoust_pairs:
                    sequence.extend(self.ID[node])
                elif el == owners[0]:
                    sequence.extend(self.ID[node][:self.elements[el].dofs_per_node[-1]])
                else:
                    sequence.extend(self.ID[node][self.elements[el - 1].dofs_per_node[-1]:])
            self.LM[el] = sequence
        self.dofs_per_elset = {name: list(set(dof for el in members for dof in self.LM[el]))
                               for name, members in self.element_sets.items()}
        self.is_coupled = len(self.struct_acoust) > 0

    def _coupling_blocks(self):
        """Fluid-structure blocks K_1_coupling from the interface vectors (ref mesh.py:493-514)."""
        for pair in self.struct_acoust_pairs:
            pair.sort()
            first, second = self.elements[pair[0]], self.elements[pair[1]]
            kind = self.element_types[pair[0]]
            if kind in ('ALAX6', 'ALAX6_core'):
                H = -second.Hs0.dot(first.Hf1.T)
                nf = H.shape[1]
                block = np.zeros([nf + H.shape[0]] * 2)
                block[:nf, nf:], block[nf:, :nf] = H.T, H
                first.K_1_coupling = block
            if kind in ('SLAX6', 'SLAX6_core'):
                H = first.Hs1.dot(second.Hf0.T)
                ns = H.shape[0]
                block = np.zeros([ns + H.shape[1]] * 2)
                block[:ns, ns:], block[ns:, :ns] = H, H.T
                first.K_1_coupling = block

    def _matrix_ids(self):
        ids = [name for name, _ in _MATRIX_ATTR]
        return ids + ['k1_coupling'] if self.is_coupled else ids

    def _scatter(self, element_list):
        """COO triplets of every global matrix for the listed elements (ref mesh.py:518-572)."""
        triplets = {name: ([], [], []) for name in self._matrix_ids()}
        leaders = {pair[0] for pair in self.struct_acoust_pairs}
        for el in element_list:
            obj = self.elements[el]
            for name, attr in _MATRIX_ATTR:
                block = getattr(obj, attr)
                lm = np.asarray(self.LM[el], dtype=np.int64)
                rows, cols = np.nonzero(block)
                store = triplets[name]
                store[0].append(lm[rows]); store[1].append(lm[cols]); store[2].append(block[rows, cols])
            if self.is_coupled and el in self.struct_acoust and el in leaders:
                block = obj.K_1_coupling
                lm = np.asarray(self.LM[el] + self.LM[el + 1], dtype=np.int64)
                rows, cols = np.nonzero(block)
                store = triplets['k1_coupling']
                store[0].append(lm[rows]); store[1].append(lm[cols]); store[2].append(block[rows, cols])
        shape = (self.no_of_dofs, self.no_of_dofs)
        out = {}
        for name, (rows, cols, data) in triplets.items():
            if rows:
                rows, cols, data = np.concatenate(rows), np.concatenate(cols), np.concatenate(data)
            else:
                rows, cols, data = np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0)
            out[name] = sps.coo_matrix((data, (rows, cols)), shape=shape)
        return out

    # --------------------------------------------------------------------- assembly
    def assemble_matrices(self, n='none'):
        """Evaluate all element matrices on the GPU and assemble the global system.

        ref mesh.py:290-619.  Like the reference, not idempotent: use one ``Mesh``
        per circumferential order ``n``.
        """
        from . import _native
        self.n = n
        self._instantiate_elements(n)
        self._number_dofs()
        order = sorted(self.IEN)
        self._device_elem = _native.element_matrices([self.elements[el] for el in order])
        self._coupling_blocks()
        for name, matrix in self._scatter(list(self.LM)).items():
            setattr(self, _GLOBAL_NAME[name], matrix)
        self._device_ops = _native.assemble_operators(self, order, self._device_elem)

    def assemble_subset(self, elements):
        """Global-size matrices restricted to a subset of elements (ref mesh.py:621-755)."""
        return {_GLOBAL_NAME[name]: matrix for name, matrix in self._scatter(elements).items()}
