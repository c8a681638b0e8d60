"""Benchmark / parity cases = the inputs of the reference's example scripts, as data.

Each builder returns ``dict(name, sets, fmax, PML, PML_props, n, f)`` ready for
``Mesh.create(sets, fmax, PML=..., PML_props=...)``, ``assemble_matrices(n)`` and
``solve(f)``.  Definitions follow BASELINE.md section 3 / SURVEY.md section 8d and cite
the example script they restate.  Pure host code.
"""
import numpy as np

from . import misc
from .mesh import suggest_PML_parameters


def _case(name, sets, f, n, PML=0, PML_props=0):
    return dict(name=name, sets=sets, fmax=f[-1], PML=PML, PML_props=PML_props, n=n, f=f)


def pipe_soil_pml(nf=150, n=0):
    """cfg1 - water-filled cast-iron pipe buried in soil (examples/water_filled_pipe_in_soil.py:23-65)."""
    cast_iron = list(misc.young2lame(156e9, 0.2)) + [7800, 0.0]
    soil = list(misc.cLcS2lame(200, 100, 1500)) + [1500, 0.]
    water = [2.25e9, 1000, 0]
    f = np.linspace(1e3, 20e3, nf, True)
    r_i, wall = 0.1, 0.01
    _, h_short, order, _, _, _ = suggest_PML_parameters(r_i + wall, 3, cast_iron, soil, 1e3, 20e3, att=1)
    thk = np.round(h_short, 4)
    sets = [['water', 0.0, r_i, water, 'ALAX6'],
            ['cast_iron', r_i, wall, cast_iron, 'SLAX6'],
            ['soil', r_i + wall, thk, soil, 'SLAX6_PML', order]]
    return _case('pipe_soil_pml', sets, f, n, 'soil', [r_i + wall, thk, 6 + 7j])


def free_steel_pipe(nf=500, n=0):
    """cfg2 - free steel pipe in vacuum, one solid layer (steel constants of Nguyen.py:23-24)."""
    steel = list(misc.cLcS2lame(5960, 3260, 7932)) + [7932, 0.0]
    f = np.linspace(1e3, 100e3, nf, True)
    return _case('free_steel_pipe', [['steel', 0.1, 0.01, steel, 'SLAX6']], f, n)


def aristegui(nf=200, n=0):
    """cfg3 - water-filled copper pipe submerged in water, fluid PML (Aristegui.py:22-38)."""
    copper = list(misc.cLcS2lame(4759, 2325, 8933)) + [8933, 0.00]
    water = [2.25e9, 1000, 0]
    _, h_short, order, _, _, _ = suggest_PML_parameters(0.0075, 8, copper, water, 1e3, 4e5, att=1)
    thk = np.round(h_short, 4)
    sets = [['water', 0.0, 0.0068, water, 'ALAX6'],
            ['copper', 0.0068, 0.0007, copper, 'SLAX6'],
            ['water_a', 0.0075, thk, water, 'ALAX6_PML', order]]
    return _case('aristegui', sets, np.linspace(1e3, 4e5, nf, True), n, 'water_a', [0.0075, thk, 6 + 7j])


def muggleton_submerged(nf=100, n=0):
    """cfg3b - water-filled MDPE pipe in water (Muggleton_submerged_experiment_spectral.py:22-38)."""
    mdpe = list(misc.young2lame(2e9, 0.4)) + [900, 0.06]
    water = [2.25e9, 1000, 0]
    _, h_short, order, _, _, _ = suggest_PML_parameters(0.09, 3, mdpe, water, 1, 1e3, att=1)
    thk = np.round(h_short, 4)
    sets = [['water', 0.0, 0.079, water, 'ALAX6'],
            ['mdpe', 0.079, 0.011, mdpe, 'SLAX6'],
            ['water_a', 0.09, thk, water, 'ALAX6_PML', order]]
    return _case('muggleton_submerged', sets, np.linspace(1, 1000, nf, True), n, 'water_a',
                 [0.09, thk, 6 + 7j])


def muggleton_buried(nf=50, n=0):
    """Water-filled MDPE pipe in soil (Muggleton_experiment_spectral.py:22-45)."""
    mdpe = list(misc.young2lame(2e9, 0.4)) + [900, 0.06]
    soil = list(misc.cLcS2lame(200, 100, 1500)) + [1500, 0.]
    water = [2.25e9, 1000, 0]
    _, h_short, order, _, _, _ = suggest_PML_parameters(0.09, 3, mdpe, soil, 1, 1e3, att=1)
    thk = np.round(h_short, 4)
    sets = [['water', 0.0, 0.079, water, 'ALAX6'],
            ['mdpe', 0.079, 0.011, mdpe, 'SLAX6'],
            ['soil', 0.09, thk, soil, 'SLAX6_PML', order]]
    return _case('muggleton_buried', sets, np.linspace(1, 1000, nf, True), n, 'soil', [0.09, thk, 6 + 7j])


def gravenkamp(nf=100, n=0):
    """Titanium rod in oil, solid core + fluid PML (Gravenkamp_titanum_cylinder.py:22-37)."""
    titanium = list(misc.G_nu2lame(46.53e9, 0.302)) + [4460, 0.0]
    oil = [2.634e9, 870, 0]
    _, h_short, order, _, _, _ = suggest_PML_parameters(0.001, 3, titanium, oil, 1e3, 3e6, att=1)
    thk = np.round(h_short, 4)
    sets = [['titanium', 0.0, 0.001, titanium, 'SLAX6'],
            ['oil', 0.001, thk, oil, 'ALAX6_PML', order]]
    return _case('gravenkamp', sets, np.linspace(1e3, 3e6, nf, True), n, 'oil', [0.001, thk, 6 + 7j])


def nguyen(nf=100, n=0):
    """Steel rod embedded in concrete (Nguyen.py:23-53)."""
    steel = list(misc.cLcS2lame(5960, 3260, 7932)) + [7932, 0.000]
    concrete = list(misc.cLcS2lame(4222.1, 2637.5, 2300)) + [2300, 0.000]
    _, h_short, order, _, _, _ = suggest_PML_parameters(0.01, 6, steel, concrete, 100, 2e5,
                                                         att=10, alpha=6, beta=7)
    thk = np.round(h_short, 4)
    sets = [['steel', 0, 0.01, steel, 'SLAX6'],
            ['concrete', 0.01, thk, concrete, 'SLAX6_PML', order]]
    return _case('nguyen', sets, np.linspace(1, 2e5, nf, True), n, 'concrete', [0.01, thk, 6 + 7j])


def coated_pipe(nf=2000, n=1, elements_per_layer=22, order=10):
    """cfg4 - steel + bitumen + soil PML, every layer ``order`` x ``elements_per_layer``
    (2N = 3966 at the defaults; synthetic)."""
    steel = list(misc.cLcS2lame(5960, 3260, 7932)) + [7932, 0.0]
    bitumen = list(misc.cLcS2lame(1860, 750, 1500)) + [1500, 0.05]
    soil = list(misc.cLcS2lame(200, 100, 1500)) + [1500, 0.0]
    sets = [['steel', 0.1, 0.01, steel, 'SLAX6', order, elements_per_layer],
            ['bitumen', 0.11, 0.005, bitumen, 'SLAX6', order, elements_per_layer],
            ['soil', 0.115, 0.02, soil, 'SLAX6_PML', order, elements_per_layer]]
    return _case('coated_pipe', sets, np.linspace(1e3, 50e3, nf, True), n, 'soil', [0.115, 0.02, 6 + 7j])


def ten_layer(nf=10000, n=1, elements_per_layer=13, order=10):
    """cfg5 - nine alternating steel/bitumen layers + soil PML (2N = 7806 at the defaults)."""
    steel = list(misc.cLcS2lame(5960, 3260, 7932)) + [7932, 0.0]
    bitumen = list(misc.cLcS2lame(1860, 750, 1500)) + [1500, 0.05]
    soil = list(misc.cLcS2lame(200, 100, 1500)) + [1500, 0.0]
    sets, r = [], 0.1
    for i in range(9):
        sets.append(['L%d' % i, r, 0.005, steel if i % 2 == 0 else bitumen, 'SLAX6', order, elements_per_layer])
        r = 0.1 + 0.005 * (i + 1)
    sets.append(['soil', r, 0.02, soil, 'SLAX6_PML', order, elements_per_layer])
    return _case('ten_layer', sets, np.linspace(1e3, 50e3, nf, True), n, 'soil', [r, 0.02, 6 + 7j])


ALL = dict(pipe_soil_pml=pipe_soil_pml, free_steel_pipe=free_steel_pipe, aristegui=aristegui,
           muggleton_submerged=muggleton_submerged, muggleton_buried=muggleton_buried,
           gravenkamp=gravenkamp, nguyen=nguyen, coated_pipe=coated_pipe, ten_layer=ten_layer)
